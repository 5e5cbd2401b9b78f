"""determinism stress: the fused kernels' forward (and the non-atomic part of the backward) must be
bit-identical run to run; any difference is a race"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
shapes = [(2, 1, 64, 32, 5, 3), (1, 2, 100, 37, 3, 3), (1, 1, 256, 24, 4, 3), (1, 1, 250, 50, 3, 2),
          (2, 1, 256, 128, 40, 3), (1, 1, 256, 128, 6, 1), (1, 1, 250, 100, 3, 3)]
bad = 0
for (B, P, K, I, O, T) in shapes:
    g = torch.Generator().manual_seed(K + I)
    base = torch.randn(B, 1, 1, 1, 4, generator=g)
    lrfs = base + 0.3 * torch.randn(B, P, K, I, 4, generator=g)
    lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
    lrfs = (lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)).to(dev).requires_grad_(True)
    act = ((torch.rand(B, P, K * I, generator=g) * 0.5 + 0.4) * (torch.rand(B, P, K * I, generator=g) < 0.9)).to(dev).requires_grad_(True)
    tb = torch.randn(1, O, 4, 1, generator=g)
    t = (tb + 0.25 * torch.randn(B * P * K, O, 4, I, generator=g)).reshape(B * P * K, 4 * I * O).to(dev).requires_grad_(True)
    alpha = torch.ones(1, device=dev, requires_grad=True); beta = torch.ones(1, device=dev, requires_grad=True)
    Gp = torch.randn(B, P, O, 4, generator=g).to(dev); Ga = torch.randn(B, P, O, generator=g).to(dev)
    ref = None
    for r in range(reps):
        for x in (lrfs, act, t, alpha, beta): x.grad = None
        pose, agr = QF.routing_fused(t, lrfs, act, alpha, beta, T)
        ((pose * Gp).sum() + (agr * Ga).sum()).backward()
        cur = (pose.detach().clone(), agr.detach().clone(), t.grad.clone(), lrfs.grad.clone(), act.grad.clone())
        if ref is None: ref = cur; continue
        for name, a, b, exact in (("pose", cur[0], ref[0], True), ("agr", cur[1], ref[1], True), ("g_t", cur[2], ref[2], True),
                                  ("g_lrf", cur[3], ref[3], False), ("g_act", cur[4], ref[4], False)):
            d = float((a - b).abs().max())
            lim = 0.0 if exact else 1e-5 * float(b.abs().max())
            if d > lim:
                bad += 1
                print("MISMATCH", (B, P, K, I, O, T), "rep", r, name, d)
    print("shape", (B, P, K, I, O, T), "done")
print("bad", bad)
