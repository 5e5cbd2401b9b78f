"""packed (routing_fused2.cu) vs scalar (routing_fused.cu) cluster-resident kernels on the same inputs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import _lib
dev = torch.device("cuda:0")

def inputs(BP, K, I, O, seed=0, zero_frac=0.1):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(BP, 1, 1, 4, generator=g)
    lrfs = base + 0.3 * torch.randn(BP, K, I, 4, generator=g)
    lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
    lrfs = lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)
    act = torch.rand(BP, K * I, generator=g) * 0.5 + 0.4
    dead = torch.rand(BP, K * I, generator=g) < zero_frac
    act[dead] = 0
    lrfs.view(BP, K * I, 4)[dead] = 0
    tb = torch.randn(1, O, 4, 1, generator=g)
    t = (tb + 0.25 * torch.randn(BP * K, O, 4, I, generator=g)).reshape(BP * K, 4 * I * O)
    return t.to(dev).contiguous(), lrfs.to(dev).contiguous(), act.to(dev).contiguous()

def fwd(name, t, lrfs, act, BP, K, I, O, T):
    alpha = torch.ones(1, device=dev); beta = torch.ones(1, device=dev)
    Nc = BP * O
    pose = torch.full((Nc, 4), 7.0, device=dev); agr = torch.full((Nc,), 7.0, device=dev)
    poses_all = torch.full((T, Nc, 4), 7.0, device=dev); stats = torch.full((Nc, 2), 7.0, device=dev)
    _lib.call(name, t.data_ptr(), lrfs.data_ptr(), act.data_ptr(), alpha.data_ptr(), beta.data_ptr(), pose.data_ptr(),
              agr.data_ptr(), poses_all.data_ptr(), stats.data_ptr(), BP, K, I, O, T, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return pose, agr, poses_all, stats

def bwd(name, t, lrfs, act, BP, K, I, O, T, fw, seed=1, chain=True):
    pose, agr, poses_all, stats = fw
    g = torch.Generator().manual_seed(seed)
    Nc = BP * O
    g_pose = torch.randn(Nc, 4, generator=g).to(dev); g_agr = torch.randn(Nc, generator=g).to(dev)
    alpha = torch.ones(1, device=dev); beta = torch.ones(1, device=dev)
    g_t = torch.full_like(t, 7.0); g_l = torch.zeros_like(lrfs); g_a = torch.zeros_like(act); g_ab = torch.zeros(2, device=dev)
    _lib.call(name, t.data_ptr(), lrfs.data_ptr(), act.data_ptr(), alpha.data_ptr(), beta.data_ptr(), poses_all.data_ptr(),
              stats.data_ptr(), g_pose.data_ptr(), g_agr.data_ptr(), g_t.data_ptr(), g_l.data_ptr(),
              g_a.data_ptr() if chain else None, g_ab.data_ptr(), BP, K, I, O, T, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return g_t, g_l, g_a, g_ab

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

worst = 0.0
shapes = [(32, 256, 128, 40, 3), (3, 256, 128, 5, 3), (2, 16, 128, 7, 3), (5, 9, 128, 3, 2), (2, 256, 64, 3, 3),
          (3, 100, 30, 4, 3), (2, 37, 6, 5, 1), (4, 512, 8, 6, 3), (2, 255, 126, 3, 3), (1, 1, 2, 1, 3)]
do_bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
for (BP, K, I, O, T) in shapes:
    t, lrfs, act = inputs(BP, K, I, O)
    if not _lib.query("qec_routing_fused_packed_supported", K, I, O, T):
        print("skip (packed unsupported)", (BP, K, I, O, T)); continue
    a = fwd("qec_routing_fused_fwd", t, lrfs, act, BP, K, I, O, T)
    b = fwd("qec_routing_fused_fwd_scalar", t, lrfs, act, BP, K, I, O, T)
    errs = [rel(x, y) for x, y in zip(a, b)]
    line = "fwd pose %.2e agr %.2e poses_all %.2e stats %.2e" % tuple(errs)
    worst = max(worst, *errs)
    if do_bwd:
        for chain in (True, False):
            ga = bwd("qec_routing_fused_bwd", t, lrfs, act, BP, K, I, O, T, b, chain=chain)
            gb = bwd("qec_routing_fused_bwd_scalar", t, lrfs, act, BP, K, I, O, T, b, chain=chain)
            e2 = [rel(x, y) for x, y in zip(ga, gb)]
            line += " | bwd%s g_t %.2e g_l %.2e g_a %.2e g_ab %.2e" % ((("" if chain else "(nochain)"),) + tuple(e2))
            worst = max(worst, *[e for e in e2 if e == e])
    print((BP, K, I, O, T), line)
print("WORST", worst)
