"""bit-reproducibility hunt on the cluster-resident routing backward (pool2 geometry of BASELINE config 2): the same
launch repeated N times; dL/dt must be bit-identical (dL/dlrfs, dL/dact and dL/d(alpha,beta) use FP32 atomics and may
differ in the last bits).  Prints where a deviating repetition differs: (sample, neighbour k) row, (capsule o,
component q, channel i) column.   usage: python scripts/race_hunt_bwd.py [reps] [B] [split]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
split = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
dev = torch.device("cuda:0")
P, K, I, O, T = 1, 256, 128, 40, 3
g = torch.Generator().manual_seed(3)
base = torch.randn(B, 1, 1, 1, 4, generator=g)
lrfs = base + 0.3 * torch.randn(B, P, K, I, 4, generator=g)
lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
lrfs = (lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)).to(dev)
act = (torch.rand(B, P, K * I, generator=g) * 0.5 + 0.4).to(dev)
tb = torch.randn(1, O, 4, 1, generator=g)
t = (tb + 0.25 * torch.randn(B * P * K, O, 4, I, generator=g)).reshape(B * P * K, 4 * I * O).to(dev)
alpha = torch.ones(1, device=dev); beta = torch.ones(1, device=dev)
Gp = torch.randn(B, P, O, 4, generator=g).to(dev); Ga = torch.randn(B, P, O, generator=g).to(dev)
pose, agr, poses_all, stats = QF.routing_fused_forward_raw(t, lrfs, act, alpha, beta, T)
fwd_ref = [x.clone() for x in (pose, agr, poses_all, stats)]
bad_fwd = 0
h = torch.randn(B * P * K, O, generator=g).to(dev)
W2 = (torch.randn(4 * I * O, O, generator=g) / 6).to(dev)
perm = torch.randperm(4 * I * O, generator=g).to(dev)
ref2 = None
bad2 = 0
ref = None
bad = 0
for r in range(reps):
    fw = QF.routing_fused_forward_raw(t, lrfs, act, alpha, beta, T)   # the forward has no atomics at all
    if not all(torch.equal(x, y) for x, y in zip(fw, fwd_ref)):
        bad_fwd += 1
        print("rep %d: forward outputs differ: %s" % (r, [int((x != y).sum()) for x, y in zip(fw, fwd_ref)]))
    gt, gl, ga, gab = QF.routing_fused_backward_raw(t, lrfs, act, alpha, beta, poses_all, stats, Gp, Ga, T, True, True, split=split)
    if not split and QF.gen2_bwd_supported(gt.shape[0], gt.shape[1], O):
        outs = QF.gen2_backward(gt, h, W2, perm)   # deterministic by design: any deviation with an identical g is a race
        if ref2 is None:
            ref2 = [x.clone() for x in outs]
        elif not all(torch.equal(x, y) for x, y in zip(outs, ref2)):
            bad2 += 1
            same_g = ref is not None and torch.equal(gt, ref[0])
            msg = []
            for name, x, y in zip(("g_h", "g_W2", "g_b2"), outs, ref2):
                if not torch.equal(x, y):
                    d = (x != y).nonzero()
                    msg.append("%s: %d elements, rows %d..%d (%d distinct), max rel %.1e" % (
                        name, d.shape[0], int(d[:, 0].min()), int(d[:, 0].max()), len(set(d[:, 0].tolist())),
                        float((x - y).abs().max() / y.abs().max())))
            print("rep %d: gen2_bwd outputs deviate (g identical: %s): %s" % (r, same_g, "; ".join(msg)))
    if split:
        gt = gt[0]
    if ref is None:
        ref = (gt.clone(), gl.clone(), ga.clone())
        continue
    if not torch.equal(gt, ref[0]):
        bad += 1
        d = (gt != ref[0]).nonzero()
        rows, cols = d[:, 0], d[:, 1]
        o = cols // (4 * I); q = (cols // I) % 4; i = cols % I
        b = rows // K; k = rows % K
        rel = float((gt - ref[0]).abs().max() / ref[0].abs().max())
        print("rep %d: %d elements differ, max rel %.2e; samples %s capsules %s; k range %d..%d (%d distinct), i range %d..%d (%d distinct), q %s" % (
            r, d.shape[0], rel, sorted(set(b.tolist())), sorted(set(o.tolist())), int(k.min()), int(k.max()), len(set(k.tolist())),
            int(i.min()), int(i.max()), len(set(i.tolist())), sorted(set(q.tolist()))))
        if bad <= 3:
            kk = sorted(set(k.tolist()))
            print("    k values:", kk[:40], "..." if len(kk) > 40 else "")
    le = float((gl - ref[1]).abs().max() / ref[1].abs().max()); ae = float((ga - ref[2]).abs().max() / ref[2].abs().max())
    if le > 1e-4 or ae > 1e-4:
        print("rep %d: dL/dlrfs rel %.2e, dL/dact rel %.2e" % (r, le, ae))
print("reps %d, deviating forward: %d, deviating dL/dt: %d, deviating gen2_bwd outputs: %d" % (reps, bad_fwd, bad, bad2))
