"""one line per library: pool2 cluster kernels of BASELINE config 2 in isolation (forward, plain and pre-split backward),
mean of 20 launches each after 5 warm-ups; QEC_LIB selects a variant library, QEC_FUSED_WIDE the 512-thread build"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF, _lib
if os.environ.get("QEC_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])
B, P, K, I, O, T = 32, 1, 256, 128, 40, 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
base = torch.randn(B, 1, 1, 1, 4, generator=g)
lrfs = base + 0.3 * torch.randn(B, P, K, I, 4, generator=g)
lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
lrfs = (lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)).to(dev)
act = (torch.rand(B, P, K * I, generator=g) * 0.5 + 0.4).to(dev)
tb = torch.randn(1, O, 4, 1, generator=g)
t_oqi = (tb + 0.25 * torch.randn(B * P * K, O, 4, I, generator=g)).reshape(B * P * K, 4 * I * O).to(dev)
alpha = torch.ones(1, device=dev)
beta = torch.ones(1, device=dev)
out = {}
with torch.no_grad():
    def timed(fn, n=20, w=5):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        _lib.start_kernel_timing()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        (key, ts), = _lib.stop_kernel_timing().items()
        return sum(ts) / len(ts)
    res = []
    def f():
        res[:] = QF.routing_fused_forward_raw(t_oqi, lrfs, act, alpha, beta, T)
    out["fwd"] = timed(f)
    pose, agr, poses_all, stats = res
    gp, ga = torch.ones_like(pose), torch.ones_like(agr)
    out["bwd"] = timed(lambda: QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, True, True, split=False))
    out["bwd_split"] = timed(lambda: QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, True, True, split=True))
print("%-28s wide=%s  fwd %.4f  bwd %.4f  bwd_split %.4f ms   checksum %.6f" % (
    os.path.basename(os.environ.get("QEC_LIB", "in-tree")), os.environ.get("QEC_FUSED_WIDE", "0"), out["fwd"], out["bwd"],
    out["bwd_split"], float(pose.abs().sum())))
