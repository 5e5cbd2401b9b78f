"""time the pool2-shaped fused routing forward / backward kernels in isolation (CUDA events)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF, _lib
if os.environ.get("QEC_LIB"):  # a timing variant built by scripts/build_variant.sh
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])

B, P, K, I, O, T = 32, 1, 256, 128, 40, 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
base = torch.randn(B, 1, 1, 1, 4, generator=g)
lrfs = base + 0.3 * torch.randn(B, P, K, I, 4, generator=g)
lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
lrfs = (lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)).to(dev).requires_grad_(True)
act = (torch.rand(B, P, K * I, generator=g) * 0.5 + 0.4).to(dev).requires_grad_(True)
tb = torch.randn(1, O, 4, 1, generator=g)
t_oqi = (tb + 0.25 * torch.randn(B * P * K, O, 4, I, generator=g)).reshape(B * P * K, 4 * I * O).to(dev).requires_grad_(True)
alpha = torch.ones(1, device=dev, requires_grad=True)
beta = torch.ones(1, device=dev, requires_grad=True)
for _ in range(3):
    pose, agr = QF.routing_fused(t_oqi, lrfs, act, alpha, beta, T)
    (pose.sum() + agr.sum()).backward()
torch.cuda.synchronize()
_lib.start_kernel_timing()
for _ in range(10):
    pose, agr = QF.routing_fused(t_oqi, lrfs, act, alpha, beta, T)
    (pose.sum() + agr.sum()).backward()
torch.cuda.synchronize()
for (name, dims), ts in _lib.stop_kernel_timing().items():
    print("%s %.3f ms (min %.3f)" % (name, sum(ts) / len(ts), min(ts)))

# the same backward through the raw entry points, plain and pre-split (what the one-node GEMM + routing path calls)
with torch.no_grad():
    pose, agr, poses_all, stats = QF.routing_fused_forward_raw(t_oqi, lrfs, act, alpha, beta, T)
    gp, ga = torch.ones_like(pose), torch.ones_like(agr)
    for split in (False, True):
        for _ in range(3):
            QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, True, True, split=split)
        torch.cuda.synchronize()
        _lib.start_kernel_timing()
        for _ in range(10):
            QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, True, True, split=split)
        torch.cuda.synchronize()
        for (name, dims), ts in _lib.stop_kernel_timing().items():
            print("raw %s %.3f ms (min %.3f)" % (name, sum(ts) / len(ts), min(ts)))
    # what the dL/dlrfs atomics (RED.ADD.F32x4 per vote, summed over the O capsules of a patch) cost: the same
    # pre-split backward without that output
    for _ in range(3):
        QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, False, True, split=True)
    torch.cuda.synchronize()
    _lib.start_kernel_timing()
    for _ in range(10):
        QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, gp, ga, T, False, True, split=True)
    torch.cuda.synchronize()
    for (name, dims), ts in _lib.stop_kernel_timing().items():
        print("raw-no-g_lrfs %s %.3f ms (min %.3f)" % (name, sum(ts) / len(ts), min(ts)))
