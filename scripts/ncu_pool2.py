"""workload for the ncu capture of the two packed cluster kernels at BASELINE config 2's pool2 geometry:
three rounds of (fused forward, pre-split fused backward); profile the third,
    ncu --set full --clock-control none --import-source on -k regex:routing_fused2 -s 4 -c 2 -o OUT python scripts/ncu_pool2.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF, _lib
if os.environ.get("QEC_LIB"):  # a variant built by scripts/build_variant.sh
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
h = torch.randn(B * P * K, O, generator=g).to(dev)
W2 = (torch.randn(4 * I * O, O, generator=g) / 6).to(dev)
perm = torch.randperm(4 * I * O, generator=g).to(dev)
with torch.no_grad():
    for _ in range(3):  # the training step's pool2 kernels: forward, plain-FP32 backward, tcgen05 backward of quater_gen2
        pose, agr, poses_all, stats = QF.routing_fused_forward_raw(t_oqi, lrfs, act, alpha, beta, T)
        g_t, _, _, _ = QF.routing_fused_backward_raw(t_oqi, lrfs, act, alpha, beta, poses_all, stats, torch.ones_like(pose),
                                                     torch.ones_like(agr), T, True, True, split=False)
        QF.gen2_backward(g_t, h, W2, perm)
torch.cuda.synchronize()
print("ok")
