"""one forward+backward of the pool2 layer through the drop-in QecModule (BASELINE config 2: B=32, 1 patch,
K=256, I=128, O=40, T=3) -- driver for ncu captures of the path the training step takes (pre-split gradient)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qecnetworks_b200 as qb
B, P, K, I, O, T = 32, 1, 256, 128, 40, 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = qb.QecModule(I, O, T, K, P).to(dev)
g = torch.Generator().manual_seed(0)
base = torch.randn(B, 1, 1, 1, 4, generator=g)
lrfs = base + 0.3 * torch.randn(B, P, K, I, 4, generator=g)
lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
lrfs = (lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)).to(dev).requires_grad_(True)
act = (torch.rand(B, P, K, I, generator=g) * 0.5 + 0.4).to(dev).requires_grad_(True)
pts = torch.randn(B, P, K, 3, generator=g).to(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    pose, agr = m(pts, lrfs, act)
    (pose.sum() + agr.sum()).backward()
torch.cuda.synchronize()
print("ok", float(agr.detach().mean()))
