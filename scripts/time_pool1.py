"""time the pool1-shaped register-resident kernels in isolation (CUDA events)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qecnetworks_b200 as qb
from qecnetworks_b200 import _lib
from oracle import qec_oracle as orc
if os.environ.get("QEC_LIB"):  # a timing variant
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = qb.QecModule(1, 128, 3, 9, 256).to(dev)
points, lrfs, act, labels = orc.synthetic_batch(32, ncls=40, seed=1)
points, lrfs, act = points.to(dev), lrfs.to(dev).unsqueeze(-2), act.to(dev).unsqueeze(-1)
for _ in range(3):
    pose, agr = m(points, lrfs, act); (pose.sum() + agr.sum()).backward()
torch.cuda.synchronize()
_lib.start_kernel_timing()
for _ in range(10):
    pose, agr = m(points, lrfs, act); (pose.sum() + agr.sum()).backward()
torch.cuda.synchronize()
for (name, dims), ts in _lib.stop_kernel_timing().items():
    print("%s %.3f ms (min %.3f)" % (name, sum(ts) / len(ts), min(ts)))
