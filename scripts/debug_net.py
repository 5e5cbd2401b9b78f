"""debug helper: per-parameter gradient errors of the CUDA net vs golden / oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import qec_oracle as orc
from tests.util import load_golden, params_of, quat_err, max_err, rel_err
import qecnetworks_b200 as qb

dev = torch.device("cuda:0")
g = load_golden("net_small")
B, C, ncls, T = [int(v) for v in g["dims"]]
net = qb.QecNet(2048, C, ncls, T); net.load_state_dict(params_of(g)); net = net.to(dev)
pose, a = net(g["points"].to(dev), g["lrfs"].to(dev), g["act"].to(dev), None)
loss = net.spread_loss(a, g["labels"].to(dev), 0); loss.backward()
p64 = {k: v.double().requires_grad_(True) for k, v in params_of(g).items()}
pr, ar = orc.qec_net_forward(p64, g["points"].double(), g["lrfs"].double(), g["act"].double(), T)
orc.spread_loss(ar, g["labels"]).backward()
print("pose err golden", quat_err(pose.cpu(), g["pose"]), "oracle64", quat_err(pose.cpu(), pr))
print("agr err golden", max_err(a.cpu(), g["agreement"]), "oracle64", max_err(a.cpu(), ar))
for k, v in net.named_parameters():
    ref = g["grad." + k]; r64 = p64[k].grad
    print(f"{k:28s} |ref|max {float(ref.abs().max()):.3e}  cuda-vs-golden {max_err(v.grad.cpu(), ref):.3e}  cuda-vs-f64 {max_err(v.grad.cpu(), r64):.3e}  golden-vs-f64 {max_err(ref, r64):.3e}")
