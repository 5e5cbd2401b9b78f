"""run-to-run variation of the Siamese full-geometry backward (BASELINE config 4 geometry, B = 2 pairs): repeat the same
forward + backward N times and report every repetition whose gradients deviate from the first by more than the noise
of the FP32 atomics (flakiness hunt; usage: python scripts/flaky_hunt.py [reps])"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qecnetworks_b200 as qb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kind = sys.argv[2] if len(sys.argv) > 2 else "sia"      # sia | qec | sia1 (Siamese net, spread loss only)
dev = torch.device("cuda:0")
from oracle import qec_oracle as orc   # (a diagnostic script: the oracle only supplies the test's parameters and inputs)
from qecnetworks_b200.synthetic import siamese_pair

params = orc.init_net_params(128, 40, seed=12)      # exactly tests/test_gpu_module.py::test_sia_net_vs_oracle_full_geometry
net = (qb.QecNet if kind == "qec" else qb.QecSiaNet)(2048, 128, 40, 3)
net.load_state_dict(params)
net = net.to(dev)
pts, lr, ac, lab = orc.synthetic_batch(2, ncls=40, seed=41)
p2, l2, a2 = siamese_pair(pts, lr, ac, seed=42)
p2, l2, a2, lab = p2.to(dev), l2.to(dev), a2.to(dev), lab.to(dev)
tgt = torch.stack([lab, lab], 1)
ref = None
worst = {}
bad = 0
for r in range(reps):
    for p in net.parameters():
        p.grad = None
    if kind == "qec":
        pose, a = net(p2[:, 0], l2[:, 0], a2[:, 0], None)
        net.spread_loss(a, lab).backward()
    elif kind == "sia1":
        pose, a = net(p2, l2, a2, None)
        net.spread_loss(a, tgt).backward()
    else:
        pose, a = net(p2, l2, a2, None)
        total, _, _ = net.class_pose_loss(a, pose, tgt)
        total.backward()
    cur = {k: v.grad.detach().clone() for k, v in net.named_parameters()}
    cur["__pose"] = pose.detach().clone(); cur["__act"] = a.detach().clone()
    if ref is None:
        ref = cur
        scale = max(float(v.abs().max()) for k, v in ref.items() if not k.startswith("__"))
        continue
    errs = {k: float((cur[k] - ref[k]).abs().max()) / (1.0 if k.startswith("__") else max(float(ref[k].abs().max()), 0.25 * scale)) for k in ref}
    m = max(errs.values())
    for k, e in errs.items():
        worst[k] = max(worst.get(k, 0.0), e)
    if m > 3e-5:
        bad += 1
        print("rep %d: " % r + ", ".join("%s %.1e" % (k.replace("quater_", "q"), e) for k, e in errs.items() if e > 3e-5))
print("reps %d, outliers %d; worst per tensor: " % (reps, bad) + ", ".join("%s %.1e" % (k.replace("quater_", "q"), e) for k, e in worst.items()))
