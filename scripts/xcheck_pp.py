"""third-generation cluster kernels (routing_fused3.cu: two capsules in flight, solver warp) against the second
generation (routing_fused2.cu) on the same inputs, forward (and backward when given `bwd`), plus their timing at
BASELINE config 2's pool2 geometry"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import _lib, functional as QF
if os.environ.get("QEC_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])
argv = sys.argv[1:]
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(__file__), "xcheck_fused.py")).read().split("worst = 0.0")[0])

shapes = [(2, 16, 128, 6, 3), (3, 32, 128, 5, 3), (2, 64, 128, 4, 2), (1, 256, 128, 6, 3), (32, 256, 128, 40, 3),
          (3, 100, 30, 4, 3), (2, 37, 6, 5, 1), (1, 250, 100, 3, 3), (2, 255, 126, 3, 3), (1, 1, 2, 1, 3)]
worst = 0.0
for (BP, K, I, O, T) in shapes:
    if not _lib.query("qec_routing_fused_pp_supported", K, I, O, T):
        print("skip", (BP, K, I, O, T)); continue
    t, lrfs, act = inputs(BP, K, I, O)
    QF.set_fused_pp(0)
    ok2 = _lib.query("qec_routing_fused_packed_supported", K, I, O, T)
    a = fwd("qec_routing_fused_fwd" if ok2 else "qec_routing_fused_fwd_scalar", t, lrfs, act, BP, K, I, O, T)
    b = fwd("qec_routing_fused_pp_fwd", t, lrfs, act, BP, K, I, O, T)
    b2 = fwd("qec_routing_fused_pp_fwd", t, lrfs, act, BP, K, I, O, T)
    errs = [rel(x, y) for x, y in zip(b, a)]
    same = all(torch.equal(x, y) for x, y in zip(b, b2))
    worst = max(worst, *errs)
    print((BP, K, I, O, T), "pose %.2e agr %.2e poses_all %.2e stats %.2e  reproducible %s" % (*errs, same))
print("worst", worst)
if "time" in argv:
    BP, K, I, O, T = 32, 256, 128, 40, 3
    t, lrfs, act = inputs(BP, K, I, O, zero_frac=0.0)
    for name in ("qec_routing_fused_fwd", "qec_routing_fused_pp_fwd"):
        for _ in range(3):
            fwd(name, t, lrfs, act, BP, K, I, O, T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        alpha = torch.ones(1, device=dev); beta = torch.ones(1, device=dev)
        Nc = BP * O
        pose = torch.empty(Nc, 4, device=dev); agr = torch.empty(Nc, device=dev)
        pa = torch.empty(T, Nc, 4, device=dev); st = torch.empty(Nc, 2, device=dev)
        e0.record()
        for _ in range(20):
            _lib.call(name, t.data_ptr(), lrfs.data_ptr(), act.data_ptr(), alpha.data_ptr(), beta.data_ptr(), pose.data_ptr(),
                      agr.data_ptr(), pa.data_ptr(), st.data_ptr(), BP, K, I, O, T, torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize()
        print("%-28s %.4f ms" % (name, e0.elapsed_time(e1) / 20))
