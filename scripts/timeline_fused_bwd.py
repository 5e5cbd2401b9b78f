"""phase timeline of the packed fused backward (library built with QEC_NVCC_EXTRA=-DQEC_PROF)"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from qecnetworks_b200 import _lib
if os.environ.get("QEC_LIB"):  # a variant built by scripts/build_variant.sh
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(__file__), "xcheck_fused.py")).read().split("worst = 0.0")[0])
BP, K, I, O, T = 32, 256, 128, 40, 3
t, lrfs, act = inputs(BP, K, I, O)
f = fwd("qec_routing_fused_fwd", t, lrfs, act, BP, K, I, O, T)
for _ in range(2):
    bwd("qec_routing_fused_bwd", t, lrfs, act, BP, K, I, O, T, f)
buf = np.zeros(8 * 64 * 32, dtype=np.int64)
lib = _lib.load()
lib.qec_prof_read.argtypes = [ctypes.c_void_p]
lib.qec_prof_read(buf.ctypes.data)
ev = buf.reshape(8, 64, 32)
names = {0: "sweep0 start", 1: "sweep0 end", 2: "r0: cta barrier", 3: "r0: pushed", 4: "r0: totals", 5: "solve0 done", 6: "sync0 done",
         7: "grad pass end", 8: "acc(2) end", 9: "r1: cta barrier", 10: "r1: pushed", 11: "r1: totals", 12: "solve1 done", 13: "sync1 done",
         15: "acc(1) end", 16: "r2: cta barrier", 17: "r2: pushed", 18: "r2: totals", 19: "solve2 done", 20: "sync2 done", 25: "acc(0) end", 22: "r0: summed16", 23: "r0: butterfly", 24: "r0: mbar done"}
order = [0, 1, 2, 22, 23, 3, 24, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 18, 19, 20, 25]
for blk in (0,):
    d = ev[blk, 2:30, :].astype(np.float64)
    rel = d - d[:, 0:1]
    prev = 0.0
    for i in order:
        m = rel[:, i].mean()
        print("  %-16s %8.0f  (+%6.0f)" % (names[i], m, m - prev))
        prev = m
    print("  capsule period   %8.0f" % (ev[blk, 3:31, 0] - ev[blk, 2:30, 0]).mean())
