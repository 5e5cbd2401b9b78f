"""phase timeline of the packed fused forward (library built with QEC_NVCC_EXTRA=-DQEC_PROF)"""
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
for _ in range(2):
    fwd("qec_routing_fused_fwd", t, lrfs, act, BP, K, I, O, T)
buf = np.zeros(8 * 64 * 32, dtype=np.int64)
lib = _lib.load()
lib.qec_prof_read.argtypes = [ctypes.c_void_p]
rc = lib.qec_prof_read(buf.ctypes.data)
ev = buf.reshape(8, 64, 32)
names = {0: "sweep0 start", 1: "sweep0 end", 2: "r0: cta barrier", 3: "r0: pushed", 4: "r0: totals", 5: "solve0 done", 6: "bcast0 done",
         7: "sweep1 end", 8: "r1: cta barrier", 9: "r1: pushed", 10: "r1: totals", 11: "solve1 done", 12: "bcast1 done",
         14: "sweep2 end", 15: "r2: cta barrier", 16: "r2: pushed", 17: "r2: totals", 18: "solve2 done", 19: "bcast2 done", 21: "epilogue end"}
names.update({22: "r0: summed16", 23: "r0: butterfly", 24: "r0: mbar done"})
order = [0, 1, 2, 22, 23, 3, 24, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 21]
for blk in (0, 3):
    print("== CTA", blk, "(mean over capsules 2..30, cycles since sweep0 start; delta from previous event)")
    d = ev[blk, 2:30, :].astype(np.float64)
    base = d[:, 0:1]
    rel = d - base
    prev = 0.0
    for i in order:
        m = rel[:, i].mean()
        print("  %-16s %8.0f  (+%6.0f)  min %6.0f max %6.0f" % (names[i], m, m - prev, rel[:, i].min(), rel[:, i].max()))
        prev = m
    nxt = ev[blk, 3:31, 0] - ev[blk, 2:30, 0]
    print("  capsule period   %8.0f" % nxt.mean())
