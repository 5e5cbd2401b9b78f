"""phase timeline of the third-generation fused forward (library built with -DQEC_PROF, scripts/build_variant3.sh)"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from qecnetworks_b200 import _lib
if os.environ.get("QEC_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(__file__), "xcheck_fused.py")).read().split("worst = 0.0")[0])
BP, K, I, O, T = 32, 256, 128, 40, 3
t, lrfs, act = inputs(BP, K, I, O, zero_frac=0.0)
os.environ["QEC_PP_VERBOSE"] = "1"
for _ in range(2):
    fwd("qec_routing_fused_pp_fwd", t, lrfs, act, BP, K, I, O, T)
buf = np.zeros(8 * 48 * 96, dtype=np.int64)
lib = _lib.load()
lib.qec_prof3_read.argtypes = [ctypes.c_void_p]
lib.qec_prof3_read(buf.ctypes.data)
ev = buf.reshape(8, 48, 96).astype(np.float64)
names = {0: "pair start"}
for s in (0, 1):
    names[10 + s * 4 + 1] = "sweep0(s%d) end" % s
    names[10 + s * 4 + 2] = "posted0(s%d)" % s
    for tt in (1, 2):
        names[10 + tt * 8 + s * 4 + 3] = "t%d s%d: at pose barrier" % (tt, s)
        names[10 + tt * 8 + s * 4 + 0] = "t%d s%d: pose received" % (tt, s)
        names[10 + tt * 8 + s * 4 + 1] = "t%d s%d: sweep end" % (tt, s)
        names[10 + tt * 8 + s * 4 + 2] = "t%d s%d: posted" % (tt, s)
    names[40 + s * 3 + 2] = "epi s%d: at pose barrier" % s
    names[40 + s * 3] = "epi s%d: pose received" % s
    names[40 + s * 3 + 1] = "epi s%d: end" % s
    for tt in range(3):
        b = 50 + s * 16 + tt * 4
        names[b] = "  solver s%d t%d: partials in" % (s, tt)
        names[b + 1] = "  solver s%d t%d: pushed" % (s, tt)
        names[b + 2] = "  solver s%d t%d: totals in" % (s, tt)
        names[b + 3] = "  solver s%d t%d: pose out" % (s, tt)
for blk in (0, 5):
    d = ev[blk, 2:30, :]
    base = d[:, 0:1]
    rel = d - base
    items = sorted(((rel[:, i].mean(), i) for i in names if d[:, i].min() > 0), key=lambda x: x[0])
    print("== CTA", blk, "(mean over capsule pairs 2..30, cycles since pair start)")
    for m, i in items:
        print("  %-32s %8.0f   min %7.0f max %7.0f" % (names[i], m, rel[:, i].min(), rel[:, i].max()))
    print("  pair period %8.0f" % (ev[blk, 3:31, 0] - ev[blk, 2:30, 0]).mean())
