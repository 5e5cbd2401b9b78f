"""FFMA2 / FFMA / FMNMX issue-rate probe (qec_pipe_probe): warp-instructions per cycle per SM sub-partition"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import _lib
dev = torch.device("cuda:0")
out = torch.zeros(1 << 20, device=dev)
blocks, iters = 148 * 16, 2048
st = torch.cuda.current_stream().cuda_stream
names = {0: "16 FFMA", 1: "16 FFMA2", 2: "8 FFMA + 8 FMNMX", 3: "8 FFMA2 + 8 FMNMX", 4: "8 FFMA2 + 8 FFMA",
         5: "8 FFMA2", 6: "8 FFMA", 7: "16 FMNMX", 8: "16 FFMA + 8 FMNMX", 9: "8 FMNMX"}
nfp = {0: (0, 16), 1: (16, 0), 2: (0, 8), 3: (8, 0), 4: (8, 8), 5: (8, 0), 6: (0, 8), 7: (0, 0), 8: (0, 16), 9: (0, 0)}
for mode in range(10):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("qec_pipe_probe", out.data_ptr(), blocks, iters, mode, st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    f2, f1 = nfp[mode]
    flop = (4.0 * f2 + 2.0 * f1) * iters * 256 * blocks
    ninst = {0: 16, 1: 16, 2: 16, 3: 16, 4: 16, 5: 8, 6: 8, 7: 16, 8: 24, 9: 8}[mode]
    cyc = best * 1e-3 * 1.965e9 / (blocks * 8 / (148 * 4)) / iters  # cycles per round per SMSP-resident warp set
    print("mode %d %-22s %8.3f ms  %7.2f TFLOP/s  %5.2f issue-cycles/round (@1965 MHz) for %d instr" % (mode, names[mode], best, flop / best / 1e9, cyc, ninst))
