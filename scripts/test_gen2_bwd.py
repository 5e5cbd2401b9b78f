"""bring-up / accuracy / timing of qec_gen2_bwd (tcgen05 tf32 x3 backward of quater_gen2) against float64 matmuls"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import _lib

dev = torch.device("cuda:0")
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
shapes_arg = sys.argv[2] if len(sys.argv) > 2 else "all"
shapes = [(128, 128, 40), (256, 384, 40), (512, 20480, 40), (8192, 20480, 40)]
for (M, N, Kc) in shapes:
    gen = torch.Generator().manual_seed(M + N)
    g = torch.randn(M, N, generator=gen).to(dev)
    h = torch.randn(M, Kc, generator=gen).to(dev)
    W2 = (torch.randn(N, Kc, generator=gen) / 6).to(dev)
    perm = torch.randperm(N, generator=gen).to(dev)
    ref_h = (g.double() @ W2.double()[perm]).float()
    haug = torch.cat([h.double(), torch.ones(M, 1, device=dev, dtype=torch.float64)], 1)
    ref_w = g.double().t() @ haug                      # [N, Kc+1] in column order of g
    ws = torch.full((_lib.query("qec_gen2_bwd_workspace", M, N, Kc),), 3.0, device=dev)
    for v in variants:
        g_h = torch.full((M, Kc), 7.0, device=dev); g_W2 = torch.full((N, Kc), 7.0, device=dev); g_b2 = torch.full((N,), 7.0, device=dev)
        _lib.launch("qec_gen2_bwd", g, h, W2, perm, g_h, g_W2, g_b2, ws, M, N, Kc, v)
        torch.cuda.synchronize()
        eh = float((g_h - ref_h).abs().max() / ref_h.abs().max())
        gw = torch.empty(N, Kc + 1, device=dev, dtype=torch.float64)
        gw[:, :Kc] = g_W2[perm].double(); gw[:, Kc] = g_b2[perm].double()
        ew = float((gw - ref_w).abs().max() / ref_w.abs().max())
        print("M=%d N=%d variant %d: rel err dh %.2e  dW %.2e   dW[0,:4] got %s ref %s" % (M, N, v, eh, ew, gw[0, :4].tolist(), ref_w[0, :4].tolist()))
    if M == 8192:
        for _ in range(3):
            _lib.launch("qec_gen2_bwd", g, h, W2, perm, g_h, g_W2, g_b2, ws, M, N, Kc, variants[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.launch("qec_gen2_bwd", g, h, W2, perm, g_h, g_W2, g_b2, ws, M, N, Kc, variants[0])
        e1.record(); torch.cuda.synchronize()
        print("qec_gen2_bwd: %.1f us per call (prep + dh + dW + fold)" % (e0.elapsed_time(e1) * 100))
        e0.record()
        for _ in range(10):
            _lib.launch("qec_gen2_bwd", g, h, W2, perm, g_h, None, None, ws, M, N, Kc, variants[0])
        e1.record(); torch.cuda.synchronize()
        print("   dh only: %.1f us" % (e0.elapsed_time(e1) * 100))
        e0.record()
        for _ in range(10):
            _lib.launch("qec_gen2_bwd", g, h, W2, perm, None, g_W2, g_b2, ws, M, N, Kc, variants[0])
        e1.record(); torch.cuda.synchronize()
        print("   dW only: %.1f us" % (e0.elapsed_time(e1) * 100))
        for probe, name in ((16, "no MMA"), (48, "no MMA, no split")):
            e0.record()
            for _ in range(10):
                _lib.launch("qec_gen2_bwd", g, h, W2, perm, g_h, None, None, ws, M, N, Kc, probe)
            e1.record(); torch.cuda.synchronize()
            print("   dh only, %s: %.1f us" % (name, e0.elapsed_time(e1) * 100))
            e0.record()
            for _ in range(10):
                _lib.launch("qec_gen2_bwd", g, h, W2, perm, None, g_W2, g_b2, ws, M, N, Kc, probe)
            e1.record(); torch.cuda.synchronize()
            print("   dW only, %s: %.1f us" % (name, e0.elapsed_time(e1) * 100))
