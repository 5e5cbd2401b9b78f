"""warm, back-to-back timing of the small kernels of the config-2 training step (the per-kernel events of bench.py
include the host's launch gap for kernels of a few microseconds; here 50 launches sit between two events)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF, _lib

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, P1, K1, C, NC = 32, 256, 9, 128, 40


def timeit(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-34s %8.1f us" % (name, e0.elapsed_time(e1) * 1000 / n))


def rq(*shape):
    q = torch.randn(*shape, 4, generator=g)
    return (q / q.norm(dim=-1, keepdim=True)).to(dev)


# pool1-side
W1 = torch.randn(C, 3, generator=g).to(dev); b1 = torch.randn(C, generator=g).to(dev)
W2 = torch.randn(4 * C, C, generator=g).to(dev); b2 = torch.randn(4 * C, generator=g).to(dev)
Wc = torch.empty(4 * C, 3, device=dev); bc = torch.empty(4 * C, device=dev)
gW1, gb1, gW2 = torch.empty_like(W1), torch.empty_like(b1), torch.empty_like(W2)
timeit("compose_fwd", lambda: _lib.launch("qec_compose_fwd", W1, b1, W2, b2, Wc, bc, 4 * C, C, 3))
timeit("compose_bwd", lambda: _lib.launch("qec_compose_bwd", W1, b1, W2, Wc, bc, gW1, gb1, gW2, None, 4 * C, C, 3))
# pool2-side: quater_gen1 = Linear(3*128, 40) over 8192 rows
M, Cin = B * 256, 3 * C
x = torch.randn(M, Cin, generator=g).to(dev); W = torch.randn(NC, Cin, generator=g).to(dev); b = torch.randn(NC, generator=g).to(dev)
y = torch.empty(M, NC, device=dev); gy = torch.randn(M, NC, generator=g).to(dev)
gx, gW, gb = torch.empty_like(x), torch.empty_like(W), torch.empty_like(b)
ws = torch.empty(int(_lib.query("qec_linear_bwd_workspace", M, Cin, NC)), device=dev)
timeit("linear_fwd  [8192x384]->40", lambda: _lib.launch("qec_linear_fwd", x, W, b, y, M, Cin, NC))
timeit("linear_bwd (+fold)", lambda: _lib.launch("qec_linear_bwd", x, W, gy, gx, gW, gb, ws, M, Cin, NC))
timeit("cublas addmm fwd", lambda: torch.addmm(b, x, W.t()))
timeit("cublas bwd (2 mm + sum)", lambda: (torch.mm(gy, W), torch.mm(gy.t(), x), gy.sum(0)))
# mean LRF of pool2: lrfs [32,1,256,128,4]
lr2 = rq(B, 1, 256, C); act2 = torch.rand(B, 1, 256, C, generator=g).to(dev); pts2 = torch.randn(B, 1, 256, 3, generator=g).to(dev)
mean = torch.empty(B, 1, C, 4, device=dev); xrot = torch.empty(B, 1, 256, C, 3, device=dev)
timeit("lrf_mean_fwd pool2 (mean+rotate)", lambda: _lib.launch("qec_lrf_mean_fwd", lr2, act2, pts2, mean, xrot, B, 256, C))
gxr = torch.randn(B, 1, 256, C, 3, generator=g).to(dev); glr = torch.empty_like(lr2)
timeit("lrf_mean_bwd pool2", lambda: _lib.launch("qec_lrf_mean_bwd", lr2, pts2, mean, gxr, None, glr, None, B, 256, C))
lr1 = rq(B, P1, K1, 1); act1 = torch.rand(B, P1, K1, 1, generator=g).to(dev); pts1 = torch.randn(B, P1, K1, 3, generator=g).to(dev)
mean1 = torch.empty(B, P1, 1, 4, device=dev); xrot1 = torch.empty(B, P1, K1, 1, 3, device=dev)
timeit("lrf_mean_fwd pool1", lambda: _lib.launch("qec_lrf_mean_fwd", lr1, act1, pts1, mean1, xrot1, B * P1, K1, 1))
