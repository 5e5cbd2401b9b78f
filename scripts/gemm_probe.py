"""probe: cuBLAS FP32 throughput of the quater_gen2 GEMMs (fwd + both bwd) for padded K"""
import torch, sys
dev = torch.device("cuda:0")
M, N = 8192, 20480
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for K in (40, 41, 44, 48, 64):
    xa = torch.randn(M, K, device=dev); wa = torch.randn(N, K, device=dev); gt = torch.randn(M, N, device=dev)
    t1 = timeit(lambda: xa @ wa.t())
    t2 = timeit(lambda: gt @ wa)          # g_xa
    t3 = timeit(lambda: gt.t() @ xa)      # g_wa
    fl = 2.0 * M * N * K / 1e9
    print(f"K={K}: fwd {t1*1e3:.0f} us ({fl/t1:.1f} TF/s)  g_x {t2*1e3:.0f} us ({fl/t2:.1f})  g_w {t3*1e3:.0f} us ({fl/t3:.1f})")
# one-pass alternative for the two backward products: [g_xa ; ...] no -- baseline numbers only
