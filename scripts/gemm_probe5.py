"""probe: alternative formulations of the two skinny backward GEMMs in cuBLAS FP32"""
import torch
dev = torch.device("cuda:0")
M, N, K = 8192, 20480, 41
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
xa = torch.randn(M, K, device=dev); wa = torch.randn(N, K, device=dev); gt = torch.randn(M, N, device=dev)
xat = xa.t().contiguous(); wat = wa.t().contiguous()
print("g_w = gt.T @ xa        ", timeit(lambda: gt.t() @ xa))
print("g_w^T = xa.T @ gt      ", timeit(lambda: xa.t() @ gt))
print("g_w^T = xat @ gt       ", timeit(lambda: xat @ gt))
print("g_x = gt @ wa          ", timeit(lambda: gt @ wa))
print("g_x^T = wa.T @ gt.T    ", timeit(lambda: wa.t() @ gt.t()))
print("g_x^T = wat @ gt.T     ", timeit(lambda: wat @ gt.t()))
# batched split-K by hand: view gt as [8, M/8, N]
g3 = gt.view(8, M // 8, N); x3 = xa.view(8, M // 8, K)
print("g_w split-R bmm        ", timeit(lambda: torch.bmm(g3.transpose(1, 2), x3).sum(0)))
g3n = gt.view(M, 8, N // 8).transpose(0, 1); w3 = wa.view(8, N // 8, K)
print("g_x split-N bmm        ", timeit(lambda: torch.bmm(g3n, w3).sum(0)))
