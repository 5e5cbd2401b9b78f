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
refw = gt.double().t() @ xa.double(); refx = gt.double() @ wa.double()
for S in (2, 4, 8, 16, 32, 64):
    g3 = gt.view(S, M // S, N); x3 = xa.view(S, M // S, K)
    fw = lambda: torch.bmm(g3.transpose(1, 2), x3).sum(0)
    g3n = gt.view(M, S, N // S).transpose(0, 1); w3 = wa.view(S, N // S, K)
    fx = lambda: torch.bmm(g3n, w3).sum(0)
    print("S", S, "g_w", round(timeit(fw)), "err", float((fw() - refw).abs().max() / refw.abs().max()),
          "| g_x", round(timeit(fx)), "err", float((fx() - refx).abs().max() / refx.abs().max()))
