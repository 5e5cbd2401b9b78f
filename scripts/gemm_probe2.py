"""probe: 3xTF32 split GEMMs through cuBLAS tensor cores -- speed and accuracy vs fp64"""
import torch
dev = torch.device("cuda:0")
M, N, K = 8192, 20480, 41
def tf32(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def split(x):
    hi = tf32(x); lo = tf32(x - hi); return hi, lo
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.manual_seed(0)
xa = torch.randn(M, K, device=dev); wa = torch.randn(N, K, device=dev) * 0.3; gt = torch.randn(M, N, device=dev)
ref = (xa.double() @ wa.double().t())
torch.backends.cuda.matmul.allow_tf32 = False
y32 = xa @ wa.t()
print("fp32 sgemm err", float((y32 - ref).abs().max() / ref.abs().max()), "time", timeit(lambda: xa @ wa.t()))
torch.backends.cuda.matmul.allow_tf32 = True
y1 = xa @ wa.t()
print("1xTF32 err", float((y1 - ref).abs().max() / ref.abs().max()), "time", timeit(lambda: xa @ wa.t()))
xh, xl = split(xa); wh, wl = split(wa)
A = torch.cat([xh, xh, xl], 1); Bm = torch.cat([wh, wl, wh], 1)
y3 = A @ Bm.t()
print("3xTF32 (K=%d) err" % A.shape[1], float((y3 - ref).abs().max() / ref.abs().max()), "time", timeit(lambda: A @ Bm.t()))
pad = torch.zeros(M, 5, device=dev); padw = torch.zeros(N, 5, device=dev)
A2 = torch.cat([xh, xh, xl, pad], 1); B2 = torch.cat([wh, wl, wh, padw], 1)
print("3xTF32 (K=%d)" % A2.shape[1], "time", timeit(lambda: A2 @ B2.t()))
def fwd_full():
    xh, xl = split(xa); wh, wl = split(wa)
    return torch.cat([xh, xh, xl], 1) @ torch.cat([wh, wl, wh], 1).t()
print("3xTF32 incl. splits", timeit(fwd_full))
# backward shapes
gh, gl = split(gt)
refx = gt.double() @ wa.double(); refw = gt.double().t() @ xa.double()
W2 = torch.cat([wh, wl], 1)   # [N, 2K]
def gx():
    t = gh @ W2
    return t[:, :K] + t[:, K:] + gl @ wh
def gw():
    X2 = torch.cat([xh, xl], 1)
    t = gh.t() @ X2
    return t[:, :K] + t[:, K:] + gl.t() @ xh
print("g_x 3xTF32 err", float((gx() - refx).abs().max() / refx.abs().max()), "time", timeit(gx))
print("g_w 3xTF32 err", float((gw() - refw).abs().max() / refw.abs().max()), "time", timeit(gw))
print("split of g_t", timeit(lambda: split(gt)))
torch.backends.cuda.matmul.allow_tf32 = False
print("fp32 g_x err", float(((gt @ wa) - refx).abs().max() / refx.abs().max()), "g_w err", float(((gt.t() @ xa) - refw).abs().max() / refw.abs().max()))
torch.backends.cuda.matmul.allow_tf32 = True
print("1xTF32 g_x err", float(((gt @ wa) - refx).abs().max() / refx.abs().max()), "time", timeit(lambda: gt @ wa), "g_w err", float(((gt.t() @ xa) - refw).abs().max() / refw.abs().max()), "time", timeit(lambda: gt.t() @ xa))
