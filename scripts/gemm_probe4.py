"""probe: accuracy of 3xTF32 backward GEMMs when the long contraction is chunked (partial sums in FP32)"""
import torch
dev = torch.device("cuda:0")
M, N, K = 8192, 20480, 41
KP = 48
def tf32(x): return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def split(x):
    hi = tf32(x); return hi, tf32(x - hi)
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.manual_seed(0)
xa = torch.randn(M, K, device=dev); wa = torch.randn(N, K, device=dev) * 0.3; gt = torch.randn(M, N, device=dev)
refx = gt.double() @ wa.double(); refw = gt.double().t() @ xa.double()
gh, gl = split(gt)
wh, wl = split(wa); xh, xl = split(xa)
z = torch.zeros(N, KP - K, device=dev); zx = torch.zeros(M, KP - K, device=dev)
Wcat = torch.cat([wh, z, wl, z], 1).contiguous(); Whp = torch.cat([wh, z], 1).contiguous()
Xcat = torch.cat([xh, zx, xl, zx], 1).contiguous(); Xhp = torch.cat([xh, zx], 1).contiguous()
torch.backends.cuda.matmul.allow_tf32 = True
for nchunk in (1, 4, 16):
    cs = N // nchunk
    def gx():
        acc = None
        for c in range(nchunk):
            sl = slice(c * cs, (c + 1) * cs)
            t = gh[:, sl] @ Wcat[sl]
            p = t[:, :K] + t[:, KP:KP + K] + (gl[:, sl] @ Whp[sl])[:, :K]
            acc = p if acc is None else acc + p
        return acc
    rs = M // nchunk
    def gw():
        acc = None
        for c in range(nchunk):
            sl = slice(c * rs, (c + 1) * rs)
            t = gh[sl].t() @ Xcat[sl]
            p = t[:, :K] + t[:, KP:KP + K] + (gl[sl].t() @ Xhp[sl])[:, :K]
            acc = p if acc is None else acc + p
        return acc
    print("chunks", nchunk, "g_x err", float((gx() - refx).abs().max() / refx.abs().max()), "time", timeit(gx),
          "| g_w err", float((gw() - refw).abs().max() / refw.abs().max()), "time", timeit(gw))
# is the error from the hi*hi product alone (tensor-core accumulation) ?
t = (gh.double() @ wh.double()); th = gh @ Whp
print("hi*hi only: TC vs exact", float((th[:, :K] - t).abs().max() / t.abs().max()))
torch.backends.cuda.matmul.allow_tf32 = False
print("fp32 sgemm: g_x", float(((gt @ wa) - refx).abs().max() / refx.abs().max()), "g_w", float(((gt.t() @ xa) - refw).abs().max() / refw.abs().max()))
