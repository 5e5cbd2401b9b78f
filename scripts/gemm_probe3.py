"""probe: backward GEMMs of quater_gen2 with pre-split gradient (g_hi, g_lo) and padded small dims"""
import torch
dev = torch.device("cuda:0")
M, N, K = 8192, 20480, 41
def tf32(x): return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def split(x):
    hi = tf32(x); return hi, tf32(x - hi)
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
refx = gt.double() @ wa.double(); refw = gt.double().t() @ xa.double()
gh, gl = split(gt)
torch.backends.cuda.matmul.allow_tf32 = True
for KP in (48, 64):
    wh, wl = split(wa); xh, xl = split(xa)
    z = torch.zeros(N, KP - K, device=dev); zx = torch.zeros(M, KP - K, device=dev)
    Wcat = torch.cat([wh, z, wl, z], 1).contiguous()   # [N, 2KP]
    Whp = torch.cat([wh, z], 1).contiguous()           # [N, KP]
    Xcat = torch.cat([xh, zx, xl, zx], 1).contiguous()
    Xhp = torch.cat([xh, zx], 1).contiguous()
    def gx():
        t = gh @ Wcat
        return t[:, :K] + t[:, KP:KP + K] + (gl @ Whp)[:, :K]
    def gw():
        t = gh.t() @ Xcat
        return t[:, :K] + t[:, KP:KP + K] + (gl.t() @ Xhp)[:, :K]
    print("KP", KP, "g_x err", float((gx() - refx).abs().max() / refx.abs().max()), "time", timeit(gx),
          "| g_w err", float((gw() - refw).abs().max() / refw.abs().max()), "time", timeit(gw))
    print("   parts: gh@Wcat", timeit(lambda: gh @ Wcat), "gl@Whp", timeit(lambda: gl @ Whp), "gh.t@Xcat", timeit(lambda: gh.t() @ Xcat), "gl.t@Xhp", timeit(lambda: gl.t() @ Xhp))
# alternative: stack hi and lo along the contraction axis: [gh | gl] (M x 2N) @ [[Wh+Wl],[Wh]] -> one pass each
G2 = torch.cat([gh, gl], 1)            # [M, 2N]  (what the kernel could write directly)
wh, wl = split(wa)
for KP in (48, 64):
    z = torch.zeros(N, KP - K, device=dev)
    # g_x = gh@wh + gh@wl + gl@wh : contraction over 2N with B = [[wh, wl],[wh, 0]] and sum of the two column blocks
    Bx = torch.cat([torch.cat([wh, z, wl, z], 1), torch.cat([wh, z, torch.zeros(N, KP, device=dev)], 1)], 0).contiguous()  # [2N, 2KP]
    def gx2():
        t = G2 @ Bx
        return t[:, :K] + t[:, KP:KP + K]
    print("KP", KP, "stacked g_x err", float((gx2() - refx).abs().max() / refx.abs().max()), "time", timeit(gx2))
