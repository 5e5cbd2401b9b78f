"""probe: the two skinny backward GEMMs of quater_gen2 on the TF32 tensor cores at FP32 accuracy, with the
gradient delivered pre-split (g = g_hi + g_lo, g_hi on the TF32 grid) and the long contraction cut into S
chunks that are summed in FP32"""
import torch
dev = torch.device("cuda:0")
M, N, K = 8192, 20480, 41
def tf32_round(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.manual_seed(0)
xa = torch.randn(M, K, device=dev); wa = torch.randn(N, K, device=dev) * 0.2
gt = torch.randn(M, N, device=dev) * torch.rand(M, 1, device=dev)
G2 = torch.empty(2, M, N, device=dev)
G2[0] = tf32_round(gt); G2[1] = gt - G2[0]
refw = (gt.double().t() @ xa.double()); refx = (gt.double() @ wa.double())
print("SGEMM g_w %.0f us err %.2e | g_x %.0f us err %.2e" % (
    timeit(lambda: gt.t() @ xa), float(((gt.t() @ xa) - refw).abs().max() / refw.abs().max()),
    timeit(lambda: gt @ wa), float(((gt @ wa) - refx).abs().max() / refx.abs().max())))
torch.backends.cuda.matmul.allow_tf32 = True
xh = tf32_round(xa); xl = tf32_round(xa - xh); wh = tf32_round(wa); wl = tf32_round(wa - wh)
KP = 48  # padded width of one operand block
def pad(t, w): return torch.cat([t, t.new_zeros(t.shape[0], w - t.shape[1])], 1)
X0 = torch.cat([pad(xh, KP), pad(xl, KP)], 1).contiguous(); X1 = pad(xh, KP).contiguous()      # [M, 96], [M, 48]
W0 = torch.cat([pad(wh, KP), pad(wl, KP)], 1).contiguous(); W1 = pad(wh, KP).contiguous()      # [N, 96], [N, 48]
for S in (1, 4, 16, 32, 64):
    R, C = M // S, N // S
    def gw():
        a = torch.bmm(G2[0].view(S, R, N).transpose(1, 2), X0.view(S, R, 2 * KP)).sum(0)
        b = torch.bmm(G2[1].view(S, R, N).transpose(1, 2), X1.view(S, R, KP)).sum(0)
        return a[:, :K] + a[:, KP:KP + K] + b[:, :K]
    def gx():
        a = torch.bmm(G2[0].view(M, S, C).transpose(0, 1), W0.view(S, C, 2 * KP)).sum(0)
        b = torch.bmm(G2[1].view(M, S, C).transpose(0, 1), W1.view(S, C, KP)).sum(0)
        return a[:, :K] + a[:, KP:KP + K] + b[:, :K]
    print("S %2d  g_w %.0f us err %.2e | g_x %.0f us err %.2e" % (
        S, timeit(gw), float((gw() - refw).abs().max() / refw.abs().max()),
        timeit(gx), float((gx() - refx).abs().max() / refx.abs().max())))

# variant: the remainder plane in BF16 (it only has to carry ~11 bits below the TF32 plane)
Glo = (gt - G2[0]).to(torch.bfloat16)
X1b = X1.to(torch.bfloat16); W1b = W1.to(torch.bfloat16)
for S in (8, 16):
    R, C = M // S, N // S
    def gw():
        a = torch.bmm(G2[0].view(S, R, N).transpose(1, 2), X0.view(S, R, 2 * KP)).sum(0)
        b = torch.bmm(Glo.view(S, R, N).transpose(1, 2), X1b.view(S, R, KP), out_dtype=torch.float32).sum(0)
        return a[:, :K] + a[:, KP:KP + K] + b[:, :K]
    def gx():
        a = torch.bmm(G2[0].view(M, S, C).transpose(0, 1), W0.view(S, C, 2 * KP)).sum(0)
        b = torch.bmm(Glo.view(M, S, C).transpose(0, 1), W1b.view(S, C, KP), out_dtype=torch.float32).sum(0)
        return a[:, :K] + a[:, KP:KP + K] + b[:, :K]
    print("bf16-lo S %2d  g_w %.0f us err %.2e | g_x %.0f us err %.2e" % (
        S, timeit(gw), float((gw() - refw).abs().max() / refw.abs().max()),
        timeit(gx), float((gx() - refx).abs().max() / refx.abs().max())))
