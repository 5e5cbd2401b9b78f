"""probe: orientation of the skinny backward bmm's (output [rows, 96] vs its transpose [96, rows])"""
import torch
dev = torch.device("cuda:0")
M, N, KP = 8192, 20480, 48
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = True
g = torch.randn(M, N, device=dev); gl = torch.randn(M, N, device=dev).to(torch.bfloat16)
X0 = torch.randn(M, 2 * KP, device=dev); X1 = torch.randn(M, KP, device=dev).to(torch.bfloat16)
W0 = torch.randn(N, 2 * KP, device=dev); W1 = torch.randn(N, KP, device=dev).to(torch.bfloat16)
for S in (8, 16):
    R, C = M // S, N // S
    a1 = lambda: torch.bmm(g.view(S, R, N).transpose(1, 2), X0.view(S, R, 2 * KP))
    a2 = lambda: torch.bmm(X0.view(S, R, 2 * KP).transpose(1, 2), g.view(S, R, N))
    b1 = lambda: torch.bmm(gl.view(S, R, N).transpose(1, 2), X1.view(S, R, KP), out_dtype=torch.float32)
    b2 = lambda: torch.bmm(X1.view(S, R, KP).transpose(1, 2), gl.view(S, R, N), out_dtype=torch.float32)
    print("S %2d g_w tf32: [N,96] %.0f us | [96,N] %.0f us (maxdiff %.1e) || bf16: [N,48] %.0f | [48,N] %.0f" % (
        S, timeit(a1), timeit(a2), float((a1().transpose(1, 2) - a2()).abs().max()), timeit(b1), timeit(b2)))
    c1 = lambda: torch.bmm(g.view(M, S, C).transpose(0, 1), W0.view(S, C, 2 * KP))
    c2 = lambda: torch.bmm(W0.view(S, C, 2 * KP).transpose(1, 2), g.view(M, S, C).permute(1, 2, 0))
    d1 = lambda: torch.bmm(gl.view(M, S, C).transpose(0, 1), W1.view(S, C, KP), out_dtype=torch.float32)
    d2 = lambda: torch.bmm(W1.view(S, C, KP).transpose(1, 2), gl.view(M, S, C).permute(1, 2, 0), out_dtype=torch.float32)
    print("S %2d g_x tf32: [M,96] %.0f us | [96,M] %.0f us (maxdiff %.1e) || bf16: [M,48] %.0f | [48,M] %.0f" % (
        S, timeit(c1), timeit(c2), float((c1().transpose(1, 2) - c2()).abs().max()), timeit(d1), timeit(d2)))
