"""does the HBM-bound dL/dW product of qec_gen2_bwd overlap with the FP32-bound pool1 backward (qec_routing_small_bwd)
when they run on two streams?  serial vs concurrent timing at the BASELINE config-2 shapes.
QEC_G2_STAGES is not a knob: the kernel's ring depth is compile-time; the probe only tells whether CTAs co-reside."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qecnetworks_b200 import functional as QF, _lib
if os.environ.get("QEC_LIB"):  # a variant library (e.g. gen2_bwd built with -DQEC_G2_STAGES=3)
    _lib.LIB_PATH = os.path.abspath(os.environ["QEC_LIB"])

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, P, K, O, T = 32, 256, 9, 128, 3
lr = torch.randn(B, P, K, 1, 4, generator=g); lr = (lr / lr.norm(dim=-1, keepdim=True)).to(dev)
act = torch.rand(B, P, K, generator=g).to(dev)
xrot = torch.randn(B, P, K, 1, 3, generator=g).to(dev)
Wc = (torch.randn(4 * O, 3, generator=g) * 0.3).to(dev).requires_grad_(True)
bc = torch.randn(4 * O, generator=g).to(dev).requires_grad_(True)
alpha = torch.ones(1, device=dev, requires_grad=True); beta = torch.ones(1, device=dev, requires_grad=True)
pose, agr = QF.routing_small(xrot, lr, act, Wc, bc, alpha, beta, T)
Gp = torch.randn_like(pose); Ga = torch.randn_like(agr)
M, N, Kc = 8192, 20480, 40
gt = torch.randn(M, N, generator=g).to(dev); h = torch.randn(M, Kc, generator=g).to(dev)
W2 = (torch.randn(N, Kc, generator=g) / 6).to(dev); perm = torch.randperm(N, generator=g).to(dev)

def small_bwd():
    torch.autograd.grad([pose, agr], [Wc, bc, alpha, beta], [Gp, Ga], retain_graph=True)

def dw():
    QF.gen2_backward(gt, h, W2, perm, need_h=False, need_w=True)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n

side = torch.cuda.Stream(priority=0)        # low priority: the HBM-bound product fills what the chain leaves
hi = torch.cuda.Stream(priority=-1)         # high priority: the serial chain of the backward
def both_prio():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    side.wait_event(ev); hi.wait_event(ev)
    with torch.cuda.stream(side): dw()
    with torch.cuda.stream(hi): small_bwd()
    e1 = torch.cuda.Event(); e1.record(side); e2 = torch.cuda.Event(); e2.record(hi)
    cur.wait_event(e1); cur.wait_event(e2)

def both(first_dw=True):
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(main)
    side.wait_event(ev)
    if first_dw:
        with torch.cuda.stream(side): dw()
        small_bwd()
    else:
        small_bwd()
        with torch.cuda.stream(side): dw()
    ev2 = torch.cuda.Event(); ev2.record(side)
    main.wait_event(ev2)

print("pool1 backward alone   %7.1f us" % timeit(small_bwd))
print("gen2 dW alone          %7.1f us" % timeit(dw))
print("serial                 %7.1f us" % timeit(lambda: (small_bwd(), dw())))
print("two streams, dW first  %7.1f us" % timeit(lambda: both(True)))
print("two streams, dW second %7.1f us" % timeit(lambda: both(False)))
print("dW low prio first, pool1 backward high prio %7.1f us" % timeit(both_prio))
