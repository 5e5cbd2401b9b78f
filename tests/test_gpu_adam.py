"""qec_adam_flat (FlatAdam) against torch.optim.Adam -- the optimizer the reference trains with
(train_cls.py:44-51) -- on the same parameters and gradient sequences."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(dev, seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


@pytest.mark.parametrize("weight_decay,scale", [(0.0, 1.0), (0.01, 0.125)])
def test_flat_adam_matches_torch_adam(weight_decay, scale):
    from qecnetworks_b200 import ddp

    dev = torch.device("cuda:0")
    shapes = [(40, 41), (7,), (3, 5, 2), (1,), (1001,)]  # total not a multiple of four
    ours = _params(dev, 1, shapes)
    ref = _params(dev, 1, shapes)
    bucket = ddp.FlatGradBucket(ours, early=ours[2:3])
    opt = ddp.FlatAdam(bucket, lr=1e-3, weight_decay=weight_decay)
    topt = torch.optim.Adam(ref, lr=1e-3, weight_decay=weight_decay)
    g = torch.Generator().manual_seed(2)
    for step in range(25):
        for p, q in zip(ours, ref):
            gr = torch.randn(p.shape, generator=g).to(dev) * (10.0 ** (step % 5 - 2))
            p.grad.copy_(gr / scale)  # the kernel reads grad * scale
            q.grad = gr.clone()
        opt.step(grad_scale=scale)
        topt.step()
        assert float(bucket.flat.abs().max()) == 0.0  # the step leaves the gradient buffer zeroed
    assert opt.steps_taken() == 25
    for p, q in zip(ours, ref):
        err = float((p.detach() - q.detach()).abs().max())
        assert err <= 2e-6 * max(1.0, float(q.detach().abs().max())), err
    sd = topt.state_dict()["state"]
    off = 0
    order = {id(p): i for i, p in enumerate(ours)}
    for p in bucket.params:
        st = sd[order[id(p)]]
        n = p.numel()
        m = opt.exp_avg[off:off + n].view_as(p)
        v = opt.exp_avg_sq[off:off + n].view_as(p)
        assert float((m - st["exp_avg"]).abs().max()) <= 1e-6 * max(1.0, float(st["exp_avg"].abs().max()))
        assert float((v - st["exp_avg_sq"]).abs().max()) <= 1e-6 * max(1.0, float(st["exp_avg_sq"].abs().max()))
        off += n


def test_flat_adam_in_cuda_graph():
    """the step count lives on the device: replays of a captured step keep advancing the bias corrections"""
    from qecnetworks_b200 import ddp

    dev = torch.device("cuda:0")
    shapes = [(64, 33), (5,)]
    ours = _params(dev, 5, shapes)
    ref = _params(dev, 5, shapes)
    bucket = ddp.FlatGradBucket(ours)
    opt = ddp.FlatAdam(bucket, lr=1e-2)
    topt = torch.optim.Adam(ref, lr=1e-2)
    src = torch.randn(bucket.flat.numel(), generator=torch.Generator().manual_seed(9)).to(dev)

    def step():
        bucket.flat.copy_(src)
        opt.step()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert opt.steps_taken() == 5  # one eager step + four replays (the capture itself executes nothing)
    off = 0
    for p, q in zip(bucket.params, ref):  # no early slice: the bucket keeps the given order
        q.grad = src[off:off + p.numel()].view_as(p).clone()
        off += p.numel()
    for _ in range(5):
        topt.step()
    for p, q in zip(ours, ref):
        err = float((p.detach() - q.detach()).abs().max())
        assert err <= 2e-6 * max(1.0, float(q.detach().abs().max())), err
