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


def test_direct_gradient_placement_matches_accumulating_bucket():
    """FlatGradBucket(direct=True): the backward kernels write every parameter gradient of QEC-Net straight into
    the flat buffer (autograd adopts the aliases as .grad: no add launch per parameter).  Three training steps give
    the same gradients and parameters as the accumulating bucket; a parameter that enters the graph twice (the
    Siamese net) is caught by check_views()."""
    import copy

    import qecnetworks_b200 as qb
    from qecnetworks_b200 import _lib, ddp
    from qecnetworks_b200.synthetic import synthetic_batch

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net_a = qb.QecNet(2048, 128, 40, 3).to(dev)   # BASELINE config 2 geometry (cluster + register-resident regimes)
    net_b = copy.deepcopy(net_a)
    buck_a = ddp.FlatGradBucket(net_a.parameters(), early=list(net_a.pool2.parameters()), direct=True)
    buck_b = ddp.FlatGradBucket(net_b.parameters(), early=list(net_b.pool2.parameters()))
    opt_a, opt_b = ddp.FlatAdam(buck_a, lr=1e-3), ddp.FlatAdam(buck_b, lr=1e-3)
    launches = []
    for step in range(3):
        points, lrfs, act, labels = [t.to(dev) for t in synthetic_batch(2, ncls=40, seed=50 + step)]
        grads = []
        for net, buck in ((net_a, buck_a), (net_b, buck_b)):
            n0 = _lib.launches()
            _, a = net(points, lrfs, act, None)
            net.spread_loss(a, labels).backward()
            launches.append(_lib.launches() - n0)
            buck.check_views()   # every .grad aliases its slice of the flat buffer
            grads.append(buck.flat.clone())
        scale = float(grads[1].abs().max())
        # same kernels either way; what differs run to run is the order of the FP32 atomics: dL/dlrfs of pool2 feeds the
        # whole pool1 backward (~1e-6 of the largest entry), and dL/d(alpha, beta) are themselves sums accumulated with
        # atomics -- pool2.beta a near-cancelling one that moves by up to 1e-4 of the largest entry from run to run
        # (scripts/flaky_hunt.py).  A lost or doubled gradient would be O(1) of the tensor's own size.
        off = 0
        for p in buck_a.params:
            n = p.numel()
            da, db = grads[0][off:off + n], grads[1][off:off + n]
            off += n
            atomic_sum = n == 1   # alpha / beta
            tol = 1e-3 if atomic_sum else 5e-5
            err = float((da - db).abs().max()) / scale
            print("direct margin grads step=%d numel=%d err=%.2e tol=%.2e" % (step, n, err, tol))
            assert err < tol, (n, err)
            own = float(db.abs().max())
            assert own > 0.0 and float((da - db).abs().max()) < 0.5 * own + 1e-4 * scale, (n, own)   # present, not doubled
        opt_a.step()
        opt_b.step()
        assert float(buck_a.flat.abs().max()) == 0.0 and all(p.grad is None for p in net_a.parameters())
        # (parameters are not compared after the Adam steps: Adam is scale-free per element, so an entry whose gradient
        # is at the level of the atomics' noise may step in either direction; the nets are re-synchronised instead)
        with torch.no_grad():
            for p, q in zip(net_a.parameters(), net_b.parameters()):
                q.copy_(p)
    assert launches[0] == launches[1]   # the same kernels of this package run either way (what is saved are ATen adds)
    # a parameter used twice in one graph: the second contribution cannot share the home
    sia = qb.QecSiaNet(2048, 16, 10, 3).to(dev)
    buck_s = ddp.FlatGradBucket(sia.parameters(), direct=True)
    points, lrfs, act, labels = [t.to(dev) for t in synthetic_batch(2, ncls=10, seed=9)]
    _, x = sia(torch.stack([points, points], 1), torch.stack([lrfs, lrfs], 1), torch.stack([act, act], 1), None)
    sia.spread_loss(x, torch.stack([labels, labels], 1)).backward()
    with pytest.raises(RuntimeError):
        buck_s.check_views()
