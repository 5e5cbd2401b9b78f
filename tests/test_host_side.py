"""CPU: host-side logic -- the C-ABI library loads and exports every symbol that
include/qec_b200.h declares, the drop-in modules keep the reference's constructor /
state_dict contract, the product refuses CPU tensors instead of falling back, and the
data-parallel gradient bucket works across 2 gloo ranks."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "qec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int|long long)\s+(qec_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    from qecnetworks_b200 import build, _lib

    build.build()  # nvcc cross-compiles for sm_100a without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.qec_b200_version() == 100
    # ... and nothing undeclared: every qec_* function the library exports is part of the documented C ABI
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("qec_")}
    # (qec_prof_read / qec_prof3_read: -DQEC_PROF builds only)
    assert exported <= set(names) | {"qec_prof_read", "qec_prof3_read"}, exported - set(names)


def test_argument_validation_without_gpu():
    """negative return = index of the offending argument; no kernel is launched"""
    from qecnetworks_b200 import _lib

    lib = _lib.load()
    assert lib.qec_routing_fwd(None, None, None, None, None, None, None, None, 1, 1, 1, 3, None) == -1
    assert lib.qec_votes_fwd(None, None, None, 1, 1, 1, 1, None) == -1
    assert lib.qec_symeig4_fwd(None, None, None, None, 4, None) == -1
    with pytest.raises(_lib.QecLibraryError):
        _lib.call("qec_votes_fwd", None, None, None, 1, 1, 1, 1, None)


def test_module_contract_matches_reference():
    import qecnetworks_b200 as qb

    m = qb.QecModule(in_channels=4, out_channels=6, num_iterations=3, num_neighbours=5, num_patches=3)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "alpha": (1,), "beta": (1,),
        "quater_gen1.weight": (6, 12), "quater_gen1.bias": (6,),
        "quater_gen2.weight": (96, 6), "quater_gen2.bias": (96,),
    }
    assert float(sd["alpha"]) == 1.0 and float(sd["beta"]) == 1.0
    net = qb.QecNet(2048, 16, 10, 3)
    keys = set(net.state_dict())
    assert keys == {p + k for p in ("pool1.", "pool2.") for k in sd}
    assert net.pool1.num_neighbours == 9 and net.pool1.num_patches == 256
    assert net.pool2.num_neighbours == 256 and net.pool2.num_patches == 1
    assert set(qb.QecSiaNet(2048, 16, 10, 3).state_dict()) == keys


def test_no_cpu_fallback():
    import qecnetworks_b200 as qb
    from qecnetworks_b200._lib import QecLibraryError

    m = qb.QecModule(1, 4, 3, 9, 8)
    with pytest.raises(QecLibraryError):
        m(torch.zeros(2, 8, 9, 3), torch.zeros(2, 8, 9, 1, 4), torch.zeros(2, 8, 9, 1))
    with pytest.raises(QecLibraryError):
        qb.symeig(torch.zeros(3, 4, 4))
    with pytest.raises(NotImplementedError):
        qb.symeig(torch.zeros(3, 5, 5))
    with pytest.raises(QecLibraryError):  # the loss, the kNN indices and the eval ops are CUDA-only too
        qb.QecNet(2048, 4, 5, 3).spread_loss(torch.rand(2, 5), torch.tensor([1, 2]))
    with pytest.raises(QecLibraryError):
        qb.functional.knn_indices(torch.zeros(1, 8, 3), 3, centres=torch.zeros(1, 2, 3))
    with pytest.raises(QecLibraryError):
        qb.eval_ops.ave_pose(torch.zeros(2, 5, 4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qecnetworks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_losses_match_oracle():
    from oracle import qec_oracle as orc
    import qecnetworks_b200 as qb

    from qecnetworks_b200._lib import QecLibraryError

    g = torch.Generator().manual_seed(1)
    net = qb.QecNet(2048, 4, 10, 3)
    # the losses are CUDA kernels (tests/test_gpu_kernels.py::test_spread_loss_kernel, ::test_class_pose_loss_kernel):
    # on the CPU they raise like the rest of the path; the oracle's statements are pinned by the golden vectors
    pose = torch.randn(2, 6, 10, 4, generator=g)
    pose = pose / pose.norm(dim=-1, keepdim=True)
    with pytest.raises(QecLibraryError):
        net.pose_diff_loss(pose)
    d = (2.0 / 3.141592653589793) * torch.acos((pose[0] * pose[1]).sum(-1).abs().clamp(max=0.9999))
    assert abs(float(d.mean()) - float(orc.pose_diff_loss(pose))) < 1e-7


def test_synthetic_batch_shapes():
    from qecnetworks_b200.synthetic import synthetic_batch, siamese_pair

    points, lrfs, act, labels = synthetic_batch(3, ncls=40, seed=7)
    assert points.shape == (3, 256, 9, 3) and lrfs.shape == (3, 256, 9, 4) and act.shape == (3, 256, 9)
    n = lrfs.norm(dim=-1)
    assert torch.all((n - act).abs() < 1e-5)  # unit where active, zero where not
    assert float(lrfs[..., 0].min()) >= 0.0
    assert torch.all(act[:, :, 0] == 1)
    p2, l2, a2 = siamese_pair(points, lrfs, act)
    assert p2.shape == (3, 2, 256, 9, 3) and l2.shape == (3, 2, 256, 9, 4) and a2.shape == (3, 2, 256, 9)
    assert torch.all((l2[:, 1].norm(dim=-1) - act).abs() < 1e-5)


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from qecnetworks_b200 import ddp
    import qecnetworks_b200 as qb

    r, w, _ = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    net = qb.QecNet(2048, 4, 5, 3)
    ddp.broadcast_parameters(net, 0)
    bucket = ddp.FlatGradBucket(net.parameters())
    assert bucket.nbytes() == 4 * sum(p.numel() for p in net.parameters())
    bucket.zero()
    for i, p in enumerate(net.parameters()):
        p.grad.add_(float(rank + 1) * (i + 1))  # grads are views into the flat buffer
    bucket.all_reduce_mean()
    expect = sum(range(1, world + 1)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, expect * (i + 1))) for i, p in enumerate(net.parameters()))
    # early slice: reduced asynchronously from the post-accumulate hooks of a real backward, joined afterwards
    net2 = qb.QecNet(2048, 4, 5, 3)
    ddp.broadcast_parameters(net2, 0)
    b2 = ddp.FlatGradBucket(net2.parameters(), early=list(net2.pool2.parameters()))
    assert b2.n_early == sum(p.numel() for p in net2.pool2.parameters())
    b2.zero()
    loss = sum((p * float(rank + 1) * (i + 1)).sum() for i, p in enumerate(net2.parameters()))
    loss.backward()
    launched_early = b2._work is not None
    b2.all_reduce_mean()
    ok = ok and launched_early and all(
        torch.allclose(p.grad, torch.full_like(p, expect * (i + 1))) for i, p in enumerate(net2.parameters()))
    # average=False: the buffer keeps the SUM and the caller is told what is owed (FlatAdam's grad_scale)
    b2.zero()
    b2.overlap = False
    for i, p in enumerate(net2.parameters()):
        p.grad.add_(float(rank + 1) * (i + 1))
    owed = b2.all_reduce_mean(average=False)
    ok = ok and owed == 1.0 / world and all(
        torch.allclose(p.grad, torch.full_like(p, expect * world * (i + 1))) for i, p in enumerate(net2.parameters()))
    first = next(net.parameters()).detach().clone()
    gathered = [torch.zeros_like(first) for _ in range(world)]
    dist.all_gather(gathered, first)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    lo, hi = ddp.shard_batch(32, rank, world)
    out[rank] = bool(ok and same and (hi - lo) == 32 // world)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29611 + os.getpid() % 200
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_flat_adam_host_layout_and_no_cpu_fallback():
    """FlatAdam re-points the parameters at one flat buffer in the bucket's order (values unchanged, early
    slice first) and refuses to step on CPU tensors: the optimizer kernel is CUDA-only like the rest."""
    import qecnetworks_b200 as qb
    from qecnetworks_b200 import ddp, _lib

    torch.manual_seed(3)
    net = qb.QecNet(2048, 4, 5, 3)
    before = {k: v.detach().clone() for k, v in net.named_parameters()}
    bucket = ddp.FlatGradBucket(net.parameters(), early=list(net.pool2.parameters()))
    opt = ddp.FlatAdam(bucket, lr=1e-3)
    assert bucket.all_reduce_mean(average=False) == 1.0  # single process: nothing owed
    off = 0
    for p in bucket.params:
        n = p.numel()
        assert p.data_ptr() == opt.flat.data_ptr() + 4 * off  # a view of the flat parameter buffer ...
        assert p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off  # ... at the offset of its gradient
        off += n
    assert off == opt.flat.numel() == bucket.flat.numel()
    for k, v in net.named_parameters():
        assert torch.equal(v.detach(), before[k])
    assert bucket.params[0] is next(net.pool2.parameters())
    with pytest.raises(_lib.QecLibraryError):
        opt.step()


def test_fused_build_selector_without_gpu():
    """qec_routing_fused_set_wide: returns the previous mode, ignores out-of-range values (query only); the
    default is the 256-thread build unless QEC_FUSED_WIDE says otherwise"""
    from qecnetworks_b200 import _lib

    lib = _lib.load()
    start = lib.qec_routing_fused_set_wide(-1)
    assert start in (0, 1, 2, 3)
    if "QEC_FUSED_WIDE" not in os.environ:
        assert start == 0
    try:
        assert lib.qec_routing_fused_set_wide(3) == start
        assert lib.qec_routing_fused_set_wide(99) == 3   # out of range: unchanged
        assert lib.qec_routing_fused_set_wide(1) == 3
        # argument validation happens before the build is chosen: no launch without a GPU either way
        assert lib.qec_routing_fused_fwd(None, None, None, None, None, None, None, None, None, 1, 256, 128, 4, 3, None) == -1
    finally:
        lib.qec_routing_fused_set_wide(start)
    assert lib.qec_routing_fused_set_wide(-1) == start


def test_gradient_homes_and_zero_split_host_logic():
    """direct gradient placement (FlatGradBucket(direct=True)): a backward node gets a parameter's home in the flat
    buffer once per release and only while the parameter has no .grad; autograd adopts the returned alias as .grad
    without a copy (checked with a toy Function on the CPU -- the real kernels are exercised by the GPU tests)"""
    import torch
    from qecnetworks_b200 import ddp
    from qecnetworks_b200 import functional as QF

    a, b, c = _z(3), _z(2), _z(1)
    ps = [torch.nn.Parameter(t + 1) for t in (a, b, c)]
    bucket = ddp.FlatGradBucket(ps, direct=True)
    assert all(p.grad is None for p in ps)

    class Toy(torch.autograd.Function):
        @staticmethod
        def forward(ctx, p, q, x):
            ctx.save_for_backward(p, q, x)
            return (p * x).sum() + q.sum() * 2.0

        @staticmethod
        def backward(ctx, g):
            p, q, x = ctx.saved_tensors
            gp = QF._home_out(p)
            assert gp is not None and float(gp.abs().max()) == 0.0   # zeroed home
            gp.copy_(x * g)
            assert QF._home_out(p) is None                          # handed out once per release
            gq = QF._home_out(q)
            gq.fill_(2.0 * float(g))
            return gp, gq, None

    x = torch.tensor([1.0, 2.0, 3.0])
    Toy.apply(ps[0], ps[1], x).backward()
    base = bucket.flat.data_ptr()
    assert ps[0].grad.data_ptr() == base and ps[1].grad.data_ptr() == base + 12
    assert bucket.flat.tolist() == [1.0, 2.0, 3.0, 2.0, 2.0, 0.0]
    assert QF._home_out(ps[0]) is None                              # .grad exists: autograd must accumulate instead
    with pytest.raises(RuntimeError):
        bucket.check_views()                                        # ps[2] got no gradient: not a view of the buffer
    bucket.zero()                                                   # zero + release
    assert all(p.grad is None for p in ps) and float(bucket.flat.abs().max()) == 0.0
    assert QF._home_out(ps[2]) is not None
    # alpha/beta pairs: adjacent homes give one 2-float tensor
    bucket.release()
    pair = QF._home_pair_out(ps[1], ps[2])
    assert pair is None                                             # ps[1] has 2 elements: ps[2] is not at offset + 1
    al, be = torch.nn.Parameter(torch.ones(1)), torch.nn.Parameter(torch.ones(1))
    b2 = ddp.FlatGradBucket([al, be], direct=True)
    pair = QF._home_pair_out(al, be)
    assert pair is not None and pair.shape == (2,) and pair.data_ptr() == b2.flat.data_ptr()
    assert QF._home_pair_out(al, be) is None
    # one zero fill, several tensors, 16-byte aligned pieces, None passes through
    z = QF._zeros_split(torch.device("cpu"), (2, 3), None, (5,), (1,))
    assert z[1] is None and [tuple(t.shape) for t in (z[0], z[2], z[3])] == [(2, 3), (5,), (1,)]
    assert all(t.data_ptr() % 16 == 0 and float(t.abs().max()) == 0.0 for t in (z[0], z[2], z[3]))
    assert z[2].data_ptr() - z[0].data_ptr() == 32 and z[3].data_ptr() - z[2].data_ptr() == 32


def _z(n):
    import torch

    return torch.zeros(n)
