"""GPU, needs >= 2 devices (skipped otherwise; run with `gpurun --gpus 2`): the data-parallel step over
NCCL gives every rank the gradient of the GLOBAL-batch mean loss."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, direct=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import qecnetworks_b200 as qb
    from qecnetworks_b200 import ddp
    from qecnetworks_b200.synthetic import synthetic_batch

    ddp.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(7)
    net = qb.QecNet(2048, 16, 10, 3).to(dev)
    ddp.broadcast_parameters(net, 0)
    # overlapped slice; direct: the backward kernels write into the flat buffer themselves (what bench.py runs)
    bucket = ddp.FlatGradBucket(net.parameters(), early=list(net.pool2.parameters()), direct=direct)
    B = 4
    points, lrfs, act, labels = [t.to(dev) for t in synthetic_batch(B * world, ncls=10, seed=11)]
    lo, hi = ddp.shard_batch(B * world, rank, world)
    bucket.zero()
    _, a = net(points[lo:hi], lrfs[lo:hi], act[lo:hi], None)
    net.spread_loss(a, labels[lo:hi]).backward()
    bucket.all_reduce_mean()
    got = bucket.flat.clone()
    # the same global batch on one device (a local backward: no collective may be launched by the hooks)
    bucket.overlap = False
    bucket.zero()
    _, a = net(points, lrfs, act, None)
    net.spread_loss(a, labels).backward()
    ref = bucket.flat.clone()
    err = float((got - ref).abs().max() / ref.abs().max())
    # the training step of bench.py: the SUM stays in the bucket, FlatAdam applies the 1/world it is owed and
    # leaves the bucket zeroed; every rank must end with the parameters torch.optim.Adam gives for the same
    # averaged gradient
    import copy

    ref_net = copy.deepcopy(net)  # parameters as they are now; its own gradient tensors
    topt = torch.optim.Adam(ref_net.parameters(), lr=1e-3)
    opt = ddp.FlatAdam(bucket, lr=1e-3)
    bucket.overlap = True
    bucket.zero()
    _, a = net(points[lo:hi], lrfs[lo:hi], act[lo:hi], None)
    net.spread_loss(a, labels[lo:hi]).backward()
    owed = bucket.all_reduce_mean(average=False)
    assert owed == 1.0 / world
    for p, q in zip(net.parameters(), ref_net.parameters()):
        q.grad = p.grad.detach().clone() * owed
    gerr = float((bucket.flat * owed - ref).abs().max() / ref.abs().max())  # still the global-batch gradient
    topt.step()
    opt.step(grad_scale=owed)
    perr = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(net.parameters(), ref_net.parameters()))
    zeroed = float(bucket.flat.abs().max()) == 0.0
    out[rank] = max(err, gerr) if (perr < 2e-6 and zeroed) else 1.0 + perr
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradients_match_global_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    assert out[0] < 1e-4 and out[1] < 1e-4, dict(out)


def test_ddp_direct_gradient_placement_match_global_batch():
    """the same with FlatGradBucket(direct=True): parameter gradients written by the kernels straight into the flat
    buffer, the early slice all-reduced from the post-accumulate hooks, FlatAdam re-arming the slices"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29741, out, True), nprocs=2, join=True)
    assert out[0] < 1e-4 and out[1] < 1e-4, dict(out)


def test_dataparallel_matches_single_gpu():
    """the reference's own multi-GPU mode (train_cls.py:27,62-65): nn.DataParallel over the drop-in QecNet --
    ONE process, one Python thread per GPU, replicas re-created every step -- forward and backward on two GPUs
    against the single-GPU result.  pool2 runs on the cluster kernels (4 096 votes per capsule, > 48 KB of
    dynamic shared memory), whose opt-in attribute is per device: the launch configuration cached per device
    ordinal is what this test pins."""
    import qecnetworks_b200 as qb
    from qecnetworks_b200.synthetic import synthetic_batch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = qb.QecNet(2048, 16, 10, 3).to(dev)
    B = 4
    points, lrfs, act, labels = [t.to(dev) for t in synthetic_batch(B, ncls=10, seed=13)]
    idx = torch.zeros(B, 256, 9, dtype=torch.int64, device=dev)   # pointspool1index: scattered like the reference's

    def step(model):
        for p in net.parameters():
            p.grad = None
        pose, a = model(points, lrfs, act, idx)
        loss = net.spread_loss(a, labels)          # on the gathered output, as train_cls.py:62-64 does
        loss.backward()
        return pose.detach().clone(), a.detach().clone(), {k: v.grad.detach().clone() for k, v in net.named_parameters()}

    pose1, a1, g1 = step(net)
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    for _ in range(2):   # twice: the second step finds every per-device cache filled
        pose2, a2, g2 = step(dp)
        assert pose2.device == dev and pose2.shape == pose1.shape
        d = torch.minimum((pose1 - pose2).abs().amax(-1), (pose1 + pose2).abs().amax(-1))
        assert float(d.max()) < 1e-5
        assert float((a1 - a2).abs().max()) < 1e-5
        scale = max(float(v.abs().max()) for v in g1.values())
        for k in g1:
            e = float((g1[k] - g2[k]).abs().max()) / max(float(g1[k].abs().max()), 0.25 * scale)
            assert e < 1e-4, (k, e)
