"""world_size-2 gloo test of the data-parallel host logic: batch sharding and the flat-arena
gradient all-reduce + Adam.  The two CUDA kernels FlatAdam calls (multi-tensor gather into the arena, fused Adam) are replaced
by torch-CPU stand-ins HERE ONLY, so that the collective / averaging logic runs without a GPU; the kernels
themselves are covered by the -m gpu tests."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_gather(srcs, offs, cnts, n, flat):
    """host stand-in for dvd_gather_flat: same (pointer, offset, count) table, plain memmove / memset"""
    import ctypes
    base = flat.data_ptr()
    for i in range(n):
        dst = base + 4 * offs[i]
        if srcs[i] is None:
            ctypes.memset(dst, 0, 4 * cnts[i])
        else:
            ctypes.memmove(dst, srcs[i], 4 * cnts[i])


def _cpu_adam(p, g, m, v, lr, b1, b2, eps, t, scale):
    g = g * scale
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    p.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)


def _worker(rank, world, port, q, shard=False, overlap=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dvdgan_b200.trainer import FlatAdam, shard_batch
    FlatAdam._gather = staticmethod(_cpu_gather)
    FlatAdam._adam = staticmethod(_cpu_adam)
    torch.manual_seed(0)                       # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    opt = FlatAdam(net, 1e-2, (0.0, 0.9), shard=(world, rank) if shard else None)
    assert all(p.data_ptr() >= opt.flat_p.data_ptr() for p in net.parameters())   # params are arena views
    if overlap:     # one all-reduce per top-level child, started by a hook as soon as the child's gradients exist
        opt.enable_overlap(net)
        assert opt._buckets == [(0, 2), (2, 4)]
    if shard:       # 41 parameters over 2 ranks: arena padded to 42, moments exist for this rank's 21 only
        assert opt.flat_p.numel() == 42 and opt.m.numel() == 21 and opt.v.numel() == 21
    x = torch.randn(8, 6)
    y = torch.randn(8, 1)
    sl = shard_batch(8, world, rank)
    for _ in range(3):
        opt.zero_grad()
        loss = ((net(x[sl]) - y[sl]) ** 2).mean()
        loss.backward()
        if overlap:
            assert len(opt._pending) == 2          # both buckets were launched during the backward pass
        opt.step(world)
    q.put((rank, opt.flat_p[:opt.numel].clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shard,overlap", [(False, False), (True, False), (False, True)],
                         ids=["allreduce", "reduce_scatter_sharded_adam_all_gather", "bucketed_allreduce_during_backward"])
def test_flat_adam_allreduce_matches_global_batch(shard, overlap):
    world = 2
    port = 29500 + (os.getpid() + 7 * shard + 13 * overlap) % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, shard, overlap)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.allclose(res[0], res[1], atol=0, rtol=0)          # replicas stay bit-identical
    # single-process reference on the GLOBAL batch with torch.optim.Adam
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    opt = torch.optim.Adam(net.parameters(), 1e-2, (0.0, 0.9))
    x = torch.randn(8, 6)
    y = torch.randn(8, 1)
    for _ in range(3):
        opt.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        opt.step()
    ref = torch.cat([p.data.reshape(-1) for p in net.parameters()])
    assert torch.allclose(res[0], ref, atol=1e-6, rtol=1e-5)


def test_shard_batch():
    from dvdgan_b200.trainer import shard_batch
    assert [shard_batch(256, 8, r) for r in (0, 7)] == [slice(0, 32), slice(224, 256)]
    with pytest.raises(ValueError):
        shard_batch(10, 4, 0)
