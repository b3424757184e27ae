"""Gradient exchange (vfm_vae_b200/sync.py, mirror of the reference's sync_grads) on 2 gloo ranks: the averaged
gradients of a batch-sharded step equal the single-process gradients of the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vfm_vae_b200 import sync
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    if rank != 0:
        for p in net.parameters():
            p.data.add_(1.0)          # replicas start different; broadcast must fix that
    sync.broadcast_module(net)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 8, generator=g)
    shard = x[rank * 4:(rank + 1) * 4]
    net(shard).square().mean().backward()
    # poison one entry on one rank: the reference semantics turn NaN/inf into finite numbers after the reduce
    params = list(net.parameters())
    if rank == 1:
        params[1].grad[0] = float('inf')
    sync.sync_grads(params, gain=1.0)
    # by value (numpy): a tensor in an mp.Queue travels as a shared-memory fd that dies with this process if the parent is slow to fetch it
    q.put((rank, [p.grad.numpy().copy() for p in params], [p.data.numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_grads_two_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict()
    for _ in range(2):
        rank, grads, params = q.get(timeout=120)
        out[rank] = ([torch.from_numpy(a) for a in grads], [torch.from_numpy(a) for a in params])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas agree
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b)
    for a, b in zip(out[0][1], out[1][1]):
        assert torch.equal(a, b)
    # and equal the full-batch gradient (mean of two equal-sized shards == mean over the batch)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 8, generator=g)
    net(x).square().mean().backward()
    ref = [p.grad for p in net.parameters()]
    for i, (a, b) in enumerate(zip(out[0][0], ref)):
        if i == 1:
            assert a[0].item() == 1e5          # inf/2 -> nan_to_num(posinf=1e5)
            assert torch.allclose(a[1:], b[1:], atol=1e-6)
        else:
            assert torch.allclose(a, b, atol=1e-6)


def test_sharded_all_mean_single_process():
    from vfm_vae_b200 import sync
    t = torch.arange(10, dtype=torch.float32)
    assert torch.equal(sync.sharded_all_mean(t.clone(), shard_size=3), t)


# ------------------------------------------------------------------ GradExchange: overlapped, bucketed, same result as sync_grads

def _make_net():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    unused = torch.nn.Parameter(torch.ones(5))           # never reached by backward: must end up with grad None, like the reference
    return net, unused


def _exchange_worker(rank, world, port, q, accumulate):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vfm_vae_b200 import sync
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 8, generator=g)
    shard = x[rank * 8:(rank + 1) * 8]
    gain = 2 if accumulate else None
    # A: the functional mirror of the reference (post hoc)
    net, unused = _make_net()
    params = list(net.parameters()) + [unused]
    if accumulate:
        net(shard[:4]).square().mean().backward()
        net(shard[4:]).square().mean().backward()
    else:
        net(shard).square().mean().backward()
    if rank == 1:
        params[1].grad[0] = float('inf')
        params[2].grad[0, 0] = float('nan')
    sync.sync_grads(params, gain=gain)
    want = [None if p.grad is None else p.grad.clone() for p in params]
    # B: the overlapped bucketed exchange (tiny buckets so that several all_reduces are in flight during backward)
    net, unused = _make_net()
    params = list(net.parameters()) + [unused]
    ex = sync.GradExchange(params, bucket_elems=16)
    assert ex.stats['buckets'] >= 3
    for step in range(2):                                  # twice: the buffer and hooks are persistent across steps
        ex.zero_grad()
        if accumulate:
            with ex.no_sync():
                net(shard[:4]).square().mean().backward()
            net(shard[4:]).square().mean().backward()
        else:
            net(shard).square().mean().backward()
        launched = ex.stats['launched_in_backward']
        if rank == 1:
            with torch.no_grad():
                params[1].grad[0] = float('inf')           # poison after the bucket left: only the mirror semantics matter for finite parts
        ex.finish(gain=gain)
    got = [None if p.grad is None else p.grad.clone() for p in params]
    views_ok = all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(ex.params, ex.views))
    q.put((rank, [None if t is None else t.numpy().copy() for t in want], [None if t is None else t.numpy().copy() for t in got], launched, views_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('accumulate', [False, True])
def test_grad_exchange_matches_sync_grads_two_ranks(accumulate):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q, accumulate)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, want, got, launched, views_ok = q.get(timeout=120)
        out[rank] = (want, got, launched, views_ok)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import numpy as np
    for rank in (0, 1):
        want, got, launched, views_ok = out[rank]
        assert views_ok
        assert launched >= 2, 'buckets must leave while backward is still running'
        assert want[-1] is None and got[-1] is None         # the unused parameter
        for i, (a, b) in enumerate(zip(want[:-1], got[:-1])):
            if i == 1:                                        # the poisoned bias: entry 0 differs by construction (poisoned after the send in B)
                assert a[0] == 1e5
                assert np.array_equal(a[1:], b[1:])
            elif i == 2:
                assert a[0, 0] == 0.0                          # nan -> 0
                assert np.array_equal(a.reshape(-1)[1:], b.reshape(-1)[1:])
            else:
                assert np.array_equal(a, b), i                 # bit-exact: a SUM over 2 ranks does not depend on the sharding
    for a, b in zip(out[0][1][:-1], out[1][1][:-1]):
        assert np.array_equal(a, b)                           # replicas agree


def test_grad_exchange_single_process_finalize_and_views():
    from vfm_vae_b200 import sync
    net, unused = _make_net()
    params = list(net.parameters()) + [unused]
    ex = sync.GradExchange(params, bucket_elems=32)
    x = torch.randn(4, 8)
    ex.zero_grad()
    net(x).square().mean().backward()
    with torch.no_grad():
        params[0].grad[0, 0] = float('-inf')
    ex.finish(gain=3)
    ref_net, _ = _make_net()
    ref_net(x).square().mean().backward()
    for i, (p, r) in enumerate(zip(net.parameters(), ref_net.parameters())):
        want = torch.nan_to_num(r.grad * 3, nan=0.0, posinf=1e5, neginf=-1e5)
        if i == 0:
            assert p.grad[0, 0].item() == -1e5
            want[0, 0] = -1e5
        assert torch.equal(p.grad, want)
    assert unused.grad is None
    # layout: reverse parameter order, 16-byte aligned slots
    assert ex.offsets[len(ex.params) - 1] == 0 and all(o % 4 == 0 for o in ex.offsets)
