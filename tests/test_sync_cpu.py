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
