"""Gradient exchange for the batch-sharded multi-GPU training step.

The decoder hot path shards by image batch (every op is per-sample; weights are replicated), so the forward/decode
needs no communication.  Training has exactly one exchange step: the gradient mean over ranks after backward.
This mirrors the reference's manual "DDP" (training/training_loop.py:272-289 ``sync_grads`` -> ``sharded_all_mean``):

    flat fp32 concat of all grads -> all_reduce(SUM) in shards of <= 2**23 elements -> / world_size -> * gain
    -> nan_to_num(nan=0, posinf=1e5, neginf=-1e5) -> scatter back

``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing; the bias / weight / noise
gradient reductions *inside* a rank are done by the kernels (bias_act db, modconv wgrad) before they join this all-reduce.
"""
import torch
import torch.distributed as dist

SHARD_ELEMS = 2 ** 23   # 32 MiB of fp32 per all_reduce, like the reference


def sharded_all_mean(tensor, shard_size=SHARD_ELEMS, group=None):
    """In-place mean over ranks of a flat tensor, reduced in fixed-size shards."""
    assert tensor.dim() == 1
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        for shard in tensor.tensor_split(max(1, -(-tensor.numel() // shard_size))):
            dist.all_reduce(shard, op=dist.ReduceOp.SUM, group=group)
        tensor /= world
    return tensor


def sync_grads(params, gain=1.0, group=None):
    """Average ``p.grad`` of every parameter over ranks (reference semantics incl. gain and nan_to_num)."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.detach().to(torch.float32).flatten() for p in params])
    flat = sharded_all_mean(flat, group=group)
    if gain != 1:
        flat = flat * gain
    torch.nan_to_num(flat, nan=0.0, posinf=1e5, neginf=-1e5, out=flat)
    for p, g in zip(params, flat.split([p.numel() for p in params])):
        p.grad = g.reshape(p.shape).to(p.grad.dtype)


def broadcast_module(module, src=0, group=None):
    """Initial replica sync (training/training_loop.py:616-619)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
