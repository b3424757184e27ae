"""Gradient exchange for the batch-sharded multi-GPU training step.

The decoder hot path shards by image batch (every op is per-sample; weights are replicated), so forward / decode needs no
communication.  Training has exactly one exchange step: the gradient mean over ranks.  The reference does it after the whole
backward (training/training_loop.py:272-289 ``sync_grads`` -> ``sharded_all_mean``, called at :726):

    flat fp32 concat of all grads -> all_reduce(SUM) in shards of <= 2**23 elements -> / world_size -> * gain
    -> nan_to_num(nan=0, posinf=1e5, neginf=-1e5) -> split back, one reshape + cast per parameter

Two implementations of those semantics live here:

* ``sync_grads`` / ``sharded_all_mean``: the functional mirror (post-hoc, unoverlapped), kept as the equality baseline.
* ``GradExchange``: the B200 path.  ONE persistent flat fp32 buffer holds every gradient (``p.grad`` are views into it, so
  there is no concat, no split and no per-parameter cast loop); the buffer is cut into buckets of <= 2**23 elements in
  *reverse* parameter order, i.e. the order in which backward finishes them; a post-accumulate-grad hook per parameter counts a
  bucket's gradients in and launches its ``all_reduce`` asynchronously (NCCL's own stream, ordered after the backward kernels
  that produced the bucket) while backward keeps running on the compute stream; ``finish()`` waits for the buckets, then
  applies ``/world``, ``*gain`` and ``nan_to_num`` in ONE in-place kernel pass (``vfm_grad_finalize``, csrc/grad_sync.cu).
  Elementwise the result is the reference's: an all-reduce SUM does not depend on how the flat buffer is sharded.

``torch.distributed`` (NCCL over NVLink / NVSwitch on the box, gloo in the CPU tests) is the plumbing; the bias / weight / noise
gradient reductions *inside* a rank are done by the kernels (bias_act db, modconv wgrad) before they join this exchange.
"""
import ctypes as C

import torch
import torch.distributed as dist

SHARD_ELEMS = 2 ** 23   # 32 MiB of fp32 per all_reduce, like the reference


def _world(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def sharded_all_mean(tensor, shard_size=SHARD_ELEMS, group=None):
    """In-place mean over ranks of a flat tensor, reduced in fixed-size shards."""
    assert tensor.dim() == 1
    world = _world(group)
    if world > 1:
        for shard in tensor.tensor_split(max(1, -(-tensor.numel() // shard_size))):
            dist.all_reduce(shard, op=dist.ReduceOp.SUM, group=group)
        tensor /= world
    return tensor


def sync_grads(params, gain=None, group=None):
    """Average ``p.grad`` of every parameter over ranks (reference semantics incl. gain and nan_to_num), post hoc."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.detach().to(torch.float32).flatten() for p in params])
    flat = sharded_all_mean(flat, group=group)
    if gain is not None and gain != 1:
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


def finalize_(flat, world, gain=None, nan=0.0, posinf=1e5, neginf=-1e5):
    """In place: ``flat = nan_to_num(flat / world [* gain])``.  CUDA buffers go through the library kernel (no fallback);
    CPU buffers (the gloo host-logic tests) through the same arithmetic in torch."""
    if flat.numel() == 0:
        return flat
    if flat.is_cuda:
        from . import _lib
        assert flat.dtype == torch.float32 and flat.is_contiguous()
        p = _lib.GradFinalizeParams()
        p.grads, p.numel, p.world_size = flat.data_ptr(), flat.numel(), int(world)
        p.use_gain, p.gain = int(gain is not None), float(gain if gain is not None else 1.0)
        p.nan, p.posinf, p.neginf = float(nan), float(posinf), float(neginf)
        with torch.cuda.device(flat.device):
            _lib.check(_lib.load().vfm_grad_finalize(C.byref(p), C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)), 'grad_finalize')
        return flat
    if world > 1:
        flat /= world
    if gain is not None:
        flat *= gain
    return torch.nan_to_num(flat, nan=nan, posinf=posinf, neginf=neginf, out=flat)


class GradExchange:
    """Overlapped, bucketed gradient mean over ranks with the reference's ``sync_grads`` result.

        ex = GradExchange(net.parameters())        # once; installs the flat buffer and the hooks
        loop:
            ex.zero_grad()                          # one memset of the flat buffer (instead of zero_grad(set_to_none=True))
            loss.backward()                         # buckets are all-reduced while backward runs
            ex.finish(gain=n_batch_acc)             # wait + ONE finalize pass; p.grad now hold the averaged gradients
            opt.step()

    Parameters that received no gradient in a backward pass get ``p.grad = None`` for that step, exactly like the reference
    (its ``sync_grads`` skips them and Adam then leaves them untouched); their slots still travel in the bucket as zeros.
    fp32 parameters only (everything in the decoder); others raise.
    """

    def __init__(self, params, bucket_elems=SHARD_ELEMS, group=None, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'GradExchange: no trainable parameters'
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError('GradExchange: fp32 parameters only (the reference concatenates to one fp32 vector)')
        self.group, self.overlap = group, overlap
        dev = self.params[0].device
        # flat layout = reverse parameter order (the order backward produces gradients in), every slot 16-byte aligned
        order = list(reversed(range(len(self.params))))
        self.offsets = [0] * len(self.params)
        off = 0
        for i in order:
            self.offsets[i] = off
            off += (self.params[i].numel() + 3) & ~3
        self.numel = off
        self.flat = torch.zeros([off], dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]
        # buckets: consecutive runs of whole parameters, closed once they hold >= bucket_elems (a parameter larger than that is
        # its own bucket); boundaries are identical on every rank because they depend on the parameter shapes only
        self.buckets = []          # (start, end, [param indices])
        start, members = 0, []
        for i in order:
            members.append(i)
            end = self.offsets[i] + ((self.params[i].numel() + 3) & ~3)
            if end - start >= bucket_elems:
                self.buckets.append((start, end, members))
                start, members = end, []
        if members:
            self.buckets.append((start, self.numel, members))
        self.bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self.bucket_of[i] = b
        self._pending = [len(m) for _, _, m in self.buckets]
        self._fired = [False] * len(self.params)      # received a gradient since zero_grad()
        self._counted = [False] * len(self.params)    # ... and was counted into its bucket (not under no_sync())
        self._next = 0             # buckets are launched strictly in index order so that every rank issues the same sequence
        self._works = []
        self._armed = False
        self._accumulating = False
        self._known_unused = set()  # parameters that got no gradient in the previous step: pre-counted so they do not hold their bucket back
        self.stats = dict(buckets=len(self.buckets), numel=self.numel, bytes=self.numel * 4, launched_in_backward=0)
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.zero_grad()

    def _make_hook(self, i):
        def hook(p):
            if not self._armed:
                return
            if p.grad is not None and p.grad.data_ptr() != self.views[i].data_ptr():
                # something replaced p.grad (e.g. zero_grad(set_to_none=True) in between): fold it back into the flat buffer
                self.views[i].copy_(p.grad)
                p.grad = self.views[i]
            self._fired[i] = True
            if i in self._known_unused and self.bucket_of[i] < self._next:
                raise RuntimeError('GradExchange: a parameter that received no gradient in the previous step received one now, after its '
                                   'bucket had been sent; call reset_unused() when the set of trained parameters changes')
            if self._accumulating or self._counted[i]:
                return
            self._counted[i] = True
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self.overlap:
                self._launch_ready(in_backward=True)
        return hook

    def _launch_ready(self, in_backward, force=False):
        world = _world(self.group)
        while self._next < len(self.buckets) and (force or self._pending[self._next] == 0):
            s, e, _ = self.buckets[self._next]
            if world > 1:
                self._works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                if in_backward:
                    self.stats['launched_in_backward'] += 1
            self._next += 1

    def zero_grad(self):
        """Zero the flat buffer and (re)attach the views as ``p.grad``; arms the hooks for the next backward."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._pending = [len(m) for _, _, m in self.buckets]
        self._fired = [False] * len(self.params)
        self._counted = [False] * len(self.params)
        for i in self._known_unused:
            self._counted[i] = True
            self._pending[self.bucket_of[i]] -= 1
        self._next = 0
        self._works = []
        self._armed = True
        self.stats['launched_in_backward'] = 0

    def no_sync(self):
        """Context manager for all but the last micro-batch of an accumulated step (the reference accumulates ``n_batch_acc``
        backward passes before one ``sync_grads(gain=n_batch_acc)``): gradients accumulate in the flat buffer, nothing is sent."""
        ex = self

        class _NoSync:
            def __enter__(self):
                ex._accumulating = True

            def __exit__(self, *a):
                ex._accumulating = False
        return _NoSync()

    def finish(self, gain=None):
        """Flush the remaining buckets, wait for the exchange and finalize in one pass (reference: / world, * gain, nan_to_num)."""
        assert self._armed, 'GradExchange.finish() without a zero_grad() / backward before it'
        self._launch_ready(in_backward=False, force=True)
        for w in self._works:
            w.wait()               # NCCL: makes the current stream wait for the collective; does not block the host
        self._works = []
        finalize_(self.flat, _world(self.group), gain)
        for p, fired in zip(self.params, self._fired):
            if not fired:
                p.grad = None
        self._known_unused = {i for i, fired in enumerate(self._fired) if not fired}
        self._armed = False

    def reset_unused(self):
        self._known_unused = set()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
