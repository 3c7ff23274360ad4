"""Data-parallel training step of the hot path: one process per GPU, identical replicas, ONE collective.

Replaces ``torch.nn.DataParallel(model)`` of the reference (train.py:97).  Windows never cross images, so the path
shards by batch with no activation exchange: the only traffic is the gradient all-reduce (average)
over NCCL / NVLink.  All gradients live in ONE flat fp32 buffer (``optim.FlatParams``, shared with ``FusedAdamW``), so
the exchange is a handful of large all-reduces instead of one per parameter:

* ``reduce_gradients()`` -- after ``loss.backward()``: packs the gradients autograd assigned into the flat buffer (one
  multi-tensor copy) and all-reduces it in ``buckets`` contiguous slices on a side stream.  This is the call to use
  after replaying a CUDA graph of forward + backward.
* ``overlap=True`` (eager backward) -- per-parameter post-accumulate hooks copy each gradient into its flat slot as soon
  as autograd has produced it and launch a bucket's all-reduce when its last gradient has arrived; the buckets are laid
  out in reverse parameter order, i.e. in the order the backward produces them, so the exchange of the deep layers
  runs under the backward of the shallow ones.

Semantics versus the reference's ``DataParallel`` (documented, not "fixed"): each rank computes its own
SiLog over its own batch and the gradients are averaged; parameters are broadcast once at construction, not every step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .optim import FlatParams


class DataParallel(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, process_group=None, buckets: int = 4, overlap: bool = False,
                 flat: FlatParams | None = None, broadcast: bool = True, bf16_copies: bool = True):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.flat = flat if flat is not None else FlatParams(params, bf16_copies=bf16_copies)
        self.overlap = bool(overlap) and self.world > 1
        self._cuda = self.flat.data.is_cuda
        self._side = torch.cuda.Stream(device=self.flat.data.device) if (self.world > 1 and self._cuda) else None
        # contiguous buckets over the flat buffer, cut at tensor boundaries; bucket 0 holds the LAST parameters
        self._bounds = self._make_buckets(max(1, int(buckets)))
        if broadcast and self.world > 1:
            dist.broadcast(self.flat.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                           group=process_group)
            if self.flat.bf16 is not None:
                self.flat.bf16.copy_(self.flat.data)
        self._pending = None
        self._hooks = []
        if self.overlap:
            self._install_hooks()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ------------------------------------------------------------------ bucket layout
    def _make_buckets(self, n):
        f = self.flat
        ends = [o + ((p.numel() + f.chunk - 1) // f.chunk) * f.chunk for o, p in zip(f.offsets, f.params)]
        target = f.total / n
        bounds, lo = [], 0
        first_param = 0
        for i, e in enumerate(ends):
            if e - lo >= target or i == len(ends) - 1:
                bounds.append((lo, e, first_param, i + 1))      # [lo, e) elements, parameters [first, i + 1)
                lo, first_param = e, i + 1
        return bounds[::-1]                                      # reverse: the order the backward fills them

    # ------------------------------------------------------------------ after-backward exchange
    def reduce_gradients(self, async_op: bool = False):
        """Pack + all-reduce (average).  Afterwards every ``p.grad`` is its flat view holding the averaged gradient and
        ``FusedAdamW(flat=...)`` can step without another copy (``grads_packed``)."""
        f = self.flat
        if not self.overlap:
            f.pack_grads()
        if self.world > 1:
            if self.overlap:
                self._finish_overlap()
            else:
                for lo, hi, _, _ in self._bounds:
                    self._launch_bucket(lo, hi)
                self._join(async_op)
        f.point_grads_at_flat()

    def wait(self):
        if self._pending is not None:
            torch.cuda.current_stream(self.flat.data.device).wait_event(self._pending)
            self._pending = None

    def _all_reduce_avg(self, t):
        if self._cuda:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def _launch_bucket(self, lo, hi):
        """All-reduce one slice of the flat gradient buffer on the side stream, after everything queued so far."""
        g = self.flat.grad[lo:hi]
        if not self._cuda:
            self._all_reduce_avg(g)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(g.device))
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._all_reduce_avg(g)

    def _join(self, async_op=False):
        if not self._cuda:
            return
        done = torch.cuda.Event()
        done.record(self._side)
        if async_op:
            self._pending = done
        else:
            torch.cuda.current_stream(self.flat.data.device).wait_event(done)

    # ------------------------------------------------------------------ overlapped exchange (eager backward)
    def _install_hooks(self):
        f = self.flat
        self._bucket_of = {}
        for b, (_, _, p0, p1) in enumerate(self._bounds):
            for i in range(p0, p1):
                self._bucket_of[i] = b
        self._left = [p1 - p0 for _, _, p0, p1 in self._bounds]

        def make(i):
            def hook(p):
                f.grad_views[i].copy_(p.grad)
                b = self._bucket_of[i]
                self._left[b] -= 1
                if self._left[b] == 0:
                    lo, hi, _, _ = self._bounds[b]
                    self._launch_bucket(lo, hi)
            return hook

        for i, p in enumerate(f.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(make(i)))

    def _finish_overlap(self):
        f = self.flat
        # parameters that received no gradient this step: their buckets still have to go out (zeros)
        for b, left in enumerate(self._left):
            if left:
                lo, hi, p0, p1 = self._bounds[b]
                for i in range(p0, p1):
                    if f.params[i].grad is None:
                        f.grad_views[i].zero_()
                self._launch_bucket(lo, hi)
        self._join()
        self._left = [p1 - p0 for _, _, p0, p1 in self._bounds]
