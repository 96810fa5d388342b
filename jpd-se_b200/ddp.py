"""Data-parallel training plumbing: gradient all-reduce overlapped with the generator backward.

The reference trains single-process (ctu/trainers/pix2pixHD_trainer.py:42-85). All its losses are batch means and
InstanceNorm is per-sample (networks.py:31), so the mean over ranks of per-rank gradients equals the global-batch
gradient: one exchange step per optimizer step.

``GradReducer`` plugs into ``GlobalGenerator.grad_reducer``. The generator backward is OUR kernel sequence, so the
hook points are exact: the plan asks the reducer for every gradient tensor (``alloc`` hands out consecutive slices of
one flat fp32 buffer, in reverse layer order) and reports it ``ready`` right after its wgrad kernels are enqueued.
Whenever ``bucket_bytes`` of contiguous gradients are ready the reducer records an event on the compute stream and
launches ``all_reduce(AVG)`` of that slice on a side stream -- no packing copies, and the transfer over NVLink runs
under the remaining dgrad / wgrad kernels. ``finish`` makes the compute stream wait for the outstanding reductions.

The discriminator's gradients (22 MB, produced by stock autograd) are reduced in one flat call after
``loss_D.backward()`` (``allreduce_grads``); the netD gradients that ``loss_G.backward()`` leaves behind are thrown
away by ``optimizer_D.zero_grad()`` (pix2pixHD_trainer.py:73) and never reduced.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class GradReducer:
    def __init__(self, group=None, bucket_bytes=64 << 20, tail_bytes=32 << 20, tail_bucket_bytes=2 << 20):
        """bucket_bytes: size at which a run of ready gradients is handed to all_reduce. The LAST `tail_bytes` of the flat
        buffer (the gradients the backward produces last: the downsampling convs and the stem) go out in buckets of
        `tail_bucket_bytes` instead, so the reduction that is still in flight when the backward ends -- the exposed part --
        is a few MB, not a 64 MB bucket."""
        self.group = group
        import os
        if os.environ.get("JPDSE_BUCKET_MB"):  # A/B measurements
            bucket_bytes = int(float(os.environ["JPDSE_BUCKET_MB"]) * (1 << 20))
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.tail_elems = max(0, tail_bytes // 4)
        self.tail_bucket_elems = max(1, tail_bucket_bytes // 4)
        # JPDSE_GRAD_BF16=1 (opt-in): buckets cross the wire as bf16 -- fp32 -> bf16 on the communication stream,
        # all_reduce(AVG) of half the bytes, bf16 -> fp32 back into the flat buffer. Gradients then carry one bf16
        # rounding (2^-9 relative) before the average; off by default so that the N-rank step equals the reference's
        # global-batch step to fp32 summation order.
        self.bf16 = os.environ.get("JPDSE_GRAD_BF16", "0") == "1"
        self._half = None
        self.total = 0
        self.flat = None
        self.stream = None
        self.copy_out = False  # set by begin(): gradients must leave as copies (an earlier .grad exists)
        self.reset_stats()

    def reset_stats(self):
        self.launched = []   # (start, end) element ranges handed to all_reduce, in launch order
        self._works = []
        self._offset = 0
        self._bucket_start = 0

    # ---- protocol used by _GeneratorFunction.backward / GeneratorPlan.backward
    def begin(self, named_params):
        total = sum((p.numel() + 3) // 4 * 4 for _, p in named_params if p.requires_grad)
        self.total = total
        ref = next(p for _, p in named_params)
        # Gradient accumulation: autograd's AccumulateGrad keeps the tensors `alloc` hands out WITHOUT copying, so after
        # one backward every p.grad is a view of `flat`. A second backward without zero_grad(set_to_none=True) would
        # overwrite those views in place and then add the same memory to itself (2*g2 instead of g1+g2). When a
        # gradient is already present, detach it from `flat` first and hand this backward's gradients out as copies.
        self.copy_out = False
        if self.flat is not None:
            lo = self.flat.data_ptr()
            hi = lo + self.flat.numel() * 4
            for _, p in named_params:
                if p.grad is not None:
                    self.copy_out = True
                    if lo <= p.grad.data_ptr() < hi:
                        p.grad = p.grad.clone()
        elif any(p.grad is not None for _, p in named_params):
            self.copy_out = True
        if self.flat is None or self.flat.numel() < total or self.flat.device != ref.device:
            self.flat = torch.empty(total, dtype=torch.float32, device=ref.device)
        if ref.is_cuda and self.stream is None:
            self.stream = torch.cuda.Stream(device=ref.device)
        if self.bf16 and ref.is_cuda and (self._half is None or self._half.numel() < total):
            self._half = torch.empty(total, dtype=torch.bfloat16, device=ref.device)
        self.reset_stats()

    def alloc(self, key, shape):
        n = 1
        for s in shape:
            n *= s
        t = self.flat[self._offset: self._offset + n].view(shape)
        self._offset += (n + 3) // 4 * 4  # keep every tensor 16-byte aligned
        return t

    def ready(self, key, tensor):
        in_tail = self.total - self._offset < self.tail_elems
        if self._offset - self._bucket_start >= (self.tail_bucket_elems if in_tail else self.bucket_elems):
            self._launch()

    def finish(self):
        self._launch()
        for w in self._works:
            if isinstance(w, torch.cuda.Event):
                torch.cuda.current_stream(self.flat.device).wait_event(w)  # bf16 buckets: converted back on the side stream
            else:
                w.wait()  # CUDA: the current (compute) stream waits; gloo: blocks
        self._works = []
        if _world(self.group) > 1 and not self.flat.is_cuda:
            self.flat[: self._offset].div_(_world(self.group))

    # ---- internals
    def _launch(self):
        a, b = self._bucket_start, self._offset
        if b <= a:
            return
        self._bucket_start = b
        self.launched.append((a, b))
        if _world(self.group) == 1:
            return
        chunk = self.flat[a:b]
        if chunk.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(chunk.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                if self.bf16 and self._half is not None:
                    half = self._half[a:b]
                    half.copy_(chunk)
                    dist.all_reduce(half, op=dist.ReduceOp.AVG, group=self.group, async_op=True).wait()  # stream-ordered
                    chunk.copy_(half)
                    done = torch.cuda.Event()
                    done.record(self.stream)
                    self._works.append(done)
                else:
                    self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))


def allreduce_grads(params, group=None):
    """Average the .grad of `params` over the ranks with one flat all-reduce (the discriminator's gradients)."""
    n = _world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if n == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(n)
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights (the reference initialises on-device, networks.py:52-55)."""
    if _world(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
