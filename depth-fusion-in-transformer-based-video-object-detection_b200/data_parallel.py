"""Frame-sharded data parallelism for the deformable-attention models.

The reference's only parallel strategy is DDP: one process per GPU, frames (or whole clips for
TransVOD++) split across ranks, and ONE collective per training step -- the all-reduce of the
trainable-parameter gradients (/root/reference/main.py:440-442, util/misc.py:441-479).  Forward /
inference needs no exchange at all: every MSDeformAttn call is independent per batch element.

This module is that plumbing on ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the
CPU tests):
  * :func:`shard_range` -- which frames / clips a rank owns;
  * :class:`GradientAllReducer` -- bucketed gradient all-reduce (SUM, then / world) that starts a
    bucket's all-reduce as soon as backward has produced all of its gradients, so communication
    overlaps the rest of backward (what DDP's reducer does; ~52 MB of fp32 gradients for the
    Encoder-Cross-Fusion transformer, a fraction of a millisecond on NVLink 5);
  * :class:`GraphedTrainStep` -- the launch-bound forward + backward of a training step captured
    once in a CUDA graph, with all gradients living in one flat buffer per dtype so that the
    step's single collective is one all-reduce of that buffer;
  * :class:`GraphedInference` -- a frame-sharded inference step (no collective) replayed from a CUDA
    graph: static input buffers in, the captured outputs out.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced split of ``n_items`` frames/clips: returns (start, stop) of ``rank``.
    The first ``n_items % world_size`` ranks get one extra item."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class GradientAllReducer:
    """Average gradients across ranks, bucket by bucket, overlapped with backward.

        reducer = GradientAllReducer(model.parameters())
        loss.backward()          # hooks launch async all-reduces as buckets fill
        reducer.finish()         # wait, divide by world size, scatter back into .grad
        optimizer.step()

    Parameters that received no gradient in a step (the reference needs
    ``find_unused_parameters=True``) are treated as zeros so that every rank issues the same
    collectives."""

    def __init__(self, params, bucket_bytes=25 << 20, process_group=None):
        self.group = process_group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        # buckets in reverse order: gradients of the last layers are ready first
        self.buckets, cur, cur_bytes = [], [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): b for b, bucket in enumerate(self.buckets) for p in bucket}
        self._flat = [torch.zeros(sum(p.numel() for p in bucket), dtype=bucket[0].dtype, device=bucket[0].device)
                      for bucket in self.buckets]
        self._pending = [len(bucket) for bucket in self.buckets]
        self._work = [None] * len(self.buckets)
        self._launched = 0            # buckets 0 .. _launched-1 have been sent this step
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _launch(self, b):
        flat, offset = self._flat[b], 0
        for p in self.buckets[b]:
            n = p.numel()
            if p.grad is not None:
                flat[offset:offset + n].copy_(p.grad.reshape(-1))
            else:
                flat[offset:offset + n].zero_()
            offset += n
        if self.world > 1:
            self._work[b] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, param):
        b = self._bucket_of[id(param)]
        if self._pending[b] <= 0:
            raise RuntimeError("GradientAllReducer: a second backward() reached a parameter before finish(); "
                               "call finish() once per backward pass")
        self._pending[b] -= 1
        # Collectives must be issued in the SAME order on every rank (the set of unused parameters, and with it the
        # order in which buckets complete, may differ between ranks): strictly by bucket index, a bucket only once
        # every earlier bucket has been sent -- what DDP's reducer does.
        while self._launched < len(self.buckets) and self._pending[self._launched] == 0:
            self._launch(self._launched)
            self._launched += 1

    def finish(self):
        """Complete the step: the remaining buckets (those waiting for an earlier one, and those whose hooks never
        all fired because of unused parameters) are sent now, in index order."""
        while self._launched < len(self.buckets):
            self._launch(self._launched)
            self._launched += 1
        for b, bucket in enumerate(self.buckets):
            if self._work[b] is not None:
                self._work[b].wait()
                self._work[b] = None
            flat, offset = self._flat[b], 0
            if self.world > 1:
                flat.div_(self.world)
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[offset:offset + n].view_as(p))
                offset += n
            self._pending[b] = len(bucket)
        self._launched = 0

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def gradient_bytes(self):
        return sum(f.numel() * f.element_size() for f in self._flat)


class GraphedTrainStep:
    """forward + loss + backward captured in ONE CUDA graph; a step is
        replay -> all-reduce of the flat gradient buffer(s) (world > 1) -> optimizer.step().

    The Encoder-Cross-Fusion training step issues ~3000 kernels; launched eagerly it is bound by
    the host (42 ms of CPU time against 27 ms of GPU time on the B200 box), replayed it is bound by
    the GPU.  Gradients are views into one flat buffer per dtype (autograd accumulates into them in
    place), so no per-parameter copies surround the collective.

        step = GraphedTrainStep(model, optimizer, lambda: loss_of(model(*static_inputs)))
        loss = step()            # static_inputs may be overwritten in place between calls

    ``loss_fn`` must be shape-static and sync-free (run it eagerly a few times first: the shape
    caches of the deformable-attention modules read ``spatial_shapes`` back once)."""

    def __init__(self, model, optimizer, loss_fn, process_group=None, warmup=3):
        self.model, self.optimizer, self.loss_fn, self.group = model, optimizer, loss_fn, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in model.parameters() if p.requires_grad]
        self._flat = {}
        by_dtype = {}
        for p in params:
            by_dtype.setdefault(p.dtype, []).append(p)
        for dtype, group in by_dtype.items():
            flat = torch.zeros(sum(p.numel() for p in group), dtype=dtype, device=group[0].device)
            offset = 0
            for p in group:
                p.grad = flat[offset:offset + p.numel()].view_as(p)
                offset += p.numel()
            self._flat[dtype] = flat
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero()
                self.loss_fn().backward()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._zero()
            self.loss = self.loss_fn()
            self.loss.backward()
        self._zero()

    def _zero(self):
        for flat in self._flat.values():
            flat.zero_()

    @property
    def gradient_bytes(self):
        return sum(f.numel() * f.element_size() for f in self._flat.values())

    def __call__(self):
        self.graph.replay()
        if self.world > 1:
            for flat in self._flat.values():
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
        self.optimizer.step()
        return self.loss


def _flatten_tensors(obj, out):
    """Depth-first list of the tensors inside nested lists / tuples / dicts (other leaves are ignored)."""
    if torch.is_tensor(obj):
        out.append(obj)
    elif isinstance(obj, (list, tuple)):
        for item in obj:
            _flatten_tensors(item, out)
    elif isinstance(obj, dict):
        for item in obj.values():
            _flatten_tensors(item, out)
    return out


class GraphedInference:
    """An inference call replayed from ONE CUDA graph.  The transformers here issue 300-700 kernels per step, most
    of them a few microseconds long: launched eagerly the step is host-bound (e.g. the TransVOD++ clip transformer
    12.8 ms eager vs 6.4 ms replayed on the B200 box).

        run = GraphedInference(lambda: model(srcs, masks, pos, None, None, None, query), inputs=(srcs, masks, pos))
        out = run(new_srcs, new_masks, new_pos)     # copied into the static buffers, then one graph launch
        out = run()                                 # or overwrite the static tensors in place yourself

    ``fn`` must be shape-static and sync-free; it is run ``warmup`` times eagerly first (the shape caches of the
    deformable-attention modules read ``spatial_shapes`` back once).  ``inputs`` is the (nested) structure of
    static tensors ``fn`` closes over; ``__call__`` takes the same structure.  The returned tensors are the
    graph's output buffers: they are overwritten by the next call."""

    def __init__(self, fn, inputs=(), warmup=2):
        self._static = _flatten_tensors(inputs, [])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.outputs = fn()

    def __call__(self, *inputs):
        if inputs:
            fresh = _flatten_tensors(inputs if len(inputs) != 1 else inputs[0], [])
            if len(fresh) != len(self._static):
                raise ValueError(f"expected {len(self._static)} input tensors, got {len(fresh)}")
            for dst, src in zip(self._static, fresh):
                if dst.shape != src.shape or dst.dtype != src.dtype:
                    raise ValueError(f"static input of shape {tuple(dst.shape)} {dst.dtype} cannot take "
                                     f"{tuple(src.shape)} {src.dtype}")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
