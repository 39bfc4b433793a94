"""Multi-GPU sampling plumbing: one process per GPU, batch sharded over ranks.

The reference never distributes sampling (the only parallelism is Lightning's `ddp`
for training, run_lib.py:55-57). Samples are independent except for the Langevin corrector's
batch-mean norms (sampling/correctors.py:72-74,102-104), so sharding a batch of B over R ranks equals
R independent reference runs of B/R. The data path therefore needs exactly one
collective - a broadcast of the flat parameter buffer at start-up - and none per step; the optional
all-gather below only collects finished samples.

torch.distributed is the plumbing (NCCL on GPUs; the same code runs over gloo on CPU tensors, which is
how tests/test_distributed_cpu.py covers it with world_size 2).
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous slice [lo, hi) of n samples owned by `rank`; the first n % world_size ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_parameters(module, src=0):
    """One broadcast of all parameters and buffers as a single flat fp32 buffer (48.5 M params = 194 MB for
    the 160 px nf96 NCSN++), then scatter back in place. Returns the number of bytes broadcast."""
    rank, world_size = world()
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers() if b.is_floating_point()]
    if world_size == 1 or not tensors:
        return 0
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    eng = getattr(module, "_engine", None)
    if eng is not None:
        eng.invalidate()      # written through .data: autograd's version counters did not move
    return flat.numel() * 4


def sample_sharded(sampler, model, y=None, gather=False, seed=None, **kwargs):
    """Run `sampler(model[, y_shard])` on this rank's contiguous slice of the global condition batch `y`.

    sampler: the function returned by sampling.get_pc_conditional_sampler / get_pc_sampler, built for the
    PER-RANK shape. seed: base seed; rank r draws from seed + r (the reference seeds nothing).
    gather=True all-gathers the finished samples (equal shard sizes required); no collective runs inside the
    sampling loop either way. Returns (samples, info)."""
    rank, world_size = world()
    if seed is not None:
        torch.manual_seed(seed + rank)
    if y is not None:
        lo, hi = shard_range(y.shape[0], rank, world_size)
        samples, info = sampler(model, y[lo:hi], **kwargs)
    else:
        samples, info = sampler(model, **kwargs)
    if gather and world_size > 1:
        parts = [torch.empty_like(samples) for _ in range(world_size)]
        dist.all_gather(parts, samples.contiguous())
        samples = torch.cat(parts, dim=0)
    return samples, info


def max_over_ranks(value, device):
    """Max of a python float over ranks (device-side timing numbers are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


# ------------------------------------------------------------------------------------------------------------
# training: data-parallel gradient all-reduce (the only collective of the backward path)
# ------------------------------------------------------------------------------------------------------------
def _shared_flat(grads):
    """If every gradient is a view into ONE contiguous fp32 buffer, in order (what engine_train.TrainPlan.param_grads
    returns), hand back a flat view over that span so it can be reduced in place with a single collective."""
    if not grads:
        return None
    base = grads[0].untyped_storage()
    end = None
    lo = grads[0].storage_offset()
    for g in grads:
        if g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base.data_ptr():
            return None
        if end is not None and g.storage_offset() < end:
            return None
        end = g.storage_offset() + g.numel()
    flat = torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(base, lo, (end - lo,), (1,))
    return flat


class OverlappedGradSync:
    """Bucketed gradient all-reduce overlapped with the backward pass (what Lightning's accelerator='ddp' gives the
    reference through torch DDP's bucket hooks, run_lib.py:55-57). The engine's backward is a planned launch list over
    ONE flat fp32 gradient buffer; `engine_train.TrainPlan.grad_segments` cuts it into `segments` pieces and knows which
    spans of the buffer are final after each piece. This hook is called after every piece: it all-reduces those spans on
    a side stream (NCCL over NVLink) while the compute stream runs the next piece, and joins the streams after the last
    one, so `loss.backward()` returns averaged gradients. No other collective runs on the backward path."""

    def __init__(self, segments=4, average=True, group=None):
        self.segments, self.average, self.group = max(1, int(segments)), average, group
        self.stream = None
        self.bytes_last = 0
        self.calls_last = 0

    def __call__(self, flat, ranges, last):
        world_size = world()[1]
        if world_size > 1 and ranges:
            if flat.is_cuda:
                if self.stream is None:
                    self.stream = torch.cuda.Stream(device=flat.device)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(flat.device))
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(ev)
                    self._reduce(flat, ranges, world_size)
            else:
                self._reduce(flat, ranges, world_size)
        if last and self.stream is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self.stream)

    def _reduce(self, flat, ranges, world_size):
        for off, n in ranges:
            span = flat[off:off + n]
            dist.all_reduce(span, group=self.group)
            if self.average:
                span.div_(world_size)
            self.bytes_last += n * 4
            self.calls_last += 1


def enable_gradient_overlap(module, segments=4, average=True, group=None):
    """Install an OverlappedGradSync on an engine-backed network: from now on `loss.backward()` returns gradients that
    are already averaged over the ranks, and `allreduce_gradients(module)` becomes a no-op for it. Returns the hook (its
    `bytes_last` / `calls_last` counters accumulate what was reduced). A single-process run is left untouched."""
    eng = getattr(module, "_engine", None)
    if eng is None:
        raise ValueError("enable_gradient_overlap needs an engine-backed network (models.ncsnpp / models.ddpm)")
    if world()[1] == 1:
        eng.grad_sync = None
        return None
    eng.grad_sync = OverlappedGradSync(segments, average, group)
    return eng.grad_sync


def disable_gradient_overlap(module):
    eng = getattr(module, "_engine", None)
    if eng is not None:
        eng.grad_sync = None


def allreduce_gradients(module_or_params, average=True, bucket_bytes=64 << 20, group=None):
    """Sum (or average) the .grad of every parameter over the ranks: the DDP gradient all-reduce the reference gets
    from Lightning's accelerator='ddp' (run_lib.py:55-57). Gradients produced by the engine's backward pass live in
    one flat buffer and are reduced in place with ONE collective over NVLink; otherwise they are packed into buckets of
    `bucket_bytes`. Returns the number of bytes reduced."""
    rank, world_size = world()
    if getattr(getattr(module_or_params, "_engine", None), "grad_sync", None) is not None:
        return 0            # enable_gradient_overlap: the backward pass has already averaged the gradients
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    grads = [p.grad for p in params if p.grad is not None]
    if world_size == 1 or not grads:
        return 0
    flat = _shared_flat(grads)
    if flat is not None:
        dist.all_reduce(flat, group=group)
        if average:
            flat.div_(world_size)
        return flat.numel() * 4
    total, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, total
        if not bucket:
            return
        buf = torch.cat([g.reshape(-1).to(torch.float32) for g in bucket])
        dist.all_reduce(buf, group=group)
        if average:
            buf.div_(world_size)
        off = 0
        for g in bucket:
            g.copy_(buf[off:off + g.numel()].view_as(g))
            off += g.numel()
        total += buf.numel() * 4
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return total
