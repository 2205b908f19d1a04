"""Batch-axis data parallelism over the GPUs of one NVSwitch box.

Every sample's correlation volume, lookups and GRU state are independent (corr.py:156-160 is a per-sample
batched matmul), so inference shards the batch with NO collective on the data path: one process per GPU,
each runs its contiguous slice.  `torch.distributed` is only used for the optional gather of results and
for timing barriers.  (The training-step gradient all-reduce belongs to SURVEY.md section 8(f) rank 2.)
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world_size):
    """Contiguous [lo, hi) slice of the batch owned by `rank`; remainders go to the lowest ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank/world_size: {rank}/{world_size}')
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every (B, ...) tensor of `tensors` to this rank's part of the batch."""
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world_size)
    return [t[lo:hi] for t in tensors]


def gather_batch(local, global_batch, group=None):
    """All-gather per-rank (b_r, ...) results back into a (B, ...) tensor in batch order (uneven shards ok)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    max_b = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def predict_sharded(predict_fn, image1, image2, group=None):
    """Run `predict_fn(image1_shard, image2_shard) -> (b_r, H, W, 2)` on this rank's shard of a replicated
    global batch and return the gathered (B, H, W, 2) result on every rank."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    s1, s2 = shard_batch([image1, image2], rank, world)
    out = predict_fn(s1, s2)
    return gather_batch(out, image1.shape[0], group)


def max_over_ranks(value, device, group=None):
    """max of a python float over all ranks (timing: a multi-GPU step takes as long as its slowest rank)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def predict_stream(predict_fn, batches, device, reuse_host_buffers=False):
    """Pipelined inference over host batches: yields `predict_fn(image1, image2)` for every `(image1, image2)` pair of
    `batches` as host tensors, in order.

    `reuse_host_buffers=True` reads the results back into a ring of three pinned buffers instead of allocating a pinned
    tensor per step (cudaHostAlloc costs as much as a forward): a yielded tensor is then valid until two further results
    have been consumed.

    On a CUDA device the upload of pair i+1 is issued on a copy stream while pair i computes on the current stream, and
    the result of pair i is read back on a second copy stream while pair i+1 computes (the result is first moved to one
    of two device staging buffers -- a few microseconds -- because `predict_fn` may return a static CUDA-graph buffer
    that the next call overwrites); pass pinned host tensors for the uploads to be asynchronous.  Every pair still makes the full host -> device -> host round trip; only
    the waiting is overlapped.  (`device` of type 'cpu' runs the same schedule synchronously: plumbing tests.)"""
    device = torch.device(device)
    cuda = device.type == 'cuda'
    main = torch.cuda.current_stream(device) if cuda else None
    copy = torch.cuda.Stream(device=device) if cuda else None
    down = torch.cuda.Stream(device=device) if cuda else None
    staging, staged = [None, None], [None, None]            # device staging buffers and the events of their last read-back
    step = 0

    # Device input buffers are allocated once (two sets, used alternately): a fresh `.to(device)` per step makes the caching
    # allocator grow under two streams (cudaMalloc synchronises) -- measured 3x slower end to end at 448x1024.
    in_bufs, in_free, up_step = [None, None], [None, None], 0

    def upload(pair):
        nonlocal up_step
        if not cuda:
            return pair[0], pair[1], None, None
        k = up_step % 2
        up_step += 1
        bufs = in_bufs[k]
        if bufs is None or any(t.shape != s.shape or t.dtype != s.dtype for t, s in zip(bufs, pair[:2])):
            bufs = in_bufs[k] = tuple(torch.empty(s.shape, dtype=s.dtype, device=device) for s in pair[:2])
            in_free[k] = None
        with torch.cuda.stream(copy):
            if in_free[k] is not None:
                copy.wait_event(in_free[k])                            # the compute that read this set has finished
            bufs[0].copy_(pair[0], non_blocking=True)
            bufs[1].copy_(pair[1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        return bufs[0], bufs[1], ready, k

    ring, ring_pos = [], 0

    def host_buffer(like):
        nonlocal ring_pos
        if not reuse_host_buffers:
            return torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
        slot = ring_pos % 3
        ring_pos += 1
        if slot >= len(ring):
            ring.append(None)                      # slots fill in order 0, 1, 2
        buf = ring[slot]
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = ring[slot] = torch.empty(like.shape, dtype=like.dtype, pin_memory=torch.cuda.is_available())
        return buf

    it = iter(batches)
    first = next(it, None)
    if first is None:
        return
    cur, pending = upload(first), None
    while cur is not None:
        nxt_pair = next(it, None)
        nxt = upload(nxt_pair) if nxt_pair is not None else None     # overlaps the compute issued below
        a, b, ready, in_k = cur
        if cuda:
            main.wait_event(ready)
        out = predict_fn(a, b)
        if cuda:
            consumed = torch.cuda.Event()
            consumed.record(main)
            in_free[in_k] = consumed                                  # the input set may be overwritten after this point
            k = step % 2
            step += 1
            if staging[k] is None or staging[k].shape != out.shape or staging[k].dtype != out.dtype:
                staging[k] = torch.empty_like(out)
            if staged[k] is not None:
                main.wait_event(staged[k])                            # the read-back two steps ago has left this buffer
            staging[k].copy_(out, non_blocking=True)                  # device -> device on the compute stream
            computed = torch.cuda.Event()
            computed.record(main)
            host_out = host_buffer(out)
            with torch.cuda.stream(down):                             # device -> host beside the next step's compute
                down.wait_event(computed)
                host_out.copy_(staging[k], non_blocking=True)
                done = torch.cuda.Event()
                done.record(down)
            staged[k] = done
        else:
            host_out, done = out, None
        if pending is not None:
            if pending[1] is not None:
                pending[1].synchronize()
            yield pending[0]
        pending = (host_out, done)
        cur = nxt
    if pending[1] is not None:
        pending[1].synchronize()
    yield pending[0]
