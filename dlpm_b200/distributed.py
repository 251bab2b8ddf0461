"""Batch-sharded sampling over the GPUs of one node.

Every sample's reverse chain is independent (GroupNorm and attention are per sample), so the batch is cut into
contiguous slices, one per rank; weights and schedules are replicated; the Philox counters are keyed by the GLOBAL
sample index (``rng.set_sample_base``), so any world size reproduces the single-GPU samples; the only collective is one
``all_gather`` of the final samples (NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no
distributed execution on this path ("implement DDP later on", ``dlpm/dlpm_experiment.py:91``); its evaluation loop
chunks generation by ``eval.batch_size`` instead (``bem/evaluate/EvaluationManager.py:181-193``).
"""
import torch
import torch.distributed as dist

from . import rng


def shard_bounds(total, world, rank):
    """Contiguous slice [start, start + count) of ``total`` samples owned by ``rank`` (first ranks take the remainder)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def init_shard(total, rank=None, world=None, state=None):
    """Set this process's Philox sample base to the start of its slice; returns (start, count)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    start, count = shard_bounds(total, world, rank)
    (state or rng.default_state()).sample_base = start
    return start, count


def gather_samples(x, total, group=None):
    """All-gather per-rank slices (possibly uneven) into the full (total, ...) tensor, in global sample order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    counts = [shard_bounds(total, world, r)[1] for r in range(world)]
    cmax = max(counts)
    if all(c == cmax for c in counts):
        out = torch.empty((total,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out
    pad = torch.zeros((cmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]].copy_(x)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def sample_sharded(method, models, shape, group=None, **sample_kwargs):
    """``method.sample`` for a GLOBAL batch ``shape[0]`` spread over the process group; every rank returns all samples."""
    total = int(shape[0])
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, count = shard_bounds(total, world, rank)
    state = rng.default_state()
    old = state.sample_base
    state.sample_base = start
    hist = None
    want_hist = bool(sample_kwargs.get("get_sample_history", False))
    tail = [int(s) for s in shape[1:]]
    if count == 0:
        # more ranks than samples: this rank owns nothing (the kernels reject empty batches) but still takes part in the
        # gather with an empty slice of the right trailing shape / dtype / device
        dev = getattr(method, "device", "cpu")
        local = torch.empty([0] + tail, dtype=torch.float32, device=dev)
    else:
        try:
            local = method.sample(models, [count] + tail, **sample_kwargs)
        finally:
            state.sample_base = old
    state.sample_base = old
    if isinstance(local, tuple):
        local, hist = local
    if want_hist and world > 1:
        # the number of history entries is the sampler's business (T for DLPM / DLIM, steps + 1 for LIM): agree on it so
        # that a rank without samples contributes an empty history of the same length
        n_hist = torch.tensor([0 if hist is None else hist.shape[0]], dtype=torch.int64, device=local.device)
        dist.all_reduce(n_hist, op=dist.ReduceOp.MAX, group=group)
        if hist is None:
            hist = torch.empty([int(n_hist[0]), 0] + tail, dtype=local.dtype, device=local.device)
    out = gather_samples(local, total, group)
    if hist is not None:
        hist = gather_samples(hist.transpose(0, 1).contiguous(), total, group).transpose(0, 1).contiguous()
        return out, hist
    return out
