"""Multi-GPU plumbing for the matcher: one process per GPU (torchrun), torch.distributed for the
collectives, the CUDA library for everything else.

The path shards by events: every query is independent given (template offset, length, first lag,
lag count), so each rank takes a contiguous slice of the (time-sorted) event list -- neighbouring
events search overlapping parts of the destination stream, which keeps a rank's block spectra hot
in its L2.  Two collectives, both outside the kernels:
    broadcast   the normalised streams from rank 0, once per pair of streams   (65 MB u8 / 90 min)
    all_gather  the per-event (diff, idx) results                              (12 B per event)
There is no collective on the data path of a query, hence no fused compute+communication kernel.

On CPU-only hosts the same functions run over the `gloo` backend with NumPy-backed tensors and a
caller-supplied matcher (tests/test_parallel_cpu.py); on GPUs the backend is `nccl` and the tensors
are the buffers the CUDA library reads and writes directly.
"""
import numpy as np


def shard_bounds(count, world_size, rank):
    """Contiguous, balanced [lo, hi) slice of `count` items for `rank`; the first count % world
    ranks get one extra item."""
    base, extra = divmod(int(count), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(count, world_size):
    return [shard_bounds(count, world_size, r)[1] - shard_bounds(count, world_size, r)[0] for r in range(world_size)]


def broadcast_stream(dist, tensor, src=0):
    """Rank `src` holds the normalised stream; afterwards every rank does (in place)."""
    dist.broadcast(tensor, src)
    return tensor


def all_gather_results(dist, torch, local_diff, local_idx, count, world_size):
    """Gather variable-length per-rank results into full arrays ordered like the event list.
    local_* are 1-D tensors of this rank's shard length (any device); returns (diff, idx) tensors of
    length `count` on the same device."""
    sizes = shard_sizes(count, world_size)
    cap = max(sizes) if sizes else 0
    dev = local_diff.device
    pad_d = torch.zeros(cap, dtype=torch.float32, device=dev)
    pad_i = torch.zeros(cap, dtype=torch.int64, device=dev)
    pad_d[:local_diff.numel()] = local_diff
    pad_i[:local_idx.numel()] = local_idx
    all_d = torch.empty(world_size * cap, dtype=torch.float32, device=dev)
    all_i = torch.empty(world_size * cap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_d, pad_d)
    dist.all_gather_into_tensor(all_i, pad_i)
    keep = torch.cat([torch.arange(r * cap, r * cap + sizes[r], device=dev) for r in range(world_size)]) if cap else \
        torch.zeros(0, dtype=torch.int64, device=dev)
    return all_d[keep], all_i[keep]


def sharded_find(dist, torch, rank, world_size, match_fn, toff, tlen, lag0, nlags):
    """Run `match_fn(toff, tlen, lag0, nlags) -> (diff, idx)` on this rank's contiguous shard of the
    planned queries and all-gather the results.  match_fn returns torch tensors (GPU: filled by
    sb_find_batch_device) or NumPy arrays (CPU tests)."""
    count = len(toff)
    lo, hi = shard_bounds(count, world_size, rank)
    d, i = match_fn(toff[lo:hi], tlen[lo:hi], lag0[lo:hi], nlags[lo:hi])
    if isinstance(d, np.ndarray):
        d, i = torch.from_numpy(np.ascontiguousarray(d, np.float32)), torch.from_numpy(np.ascontiguousarray(i, np.int64))
    return all_gather_results(dist, torch, d, i, count, world_size)
