"""Data-parallel sharding of plate-crop batches across the GPUs of one box.

Crops are independent on the inference path (no batch statistics anywhere in GFPGANv1OCR), so a batch is split into
contiguous per-rank shards — the counterpart of the reference's `indices[rank::world]` sampler split
(basicsr/data/data_sampler.py:29-42) for an in-memory batch — each rank runs its shard in micro-batches through its
own engine, and NO collective is on the data path.  `gather_shards` (one all_gather of the outputs) exists for callers
that want the full result on every rank; bench.py does not use it.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous shard of n items; sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def micro_batches(lo, hi, micro_batch):
    return [(s, min(s + micro_batch, hi)) for s in range(lo, hi, micro_batch)]


def run_shard(fn, shard, micro_batch=64, out=None):
    """Applies `fn` to one rank's shard in micro-batches.  With `out` (a preallocated tensor of the shard's output shape)
    the results are written in place and `out` is returned; otherwise they are concatenated."""
    spans = micro_batches(0, shard.shape[0], micro_batch)
    if out is not None:
        for s, e in spans:
            out[s:e].copy_(fn(shard[s:e]))
        return out
    outs = [fn(shard[s:e]) for s, e in spans]
    return torch.cat(outs, 0) if outs else None


def run_sharded(fn, batch, rank, world, micro_batch=64):
    """Applies `fn` (micro-batch tensor -> tensor) to this rank's shard of `batch` and returns the shard's outputs
    concatenated (an empty tensor with the right trailing shape when the shard is empty)."""
    lo, hi = shard_bounds(batch.shape[0], rank, world)
    outs = [fn(batch[s:e]) for s, e in micro_batches(lo, hi, micro_batch)]
    if outs:
        return torch.cat(outs, 0)
    probe = fn(batch[:1])
    return probe[:0]


def gather_shards(local_out, n_total, group=None):
    """all_gather of ragged shards (padded to the largest shard) -> full [n_total, ...] tensor on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = local_out.new_zeros((max_n,) + tuple(local_out.shape[1:]))
    lo, hi = sizes[rank]
    pad[:hi - lo] = local_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:h - l] for b, (l, h) in zip(bufs, sizes)], 0)
