"""Stream sharding for multi-GPU jobs (one process per GPU, torch.distributed for the plumbing).

The scan has no data-path exchange: streams are independent (Design/FPGA.v:54-57 keeps the two
streams' bitmaps separate and only shares reads of the CSR), so rank r scans the contiguous stream range
shard_range(n, r, world) against its own copy of the NFA.  The only exchange is at the end of a step:
a SUM all-reduce of the per-state match counts (<= 8 * n_states bytes) and, when the caller wants the
match records in one place, a gather of the per-rank record arrays to rank 0.
"""
import numpy as np


def shard_range(n_streams, rank, world):
    """Contiguous, balanced split: returns (first_stream, count) of `rank`."""
    first = n_streams * rank // world
    last = n_streams * (rank + 1) // world
    return first, last - first


def reduce_counts(counts, dist):
    """In-place SUM all-reduce of a torch int64 per-state count vector (NCCL on GPU, gloo on CPU)."""
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def gather_records(records, dist, device=None):
    """Gathers every rank's match records (numpy structured array, regex_fpga_b200.MATCH_DTYPE, stream ids
    already global via stream_id_base) to all ranks and returns them in canonical (stream,pos,state) order.
    Ranks own disjoint ascending stream ranges, so concatenation in rank order preserves the order of
    already-sorted per-rank arrays."""
    import torch
    world = dist.get_world_size()
    flat = torch.from_numpy(np.ascontiguousarray(records).view(np.uint32).astype(np.int64).reshape(-1))
    if device is not None:
        flat = flat.to(device)
    n = torch.tensor([flat.numel()], dtype=torch.int64, device=flat.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=torch.int64, device=flat.device)
    padded[: flat.numel()] = flat
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = np.concatenate([p[:s].cpu().numpy() for p, s in zip(parts, sizes)]).astype(np.uint32)
    return out.view(records.dtype).reshape(-1)
