"""Host-side plumbing for multi-GPU runs: replicas only.

A single scan registration does not shard (its only cross-point reduction is 21 + 6 + 1 doubles), so N GPUs
run N independent groups of sequences: rank r owns sequences r*S .. r*S+S-1, there is no collective on the data
path, and torch.distributed is used only to bracket the timed region and to take the max over ranks.
"""
from __future__ import annotations


def sequence_seeds(rank: int, world_size: int, seqs_per_rank: int, base: int = 0) -> list[int]:
    """Seeds of the sequences rank `rank` owns; disjoint across ranks, contiguous over the job."""
    if not (0 <= rank < world_size) or seqs_per_rank < 1:
        raise ValueError("bad rank / world_size / seqs_per_rank")
    return [base + rank * seqs_per_rank + s for s in range(seqs_per_rank)]


def job_scans(world_size: int, seqs_per_rank: int, steps: int) -> int:
    """Scans the whole job processes in `steps` lock-step frames (weak scaling: per-rank work is fixed)."""
    return world_size * seqs_per_rank * steps


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def barrier(sync_cuda: bool = True) -> None:
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if sync_cuda and torch.cuda.is_available():
        torch.cuda.synchronize()
