"""Multi-GPU sharding of the two hot paths: one process per GPU, `torch.distributed` for the
plumbing (NCCL on GPUs, gloo in the CPU tests).

MSM is linear in its points: rank g takes the contiguous chunk [g*N/G, (g+1)*N/G), reduces it to
ONE Jacobian point with the single-GPU pipeline, the 144-byte partials are all-gathered and added
(`sppark_b200_msm_combine`; NCCL's reductions cannot add curve points).  The reference has no
multi-GPU path ("Parallelism inventory"); this is new functionality behind the same
per-GPU entry points.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous chunk of [0, n) owned by `rank`; chunks differ by at most one element and are
    multiples of nothing in particular (the MSM kernels accept any length, including 0)."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def all_gather_partials(partial, device=None):
    """partial: (k,) uint64 host array -> (world, k) uint64, same order on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(partial, dtype=np.uint64)[None, :].copy()
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy().view(np.uint64)


def msm_sharded(local_msm, combine, partial_words, device=None):
    """Run `local_msm()` (this rank's chunk -> Jacobian words), exchange, and fold with
    `combine((world, k) array) -> (k,)`.  Every rank returns the full result."""
    part = np.asarray(local_msm(), dtype=np.uint64)
    assert part.shape == (partial_words,)
    return combine(all_gather_partials(part, device))
