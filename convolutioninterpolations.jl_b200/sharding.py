"""Multi-GPU plumbing for the path: the work units are independent (queries; output slabs of a sweep), so ranks
share nothing but the replicated inputs - no data-path collective.  `torch.distributed` is used for
the optional final gather and for timing reductions only."""
import numpy as np


def shard_bounds(n, world, rank, align=32):
    """Contiguous, balanced, `align`-aligned split of range(n) over `world` ranks -> (lo, hi)."""
    if world <= 1:
        return 0, n
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def eval_sharded(evaluate, coords, world, rank, gather=None):
    """Evaluate this rank's slice of a point list.  `evaluate(*coords_slice)` is `itp.eval`; `gather(local, lo, hi, n)`
    (optional) assembles the full result, e.g. with all_gather_into_tensor.  Returns (local_result, (lo, hi))."""
    n = len(coords[0])
    lo, hi = shard_bounds(n, world, rank)
    local = evaluate(*[c[lo:hi] for c in coords])
    if gather is not None:
        return gather(local, lo, hi, n), (lo, hi)
    return local, (lo, hi)


def output_slab(knots_out, world, rank):
    """Slab decomposition of a separable sweep's OUTPUT along the last axis: every rank keeps the (replicated) input
    and computes the output points whose last-axis knot falls in its slab.  Because an output point depends only on
    its own knot, the concatenation of the slabs along the last axis is bit-identical to the single-device result."""
    knots_out = tuple(knots_out)
    lo, hi = shard_bounds(len(knots_out[-1]), world, rank, align=1)
    return knots_out[:-1] + (np.asarray(knots_out[-1])[lo:hi],), (lo, hi)


def max_over_ranks_ms(local_ms, world, device=None):
    """Device-time reduction used by bench.py: the job is as slow as its slowest rank."""
    if world <= 1:
        return float(local_ms)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
