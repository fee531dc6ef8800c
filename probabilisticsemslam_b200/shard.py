"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed for plumbing.

* Independent association problems (frames x sliding windows) shard with NO data-path collective:
  rank r takes a contiguous slice of the batch, runs the same kernel, and results are gathered once.
* One large permanent is the only path with a real exchange step: the Gray-code index range
  [0, 2^(n-1)) is cut into `world` aligned pieces, each rank walks its piece, and the (hi, lo)
  double-double partial sums are all-gathered (16 bytes per rank) and added in rank order on every
  rank -- deterministic, and for a 16-byte message as cheap as an all-reduce.

The compute callables default to the CUDA path (api / device); the CPU tests of the host logic
(tests/test_sharding_gloo.py, gloo backend) inject stand-ins.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .synth import ProblemBatch


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gray_range(n_dim: int, world: int, rank: int) -> tuple[int, int]:
    """This rank's share of the NW Gray index range [0, 2^(n-1)).  world is rounded down to a power of
    two internally so that every boundary is a multiple of a large power of two (aligned chunks keep the
    kernel's column reads uniform); ranks beyond that get an empty range."""
    total = 1 << (n_dim - 1)
    parts = 1
    while parts * 2 <= world and parts * 2 <= total:
        parts *= 2
    if rank >= parts:
        return total, total
    return total * rank // parts, total * (rank + 1) // parts


def combine_partials(partials: np.ndarray, n_dim: int) -> float:
    """Adds (hi, lo) pairs in rank order (error-free two-sum on the high parts) and applies the NW factor."""
    hi, lo = 0.0, 0.0
    for h, l in np.asarray(partials, dtype=np.float64).reshape(-1, 2):
        s = hi + h
        bb = s - hi
        e = (hi - (s - bb)) + (h - bb)
        hi, lo = s, lo + e + l
    return float((4 * (n_dim & 1) - 2) * (hi + lo))


def sharded_assignment_prob(pb: ProblemBatch, k: int, *, compute: Callable | None = None, group=None,
                            gather: bool = True):
    """assignmentProb over a batch split across the ranks of `group`.

    Every rank passes the SAME full batch description (cheap: it is a few arrays) and computes only its
    slice.  Returns the flat weight vector of the whole batch on every rank when gather=True (one
    all_gather of variable-length pieces), else just (lo, hi, local flat weights)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(pb), world, rank)
    if compute is None:
        from . import api
        compute = lambda part, kk: api.assignment_prob_batch(part, kk, device=torch.cuda.current_device()).probs  # noqa: E731
    local = np.asarray(compute(pb.slice(lo, hi), k), dtype=np.float64) if hi > lo else np.zeros(0)
    if not gather or world == 1:
        return local if world == 1 else (lo, hi, local)
    sizes = (pb.nM.astype(np.int64) * (pb.nL.astype(np.int64) + 1))
    counts = [int(sizes[slice(*shard_bounds(len(pb), world, r))].sum()) for r in range(world)]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = max(counts) if counts else 0
    send = torch.zeros(width, dtype=torch.float64, device=dev)
    send[:local.size] = torch.from_numpy(local).to(dev)
    recv = [torch.zeros(width, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return np.concatenate([recv[r][:counts[r]].cpu().numpy() for r in range(world)])


def sharded_association_from_moments(frames, nonassign: float, k: int, *, compute: Callable | None = None, group=None):
    """getAssignmentProbs (assignment.cpp:38-74) for a list of frames -- (land_mean, land_cov, meas_mean, meas_cov) each,
    e.g. the frames x window frames of a sliding-window update -- split across the ranks of `group`: rank r builds the
    cost matrices of its contiguous slice on its GPU and runs the association pipeline there; one all_gather of the
    variable-length weight vectors at the end, no other exchange.  Returns the list of (nM, nL+1) tables on every rank.
    `compute(frames, nonassign, k) -> list of tables` defaults to the CUDA pipeline."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(frames), world, rank)
    if compute is None:
        from . import api
        compute = lambda fr, na, kk: api.association_from_moments_batch(fr, na, kk, device=torch.cuda.current_device())  # noqa: E731
    shapes = [(int(np.asarray(f[2]).reshape(-1, 3).shape[0]), int(np.asarray(f[0]).reshape(-1, 3).shape[0]) + 1) for f in frames]
    local_tabs = compute(frames[lo:hi], nonassign, k) if hi > lo else []
    local = np.concatenate([np.asarray(t, np.float64).reshape(-1) for t in local_tabs]) if local_tabs else np.zeros(0)
    if world > 1:
        counts = [sum(m * w for m, w in shapes[slice(*shard_bounds(len(frames), world, r))]) for r in range(world)]
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        width = max(counts) if counts else 0
        send = torch.zeros(width, dtype=torch.float64, device=dev)
        send[:local.size] = torch.from_numpy(local).to(dev)
        recv = [torch.zeros(width, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
        flat = np.concatenate([recv[r][:counts[r]].cpu().numpy() for r in range(world)])
    else:
        flat = local
    out, o = [], 0
    for m, w in shapes:
        out.append(flat[o:o + m * w].reshape(m, w))
        o += m * w
    return out


def sharded_permanent(a: np.ndarray, *, partial: Callable | None = None, group=None) -> float:
    """Exact permanent of ONE square matrix with the Gray range split over the ranks of `group`
    (BASELINE.json configs[4]).  `partial(a, begin, end) -> (hi, lo)`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(a.shape[0])
    begin, end = gray_range(n, world, rank)
    if partial is None:
        from . import api
        partial = lambda m, b, e: api.permanent_range(m, b, e, device=torch.cuda.current_device())  # noqa: E731
    hi, lo = partial(a, begin, end) if end > begin else (0.0, 0.0)
    if world == 1:
        return combine_partials(np.array([hi, lo]), n)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([hi, lo], dtype=torch.float64, device=dev)
    parts = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return combine_partials(torch.stack(parts).cpu().numpy(), n)
