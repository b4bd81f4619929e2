"""Multi-GPU plumbing: independent trajectories shard across ranks, one final gather.

The reference runs every trajectory as an independent ACE process behind a thread pool
(SURVEY 2.3), so the batch axis shards trivially: contiguous, step-count-balanced
blocks per rank, the PT and problem replicated per GPU, and a single collective at the end --
an all-gather of the per-rank result blocks over NCCL/NVLink (gloo in the CPU tests).  There is
no per-step communication.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def balanced_blocks(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Split indices 0..n-1 into `world` contiguous blocks of nearly equal total cost.

    `costs[i]` is the step count of trajectory i (``tend_i`` varies across a G(t,tau) sweep,
    reference ``two_time/correlations.py:156``).  Returns ``[(start, stop)] * world``.
    """
    c = np.asarray(costs, dtype=float)
    n = len(c)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard(items: Sequence, rank: int, world: int, costs: Optional[Sequence[float]] = None):
    """This rank's contiguous slice of `items` (and its (start, stop))."""
    if costs is None:
        costs = [1.0] * len(items)
    a, b = balanced_blocks(costs, world)[rank]
    return items[a:b], (a, b)


def all_gather_blocks(local: np.ndarray, counts: Sequence[int], device=None, group=None) -> np.ndarray:
    """All-gather ragged per-rank blocks (rank r contributes ``counts[r]`` rows of `local`'s
    row shape); returns the concatenation on every rank.  complex128 travels as float64 pairs.
    Uses the default process group's backend: NCCL with CUDA tensors, gloo on the CPU."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(counts) != world or counts[rank] != local.shape[0]:
        raise ValueError("counts must list every rank's row count")
    row_shape = local.shape[1:]
    row_elems = int(np.prod(row_shape)) if row_shape else 1
    is_c = np.iscomplexobj(local)
    flat = np.ascontiguousarray(local).reshape(local.shape[0], row_elems)
    if is_c:
        flat = flat.view(np.float64)
    width = flat.shape[1] if flat.ndim == 2 and flat.shape[0] else row_elems * (2 if is_c else 1)
    cmax = max(counts) if counts else 0
    send = torch.zeros((cmax, width), dtype=torch.float64, device=device)
    if local.shape[0]:
        send[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64)).to(send.device)
    recv = torch.empty((world * cmax, width), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    host = recv.cpu().numpy().reshape(world, cmax, width)
    parts = [host[r, :counts[r]] for r in range(world)]
    out = np.concatenate(parts, axis=0) if parts else np.zeros((0, width))
    if is_c:
        out = np.ascontiguousarray(out).view(np.complex128)
    return out.reshape((out.shape[0],) + tuple(row_shape))


def is_multi_rank(group=None) -> bool:
    """True inside an initialised torch.distributed job with more than one rank (never imports torch
    itself: a process that has not imported torch cannot have a process group)."""
    import sys
    if "torch" not in sys.modules:
        return False
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _check_same_jobs(jobs, device=None, group=None) -> None:
    """Every rank must have submitted the same job list in the same order (the gather below assembles blocks by
    position): compare a digest of (count, step counts, tail lengths, start times) across the ranks."""
    import hashlib
    import torch
    import torch.distributed as dist
    h = hashlib.sha256()
    h.update(np.asarray([len(jobs)], dtype=np.int64).tobytes())
    h.update(np.asarray([[j.n_steps, getattr(j, "tail_rows", 0) or 0, len(getattr(j, "mtos", ()))] for j in jobs],
                        dtype=np.int64).tobytes())
    h.update(np.asarray([[getattr(j, "t_start", 0.0), getattr(j, "dt", 0.0)] for j in jobs], dtype=np.float64).tobytes())
    mine = torch.from_numpy(np.frombuffer(h.digest()[:16], dtype=np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    world = dist.get_world_size(group)
    allv = torch.empty(world * 2, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().numpy().reshape(world, 2)
    if not (allv == allv[0]).all():
        raise RuntimeError("run_jobs_sharded: the ranks submitted different job lists (count / lengths / order); "
                           "distributed sharding needs the same sweep on every rank")


def run_jobs_sharded(engine, prob, pt, jobs, device=None, group=None, **kw) -> List[np.ndarray]:
    """Propagate this rank's share of `jobs` and all-gather the results (every rank returns the
    full list, like ``wait(futures)`` in the reference).  Requires an initialised process group;
    with world size 1 it is `engine.run_jobs`."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return engine.run_jobs(prob, pt, jobs, **kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None and dist.get_backend(group) == "nccl":
        import torch
        device = torch.device("cuda", int(getattr(engine, "device", 0)))
    _check_same_jobs(jobs, device, group)
    costs = [max(1, j.n_steps) for j in jobs]
    blocks = balanced_blocks(costs, world)
    a, b = blocks[rank]
    mine = engine.run_jobs(prob, pt, jobs[a:b], **kw) if b > a else []
    n_out = prob.n_out
    # rows every rank can predict: all output rows, or only the requested tail
    rows = [min(j.n_steps + 1, getattr(j, "tail_rows", 0) or j.n_steps + 1) for j in jobs]
    # pack ragged [n_out, rows_i] results into one [sum rows, n_out] block per rank
    local = (np.concatenate([m.T for m in mine], axis=0) if mine else np.zeros((0, n_out), complex))
    counts = [int(sum(rows[x:y])) for (x, y) in blocks]
    full = all_gather_blocks(np.ascontiguousarray(local), counts, device=device, group=group)
    out, off = [], 0
    for r in rows:
        out.append(np.ascontiguousarray(full[off:off + r].T))
        off += r
    return out


def run_arrays_sharded(engine, prob, pt, arr, device=None, group=None, **kw):
    """:func:`run_jobs_sharded` for the array route (``Engine.run_arrays``): contiguous, step-count-balanced row
    blocks of the job table per rank, one all-gather of the kept rows.  Returns ``(out, out_off, n_rows)`` of the
    WHOLE sweep on every rank."""
    import hashlib
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return engine.run_arrays(prob, pt, arr, **kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", int(getattr(engine, "device", 0)))
    h = hashlib.sha256()
    for a in (arr.n_steps, arr.tail, arr.n_ev, arr.ev_step, arr.shift):
        h.update(np.ascontiguousarray(a, dtype=np.int64).tobytes())
    mine = torch.from_numpy(np.frombuffer(h.digest()[:16], dtype=np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    allv = torch.empty(world * 2, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().numpy().reshape(world, 2)
    if not (allv == allv[0]).all():
        raise RuntimeError("run_arrays_sharded: the ranks submitted different job lists (count / lengths / order); "
                           "distributed sharding needs the same sweep on every rank")
    n_out = prob.n_out
    n_steps = arr.n_steps.astype(np.int64)
    n_rows = np.where(arr.tail > 0, np.minimum(n_steps + 1, arr.tail), n_steps + 1).astype(np.int64)
    blocks = balanced_blocks(np.maximum(1, n_steps), world)
    a, b = blocks[rank]
    if b > a:
        out, _, _ = engine.run_arrays(prob, pt, arr.rows(a, b), **kw)
        local = np.ascontiguousarray(out).reshape(-1, n_out)
    else:
        local = np.zeros((0, n_out), complex)
    counts = [int(n_rows[x:y].sum()) for (x, y) in blocks]
    full = all_gather_blocks(local, counts, device=device, group=group)
    out_off = np.zeros(len(n_rows), dtype=np.int64)
    out_off[1:] = np.cumsum(n_rows[:-1] * n_out)
    return np.ascontiguousarray(full).reshape(-1), out_off, n_rows
