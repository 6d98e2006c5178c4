"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

* Independent profiles (the vbp -> hfilt -> Stolt pipeline, any per-profile call) shard round-robin with no
  collective at all: ``profiles_for_rank``.
* One large Kirchhoff radargram shards by contiguous OUTPUT-trace ranges.  Every output trace needs input
  traces up to one aperture away, so the input radargram is broadcast once (NCCL over NVLink) and the
  (snum, range) output blocks are all-gathered.  Ranges are balanced by pair count, not width: traces near
  the ends of the profile see half an aperture.
"""
import numpy as np


def profiles_for_rank(n_profiles, rank, world):
    """Indices of the profiles rank `rank` owns (p mod world == rank, process.py:151-193's serial loop split)."""
    return list(range(rank, n_profiles, world))


def kirchhoff_trace_cost(travel_time_us, dist_km, vel, n_depths=32, n_traces=2048):
    """Relative work per output trace: number of in-aperture input traces summed over a subsample of depths.

    The count is a smooth (piecewise linear) function of the trace position, so for monotone trace positions it is
    evaluated at `n_traces` evenly spaced traces and interpolated: O(n_depths * n_traces * log tnum) host work,
    well under a millisecond, instead of a searchsorted over every trace per depth."""
    tt = np.asarray(travel_time_us, dtype=np.float64) / 1e6
    dist = np.asarray(dist_km, dtype=np.float64) * 1e3
    T = len(dist)
    tmax = tt.max()
    zs = vel * tt / 2.0
    idx = np.unique(np.linspace(0, len(tt) - 1, min(n_depths, len(tt))).astype(int))
    a2 = (vel * tmax / 2.0) ** 2 - zs[idx] ** 2
    a = np.sqrt(a2[a2 >= 0])
    if T < 2 or not np.all(np.diff(dist) >= 0):
        return np.full(T, float(len(a) * T) + 1.0)
    pick = np.unique(np.linspace(0, T - 1, min(n_traces, T)).astype(int))
    d = dist[pick]
    hi = np.searchsorted(dist, (d[None, :] + a[:, None]).ravel(), side='right')
    lo = np.searchsorted(dist, (d[None, :] - a[:, None]).ravel(), side='left')
    cost = (hi - lo).reshape(len(a), len(pick)).sum(axis=0).astype(np.float64)
    if len(pick) == T:
        return cost + 1.0
    return np.interp(np.arange(T), pick, cost) + 1.0


_range_cache = {}


def kirchhoff_output_ranges(tnum, world, travel_time_us, dist_km, vel, align=8):
    """[(x_begin, x_end)] per rank: contiguous, covering [0, tnum), balanced by pair count, boundaries
    aligned to the kernel's 8-trace CTA tile."""
    tt = np.ascontiguousarray(travel_time_us, dtype=np.float64)
    dk = np.ascontiguousarray(dist_km, dtype=np.float64)
    key = (int(tnum), int(world), float(vel), int(align), hash(tt.tobytes()), hash(dk.tobytes()))
    hit = _range_cache.get(key)
    if hit is not None:
        return list(hit)
    cost = kirchhoff_trace_cost(tt, dk, vel)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world):
        x = int(np.searchsorted(cum, cum[-1] * r / world))
        x = int(round(x / align)) * align
        x = min(max(x, bounds[-1]), tnum)
        bounds.append(x)
    bounds.append(tnum)
    ranges = [(bounds[r], bounds[r + 1]) for r in range(world)]
    if len(_range_cache) > 64:
        _range_cache.clear()
    _range_cache[key] = tuple(ranges)
    return ranges


def kirchhoff_output_range(tnum, rank, world, travel_time_us, dist_km, vel):
    return kirchhoff_output_ranges(tnum, world, travel_time_us, dist_km, vel)[rank]


def kirchhoff_sharded_device(x, travel_time_us, dist_km, vel, nearfield, rank=None, world=None, gather=True,
                             compute=None, group=None, src=0):
    """Kirchhoff migration of one radargram over all ranks of the process group.

    x : (snum, tnum) float32 tensor on this rank's device; only rank `src`'s content matters (it is
        broadcast to the others).  Returns the full (snum, tnum) migrated image on every rank if `gather`,
        else this rank's (snum, x_end - x_begin) block and its range.
    compute(x, travel_time_us, dist_km, vel, nearfield, x_begin, x_end) -> (snum, x_end-x_begin) tensor;
    defaults to the CUDA kernel (tests on CPU/gloo inject their own)."""
    import torch
    import torch.distributed as dist
    if compute is None:
        from .migrationlib import kirchhoff_device
        compute = kirchhoff_device
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    S, T = x.shape
    if world > 1:
        dist.broadcast(x, src=src, group=group)          # the one exchange step on the input side
    ranges = kirchhoff_output_ranges(T, world, travel_time_us, dist_km, vel)
    xb, xe = ranges[rank]
    block = compute(x, travel_time_us, dist_km, vel, nearfield, xb, xe) if xe > xb else \
        torch.empty((S, 0), dtype=x.dtype, device=x.device)
    if not gather:
        return block, (xb, xe)
    if world == 1:
        return block
    wmax = max(e - b for b, e in ranges)
    padded = torch.zeros((S, wmax), dtype=x.dtype, device=x.device)
    padded[:, :xe - xb] = block
    allb2 = torch.empty((world * S, wmax), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(allb2, padded, group=group)      # (S, range) blocks, concatenated along rows
    allb = allb2.view(world, S, wmax)
    out = torch.empty((S, T), dtype=x.dtype, device=x.device)
    for r, (b, e) in enumerate(ranges):
        if e > b:
            out[:, b:e] = allb[r, :, :e - b]
    return out
