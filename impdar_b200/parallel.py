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


def row_chunks(snum, nchunks):
    """[(r0, r1)] of the bottom-up row pipeline, listed top-down (chunk j = rows [snum j / n, snum (j+1) / n))."""
    n = max(1, min(int(nchunks), int(snum)))
    return [(snum * j // n, snum * (j + 1) // n) for j in range(n)]


def _kirchhoff_sharded_pipelined(x, travel_time_us, dist_km, vel, nearfield, rank, world, group, src, ranges,
                                 nchunks, compute_rows):
    """The exchange steps overlapped with the diffraction sum.  An output row only reads input rows at or below it
    (minus one for the d/dt stencil), so the input is broadcast bottom-up in row chunks, every chunk of this rank's
    output range is computed as soon as its rows have arrived, and its all-gather runs while the next chunk is
    computed.  Returns the assembled (snum, tnum) image, or None when `compute_rows` reports irregular trace spacing
    on the first chunk (nothing has been computed then; the input is completely broadcast on return)."""
    import torch
    import torch.distributed as dist
    S, T = x.shape
    xb, xe = ranges[rank]
    wmax = max(e - b for b, e in ranges)
    chunks = row_chunks(S, nchunks)
    # 1. every broadcast is enqueued up front, bottom-up; the collective stream runs them back to back
    arrive = {}
    u_hi = S
    for j in reversed(range(len(chunks))):
        u0 = max(chunks[j][0] - 1, 0)
        arrive[j] = dist.broadcast(x[u0:u_hi], src=src, group=group, async_op=True) if u0 < u_hi else None
        u_hi = min(u_hi, u0)
    # 2. chunk by chunk: wait for its rows, compute, start its all-gather
    block = torch.empty((S, max(xe - xb, 0)), dtype=x.dtype, device=x.device)
    padded = torch.zeros((S, wmax), dtype=x.dtype, device=x.device)
    stages, gathers = {}, {}
    g_hi = S
    for j in reversed(range(len(chunks))):
        r0, r1 = chunks[j]
        if arrive[j] is not None:
            arrive[j].wait()                       # the compute stream waits; the host does not
        if xe > xb:
            try:
                compute_rows(x, travel_time_us, dist_km, vel, nearfield, xb, xe, r0, r1, g_hi, block)
            except ValueError:
                if g_hi != S:
                    raise
                for w in arrive.values():          # irregular spacing: finish the broadcast, let the caller fall back
                    if w is not None:
                        w.wait()
                return None
            g_hi = r0
            padded[r0:r1, :xe - xb] = block[r0:r1]
        stages[j] = torch.empty((world, r1 - r0, wmax), dtype=x.dtype, device=x.device)
        gathers[j] = dist.all_gather_into_tensor(stages[j].view(world * (r1 - r0), wmax), padded[r0:r1], group=group,
                                                 async_op=True)
    # 3. assemble: (rank, rows, wmax) stages -> column blocks of the image
    out = torch.empty((S, T), dtype=x.dtype, device=x.device)
    for j in reversed(range(len(chunks))):
        r0, r1 = chunks[j]
        gathers[j].wait()
        for r, (b, e) in enumerate(ranges):
            if e > b:
                out[r0:r1, b:e] = stages[j][r, :, :e - b]
    return out


def kirchhoff_sharded_device(x, travel_time_us, dist_km, vel, nearfield, rank=None, world=None, gather=True,
                             compute=None, group=None, src=0, pipeline_chunks=4, compute_rows=None):
    """Kirchhoff migration of one radargram over all ranks of the process group.

    x : (snum, tnum) float32 tensor on this rank's device; only rank `src`'s content matters (it is
        broadcast to the others).  Returns the full (snum, tnum) migrated image on every rank if `gather`,
        else this rank's (snum, x_end - x_begin) block and its range.
    compute(x, travel_time_us, dist_km, vel, nearfield, x_begin, x_end) -> (snum, x_end-x_begin) tensor;
    defaults to the CUDA kernel (tests on CPU/gloo inject their own).
    pipeline_chunks > 1 (with `gather`): the broadcast, the kernels and the all-gather overlap in bottom-up row chunks
    through compute_rows(x, tt, dist, vel, nearfield, x_begin, x_end, s_begin, s_end, g_hi, out) - the CUDA row-range
    entry by default; irregular trace spacing falls back to the three phases back to back."""
    import torch
    import torch.distributed as dist
    default_compute = compute is None
    if compute is None:
        from .migrationlib import kirchhoff_device
        compute = kirchhoff_device
    if compute_rows is None and default_compute:
        from .migrationlib import kirchhoff_rows_device
        compute_rows = kirchhoff_rows_device
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    S, T = x.shape
    broadcast_done = False
    if world > 1 and gather and pipeline_chunks > 1 and compute_rows is not None:
        ranges = kirchhoff_output_ranges(T, world, travel_time_us, dist_km, vel)
        out = _kirchhoff_sharded_pipelined(x, travel_time_us, dist_km, vel, nearfield, rank, world, group, src, ranges,
                                           pipeline_chunks, compute_rows)
        if out is not None:
            return out
        broadcast_done = True
    if world > 1 and not broadcast_done:
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
