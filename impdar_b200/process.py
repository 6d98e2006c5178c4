"""Batched, device-resident orchestration of the hot path: the B200 counterpart of ``impdar.lib.process.process``
(lib/process.py:72-197) - the caller of every hot-path method.

The reference runs ``vertical_band_pass -> hfilt -> adaptive hfilt -> ... -> migrate('stolt')`` profile by profile on
host arrays (process.py:151-193).  Calling the drop-in methods one by one costs a host->device and a device->host
copy per step; here a profile is uploaded once, every hot-path step runs on the GPU on the profile's own CUDA
stream, and the result is downloaded once.  Profiles are independent, so ``n_streams`` of them are in flight at the
same time: the upload of profile p+1, the kernels of profile p and the download of profile p-1 overlap (two copy
engines + SMs).  Under ``torchrun`` every rank takes the profiles ``p mod world == rank`` (no collective,
SURVEY.md 8e).

Same argument checks, same order of steps, same flags and same dtypes as the reference.  Steps that are not on the
hot path (hcrop, restack, reverse, nmo, denoise, interp, crop) are delegated to the object's own methods - with
``impdar_b200.install()`` on ImpDAR's RadarData these are the reference's - and split the device-resident chain
where the reference orders them between filters and migration.  There is no CPU fallback for the hot-path steps.
"""
import numpy as np

from . import device, parallel


def _need(dat, name):
    fn = getattr(dat, name, None)
    if fn is None:
        raise NotImplementedError('%s is outside the B200 hot path and %s does not provide it; use ImpDAR\'s RadarData '
                                  '(impdar_b200.install())' % (name, type(dat).__name__))
    return fn


def _host_dtype_after(steps, in_dtype):
    """dtype the reference leaves in dat.data after `steps` (filters keep it, Stolt follows np.fft.irfft2)."""
    dt = np.dtype(in_dtype)
    for name, _ in steps:
        if name == 'migrate':
            dt = np.dtype(np.float32) if dt == np.float32 else np.dtype(np.float64)
    return dt


def _run_chain_on_device(dat, steps):
    """Apply (name, args) steps through the drop-in methods; dat.data is a CUDA tensor, so results stay on the GPU."""
    for name, args in steps:
        if name == 'vbp':
            dat.vertical_band_pass(*args)
        elif name == 'hfilt':
            dat.hfilt(ftype='hfilt', bounds=args)
        elif name == 'ahfilt':
            dat.hfilt(ftype='adaptive', window_size=args)
        elif name == 'migrate':
            dat.migrate(mtype='stolt')


_stream_pool = {}   # device index -> [torch.cuda.Stream]; persistent, because scratch buffers and cuFFT plans are per stream


def _streams(n):
    import torch
    pool = _stream_pool.setdefault(torch.cuda.current_device(), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream())
    return pool[:n]


def run_device_chain(dats, steps, n_streams=3):
    """Upload once, run `steps` on the GPU, download once - `n_streams` profiles in flight."""
    import torch
    if not steps or not dats:
        return
    device.require_cuda()
    n_streams = max(1, min(int(n_streams), len(dats), 8))
    streams = _streams(n_streams)
    pending = [None] * n_streams   # (dat, pinned host tensor, event, final dtype)

    def finish(slot):
        if pending[slot] is None:
            return
        dat, host, ev, np_dtype = pending[slot]
        ev.synchronize()
        out = host.numpy()
        dat.data = out if out.dtype == np_dtype else out.astype(np_dtype)
        pending[slot] = None

    for i, dat in enumerate(dats):
        slot = i % n_streams
        finish(slot)
        if device.is_device_array(dat.data):
            _run_chain_on_device(dat, steps)      # already device resident: stays there, caller's stream
            continue
        src = np.asarray(dat.data)
        if not np.issubdtype(src.dtype, np.floating):
            # integer radargrams need the reference's cast-back (truncation) after every step: per-step path
            _run_chain_on_device(dat, steps)
            continue
        final_dtype = _host_dtype_after(steps, src.dtype)
        with torch.cuda.stream(streams[slot]):
            dat.data = device.to_device(src, torch.float32 if src.dtype == np.float32 else torch.float64)
            _run_chain_on_device(dat, steps)
            res = dat.data
            want = torch.float64 if final_dtype == np.float64 else torch.float32
            if res.dtype != want:
                res = res.to(want)                 # cast on the device: the download is one DMA into pinned memory
            host = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
            host.copy_(res.contiguous(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(streams[slot])
        dat.data = None                             # filled in by finish(); never left pointing at stale input
        pending[slot] = (dat, host, ev, final_dtype)
    for slot in range(n_streams):
        finish(slot)


def process(RadarDataList, interp=None, rev=False, vbp=None, hfilt=None, ahfilt=None, nmo=None, crop=None,
            hcrop=None, restack=None, denoise=None, migrate=None, n_streams=3, **kwargs):
    """Perform one or more processing steps on a list of RadarData; mirrors lib/process.py:72-197.

    Returns True if a step was performed.  The hot-path steps (vbp, hfilt, ahfilt, migrate) run device resident."""
    done_stuff = False

    # ---- argument checking, as the reference (process.py:101-134)
    if crop is not None:
        try:
            crop = (float(crop[0]), crop[1], crop[2])
        except ValueError:
            raise ValueError('First element of crop must be a float')
        except TypeError:
            raise TypeError('Crop must be subscriptible')
    if hcrop is not None:
        try:
            hcrop = (float(hcrop[0]), hcrop[1], hcrop[2])
        except ValueError:
            raise ValueError('First element of hcrop must be a float')
        except TypeError:
            raise TypeError('hcrop must be subscriptible')
        for dat in RadarDataList:
            _need(dat, 'hcrop')(*hcrop)
        done_stuff = True
    if denoise is not None:
        try:
            assert (type(denoise[0]) is int)
            assert (type(denoise[1]) is int)
        except (ValueError, TypeError, AssertionError, IndexError):
            raise ValueError('Denoise must be two integers giving vertical and horizontal window sizes')
    if vbp is not None:
        if not hasattr(vbp, '__iter__'):
            raise TypeError('vbp must be a tuple with first two elements \
                            [low] [high] MHz')
    if interp is not None:
        try:
            float(interp[0])
            interp[1]
        except (ValueError, TypeError, IndexError):
            raise ValueError('interp must be a target spacing (float) then a gps filename')

    if restack is not None:
        for dat in RadarDataList:
            if isinstance(restack, (list, tuple)):
                restack = int(restack[0])
            _need(dat, 'restack')(restack)
        done_stuff = True

    if rev:
        for dat in RadarDataList:
            _need(dat, 'reverse')()
        done_stuff = True

    # ---- the device-resident chain(s): filters, [host steps the reference orders in between], migration
    filters = []
    if vbp is not None:
        filters.append(('vbp', tuple(vbp)))
    if hfilt is not None:
        filters.append(('hfilt', hfilt))
    if ahfilt:
        filters.append(('ahfilt', ahfilt))
    tail = [('migrate', None)] if migrate is not None else []
    host_between = nmo is not None or denoise is not None or interp is not None or crop is not None

    if not host_between:
        run_device_chain(RadarDataList, filters + tail, n_streams)
        done_stuff = done_stuff or bool(filters or tail)
        return done_stuff

    run_device_chain(RadarDataList, filters, n_streams)
    done_stuff = done_stuff or bool(filters)

    if nmo is not None:
        if isinstance(nmo, (float, int)):
            print('One nmo value given. Assuming that this is the separation. \
                  Uice=1.6')
            nmo = (nmo, 1.6)
        for dat in RadarDataList:
            _need(dat, 'nmo')(*nmo)
        done_stuff = True

    if denoise is not None:
        for dat in RadarDataList:
            _need(dat, 'denoise')(*denoise)
        done_stuff = True

    if interp is not None:
        from impdar.lib.gpslib import interp as interpdeep   # the reference's own (process.py:24, :178)
        interpdeep(RadarDataList, float(interp[0]), interp[1])
        done_stuff = True

    if crop is not None:
        for dat in RadarDataList:
            _need(dat, 'crop')(*crop)
        done_stuff = True

    run_device_chain(RadarDataList, tail, n_streams)
    done_stuff = done_stuff or bool(tail)
    return done_stuff


def process_sharded(RadarDataList, rank=None, world=None, **kwargs):
    """`process` on this rank's share of the profiles (p mod world == rank; SURVEY.md 8e: independent units, no
    collective).  Returns (performed, indices of the profiles this rank processed)."""
    if rank is None or world is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    mine = parallel.profiles_for_rank(len(RadarDataList), rank, world)
    done = process([RadarDataList[p] for p in mine], **kwargs)
    return done, mine
