"""Batched, device-resident orchestration of the hot path: the B200 counterpart of ``impdar.lib.process.process``
(lib/process.py:72-197) - the caller of every hot-path method.

The reference runs ``vertical_band_pass -> hfilt -> adaptive hfilt -> ... -> migrate('stolt')`` profile by profile on
host arrays (process.py:151-193).  Calling the drop-in methods one by one costs a host->device and a device->host
copy per step; here a profile is uploaded once, every hot-path step runs on the GPU on the profile's own CUDA
stream, and the result is downloaded once.  Profiles are independent, so ``n_streams`` of them are in flight at the
same time: the upload of profile p+1, the kernels of profile p and the download of profile p-1 overlap (two copy
engines + SMs).  Under ``torchrun`` every rank takes the profiles ``p mod world == rank`` (no collective,
SURVEY.md 8e).

Same argument checks, same order of steps, same flags and same dtypes as the reference.  The index / resampling
steps (hcrop, restack, reverse, nmo, crop - impdar_b200.processing) run on the device inside the
same chain, so  hcrop -> restack -> reverse -> vbp -> hfilt -> ahfilt -> nmo -> denoise -> crop -> migrate  is one upload
and one download per profile.  Only interp (GPS file I/O through impdar.lib.gpslib) stays a host step and splits the
chain where the reference orders it (between denoise and crop).  One deviation: the reference
applies hcrop while it is still checking arguments (process.py:111-119); here every argument is checked first, so a
bad later argument leaves the profiles untouched.  There is no CPU fallback for the device steps.
"""
import numpy as np

from . import device, parallel


def _need(dat, name):
    fn = getattr(dat, name, None)
    if fn is None:
        raise NotImplementedError('%s is outside the B200 hot path and %s does not provide it; use ImpDAR\'s RadarData '
                                  '(impdar_b200.install())' % (name, type(dat).__name__))
    return fn


def _host_dtype_after(steps, in_dtype, dat=None):
    """dtype the reference leaves in dat.data after `steps`: filters, block crops and reverse keep it; restack
    (np.zeros), nmo (np.empty) and the per-trace pretrigger crop allocate float64; Stolt follows np.fft.irfft2."""
    dt = np.dtype(in_dtype)
    for name, args in steps:
        if name == 'migrate':
            dt = np.dtype(np.float32) if dt == np.float32 else np.dtype(np.float64)
        elif name in ('restack', 'nmo', 'denoise'):
            dt = np.dtype(np.float64)
        elif name == 'crop' and args[2] == 'pretrig' and isinstance(getattr(dat, 'trig', None), np.ndarray):
            dt = np.dtype(np.float64)
    return dt


def _run_chain_on_device(dat, steps):
    """Apply (name, args) steps through the drop-in methods; dat.data is a CUDA tensor, so results stay on the GPU."""
    for name, args in steps:
        if name == 'hcrop':
            _need(dat, 'hcrop')(*args)
        elif name == 'restack':
            _need(dat, 'restack')(args)
        elif name == 'rev':
            _need(dat, 'reverse')()
        elif name == 'nmo':
            _need(dat, 'nmo')(*args)
        elif name == 'crop':
            _need(dat, 'crop')(*args)
        elif name == 'denoise':
            _need(dat, 'denoise')(*args)
        elif name == 'vbp':
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

    # Exception safety (the reference processes profile after profile: when a step raises for profile i, profiles
    # < i are completely processed and profile i keeps whatever state the failing step left): every profile already
    # queued is finished on the way out, and the failing profile gets a host array back - the partial result the
    # device holds (downloaded, like the reference's partially processed dat.data), or its untouched input.
    try:
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
            final_dtype = _host_dtype_after(steps, src.dtype, dat)
            with torch.cuda.stream(streams[slot]):
                try:
                    dat.data = device.to_device(src, torch.float32 if src.dtype == np.float32 else torch.float64)
                    dat._b200_reference_dtypes = True   # restack / nmo produce the reference's float64 on the device
                    try:
                        _run_chain_on_device(dat, steps)
                    finally:
                        del dat._b200_reference_dtypes
                    res = dat.data
                    want = torch.float64 if final_dtype == np.float64 else torch.float32
                    if res.dtype != want:
                        res = res.to(want)             # cast on the device: the download is one DMA into pinned memory
                    host = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
                    host.copy_(res.contiguous(), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(streams[slot])
                except BaseException:
                    partial = dat.data
                    if device.is_device_array(partial):
                        try:
                            streams[slot].synchronize()
                            dat.data = partial.cpu().numpy()
                        except Exception:
                            dat.data = src
                    elif partial is None:
                        dat.data = src
                    raise
            dat.data = None                             # filled in by finish(); never left pointing at stale input
            pending[slot] = (dat, host, ev, final_dtype)
    finally:
        for slot in range(n_streams):
            finish(slot)


def process(RadarDataList, interp=None, rev=False, vbp=None, hfilt=None, ahfilt=None, nmo=None, crop=None,
            hcrop=None, restack=None, denoise=None, migrate=None, n_streams=3, **kwargs):
    """Perform one or more processing steps on a list of RadarData; mirrors lib/process.py:72-197.

    Returns True if a step was performed.  Every step except interp runs device resident."""
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

    if restack is not None and isinstance(restack, (list, tuple)):
        restack = int(restack[0])
    if nmo is not None and isinstance(nmo, (float, int)):
        print('One nmo value given. Assuming that this is the separation. \
              Uice=1.6')
        nmo = (nmo, 1.6)

    # ---- the device-resident chain, in the reference's order (process.py:111-193)
    head = []
    if hcrop is not None:
        head.append(('hcrop', hcrop))
    if restack is not None:
        head.append(('restack', restack))
    if rev:
        head.append(('rev', None))
    if vbp is not None:
        head.append(('vbp', tuple(vbp)))
    if hfilt is not None:
        head.append(('hfilt', hfilt))
    if ahfilt:
        head.append(('ahfilt', ahfilt))
    if nmo is not None:
        head.append(('nmo', tuple(nmo)))
    if denoise is not None:
        head.append(('denoise', tuple(denoise)))
    tail = []
    if crop is not None:
        tail.append(('crop', crop))
    if migrate is not None:
        tail.append(('migrate', None))
    host_between = interp is not None

    if not host_between:
        run_device_chain(RadarDataList, head + tail, n_streams)
        return bool(head or tail)

    run_device_chain(RadarDataList, head, n_streams)

    if interp is not None:
        from impdar.lib.gpslib import interp as interpdeep   # the reference's own (process.py:24, :178)
        interpdeep(RadarDataList, float(interp[0]), interp[1])

    run_device_chain(RadarDataList, tail, n_streams)
    return True


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
