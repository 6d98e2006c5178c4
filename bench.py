#!/usr/bin/env python
"""Benchmark of the B200 migration/filtering hot path (contract: task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path over one synthetic radargram (or one
batch of profiles) per GPU.  Default workload = BASELINE.json configs[1]: Kirchhoff migration of a
4096-trace x 2048-sample radargram, constant velocity 1.69e8 m/s.

  value      migrated samples/s, whole job, inputs resident in HBM, CUDA-event timed (max over ranks)
  e2e        same metric through the reference-facing plugin call (RadarData.migrate on HOST numpy data:
             H2D + kernels + D2H inside the timed region)
  roofline   the workload's bound: HBM bytes for Stolt/filters, (sample, trace) pairs for
             Kirchhoff, complex MACs for phase shift
  cpu_baseline  the oracle's reference-cost port (oracle/*_loops) on a bounded sample, 1 host core
  --impl reference   the same CPU port on all host cores (the reference is pure Python/numpy and cannot
             travel to the GPU box; the oracle is pinned to it by tests/golden)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VEL_K = 1.69e8
VEL_S = 1.68e8


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock / throttle-reason sampling DURING the timed region (NVML, 5 ms period; the nvidia-smi
    query line of B200_PROFILING.md reads the same counters)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm = []
        self.reasons = set()
        self.smmax = None
        self.stop_flag = False
        self.collect = False   # the thread starts (and initialises NVML) before the warm-up; samples count only
        self.th = None         # while the timed region runs
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if visible:
                try:
                    idx = int(visible.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.smmax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self.stop_flag:
                if self.collect:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for n, bit in names.items():
                        if r & bit:
                            self.reasons.add(n)
                    time.sleep(0.001)
                else:
                    time.sleep(0.0005)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smmax,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------ per-kernel timing
def kernel_times(names):
    """{kernel: (avg ms per launch, launches)} from the library's CUDA-event brackets (recorded on the stream
    the kernel is launched on, inside the timed region)."""
    import ctypes
    from impdar_b200 import _lib
    lib = _lib.load()
    out = {}
    for n in names:
        ms, cnt = ctypes.c_double(0.0), ctypes.c_int(0)
        _lib.check(lib.impdar_b200_kernel_timer_read(n.encode(), ctypes.byref(ms), ctypes.byref(cnt)))
        if cnt.value:
            out[n] = (ms.value / cnt.value, cnt.value)
    return out


def ncu_traffic(kernel, kernel_ms=None):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/traffic.json,
    written by scripts/ncu_summary.py), or None.  The capture is of ONE launch shape per kernel (the workload named
    in scripts/gpu_round.sh): when the launch measured here lasts very differently from the captured one it is a
    different shape and the figure does not apply."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get(kernel)
    if not e:
        return None
    if kernel_ms and e.get("ncu_ms") and not (0.6 < e["ncu_ms"] / kernel_ms < 1.6):
        return None
    return e.get("dram_bytes")


def hbm_kernel_roofline(wl, table, ms_step, bytes_per_sample_step, hbm_gbs):
    """HBM-bound paths: `table` = {kernel: algorithmic bytes per launch}.  The dominant kernel (largest total time
    in the timed region) gives achieved/peak/frac; the whole step on SURVEY 8d's byte model is reported beside it."""
    kt = kernel_times(list(table))
    step_gbs = wl.units * bytes_per_sample_step / (ms_step * 1e-3) / 1e9
    r = {"bound": "hbm", "peak": hbm_gbs, "unit": "GB/s", "step_bytes_per_sample": bytes_per_sample_step,
         "step_achieved": step_gbs, "step_frac": step_gbs / hbm_gbs}
    if not kt:
        r.update({"achieved": step_gbs, "frac": step_gbs / hbm_gbs, "traffic": None, "kernel": "whole step"})
        return r
    dom = max(kt, key=lambda k: kt[k][0] * kt[k][1])
    ms, cnt = kt[dom]
    gbs = table[dom] / (ms * 1e-3) / 1e9
    r.update({"achieved": gbs, "frac": gbs / hbm_gbs, "kernel": dom, "kernel_ms": ms, "kernel_launches": cnt,
              "algorithmic_bytes_per_launch": table[dom], "traffic": ncu_traffic(dom, ms),
              "kernel_share_of_step": ms * cnt / (ms_step * wl.args.steps),
              "kernels": {k: {"ms": v[0], "launches": v[1], "GBps": table[k] / (v[0] * 1e-3) / 1e9} for k, v in kt.items()}})
    return r


# --------------------------------------------------------------------------------------- workloads
class Workload(object):
    name = ""
    dtype = "f32"
    scaling = "weak"

    def __init__(self, args, rank, world):
        self.args, self.rank, self.world = args, rank, world

    def l2_note(self):
        return "L2 flushed between timed steps (256 MiB write)"


class KirchhoffC2(Workload):
    """configs[1]: Kirchhoff, 4096 traces x 2048 samples, v = 1.69e8; one independent radargram per GPU."""
    name = "kirchhoff_4096tr_x_2048smp_v1.69e8"
    e2e_api = ("impdar_b200.RadarData.migrate(mtype='kirch') on host numpy data (pinned input): "
               "impdar_kirchhoff_host_pipelined_f64, 8 row chunks, upload | kernels | float64 download overlapped")
    S, T = 2048, 4096
    nearfield = False

    def setup(self):
        import torch
        from impdar_b200 import synthetic
        self.tt, self.dist, self.trace_int = synthetic.geometry(self.S, self.T)
        self.x = synthetic.diffractor_radargram(self.S, self.T, seed=2 + self.rank, n_diffractors=64)
        self.host = torch.empty((self.S, self.T), dtype=torch.float32).pin_memory()
        self.host.copy_(self.x)
        self.out = torch.empty((self.S, self.T), dtype=torch.float32, device="cuda")
        self.units = self.S * self.T
        from impdar_b200 import migrationlib as ml
        ml.enable_kirchhoff_stats(True)
        ml.kirchhoff_device(self.x, self.tt, self.dist, VEL_K, self.nearfield, out=self.out)
        torch.cuda.synchronize()
        self.pairs, self.exact_pairs = ml.kirchhoff_stats()
        ml.enable_kirchhoff_stats(False)

    def step(self):
        from impdar_b200 import migrationlib as ml
        ml.kirchhoff_device(self.x, self.tt, self.dist, VEL_K, self.nearfield, out=self.out)

    def e2e_step(self):
        import impdar_b200
        d = impdar_b200.RadarData(self.host.numpy(), dt=1e-8, travel_time=self.tt, dist=self.dist,
                                  trace_int=self.trace_int)
        with _quiet():
            d.migrate(mtype='kirch', vel=VEL_K, nearfield=self.nearfield)
        return self.S * self.T * 4, d.data.nbytes, float(d.data[self.S // 2, self.T // 2])

    def roofline(self, ms, hbm_gbs, src):
        # Kirchhoff is not HBM-bound.  Unit of work = one (output sample, in-aperture input trace)
        # pair.  The uniform-geometry table kernel issues exactly one 4-byte L1 load per pair as contiguous
        # (generally misaligned -> two wavefronts) 128-byte warp loads, so its bound is the L1 load path:
        # 148 SM x 32 lanes x 1.965 GHz / 2 = 4.65e12 pair/s.  SURVEY's issue-model ceiling (4.0e12) and the
        # compulsory HBM traffic (8 B/sample) are reported beside it.
        from impdar_b200 import migrationlib as ml
        path = ml.kirchhoff_last_path()
        kname = "kirch_table_kernel" if path == "table" else "kirch_general_kernel"
        kt = kernel_times([kname])
        kms, kcnt = kt.get(kname, (ms, self.args.steps))
        pairs_s = self.pairs / (kms * 1e-3)
        hbm = self.units * 8 / (ms * 1e-3) / 1e9
        peak = 148 * 32 * 1.965e9 / 2.0 if path == "table" else 148 * 128 * 1.965e9 / 18.0
        return {"bound": "l1_load_wavefronts" if path == "table" else "sm_issue", "achieved": pairs_s, "peak": peak,
                "unit": "pair/s", "frac": pairs_s / peak, "traffic": ncu_traffic(kname, kms),
                "kernel": kname, "kernel_ms": kms, "kernel_launches": kcnt,
                "kernel_share_of_step": kms * kcnt / (ms * self.args.steps),
                "step_pairs_per_s": self.pairs / (ms * 1e-3),
                "peak_model": ("one misaligned 128B L1 load per 32 pairs = 2 wavefronts/SM/clk" if path == "table"
                               else "18 issue slots per pair (measured SASS: 23)"),
                "survey_issue_model_peak": 4.0e12, "frac_of_survey_model": pairs_s / 4.0e12,
                "pairs_per_launch": self.pairs, "exact_fp64_pairs": self.exact_pairs,
                "hbm_compulsory_gbs": hbm, "hbm_frac_of_%s_peak" % src: hbm / hbm_gbs}

    def cpu_sample(self, n_samples, xi=None):
        """Reference-cost port on n_samples output samples of one output trace; returns seconds."""
        from oracle import migration as om
        x64 = self.x.double().cpu().numpy()
        xi = self.T // 2 if xi is None else xi
        ti = list(np.linspace(0, self.S - 1, n_samples).astype(int))
        t0 = time.perf_counter()
        om.kirchhoff_loops(x64, self.tt, self.dist, VEL_K, self.nearfield, xi_list=[xi], ti_list=ti)
        return time.perf_counter() - t0, n_samples

    cpu_sample_desc = "oracle.migration.kirchhoff_loops: %d output samples of trace tnum/2 against the full 2048x4096 input"
    cpu_default_n = 160


class KirchhoffC5(KirchhoffC2):
    """configs[4]: 65536 traces x 8192 samples, output-trace ranges sharded over the ranks, NCCL broadcast of the
    input and all_gather of the output blocks (strong scaling)."""
    name = "kirchhoff_65536tr_x_8192smp_sharded"
    S, T = 8192, 65536
    scaling = "strong"
    cpu_default_n = 4
    cpu_sample_desc = "oracle.migration.kirchhoff_loops: %d output samples of trace tnum/2 against the full 8192x65536 input"

    def setup(self):
        import torch
        import torch.distributed as dist
        from impdar_b200 import synthetic, parallel
        self.tt, self.dist, self.trace_int = synthetic.geometry(self.S, self.T)
        if self.rank == 0:
            self.x = synthetic.diffractor_radargram(self.S, self.T, seed=5, n_diffractors=1024)
        else:
            self.x = torch.empty((self.S, self.T), dtype=torch.float32, device="cuda")
        self.units = self.S * self.T / self.world   # per rank share; value is whole-job
        self.parallel = parallel
        self.xb, self.xe = parallel.kirchhoff_output_range(self.T, self.rank, self.world, self.tt, self.dist, VEL_K)
        self.host = None
        self.count_pairs()

    def step(self):
        kw = {}
        if os.environ.get("IMPDAR_C5_CHUNKS"):          # development A/B switch for the exchange pipeline depth
            kw["pipeline_chunks"] = int(os.environ["IMPDAR_C5_CHUNKS"])
        self.result = self.parallel.kirchhoff_sharded_device(self.x, self.tt, self.dist, VEL_K, False,
                                                             rank=self.rank, world=self.world, gather=True, **kw)

    def e2e_step(self):
        return None

    def count_pairs(self):
        """(sample, trace) pairs inside the aperture, counted by the kernel itself over this rank's output range and
        summed over ranks (one extra untimed step)."""
        import torch
        import torch.distributed as dist
        from impdar_b200 import migrationlib as ml
        ml.enable_kirchhoff_stats(True)
        self.parallel.kirchhoff_sharded_device(self.x, self.tt, self.dist, VEL_K, False, rank=self.rank,
                                               world=self.world, gather=False)
        torch.cuda.synchronize()
        pairs, exact = ml.kirchhoff_stats()
        ml.enable_kirchhoff_stats(False)
        t = torch.tensor([float(pairs), float(exact)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t)
        self.pairs, self.exact_pairs = float(t[0].item()), float(t[1].item())

    def roofline(self, ms, hbm_gbs, src):
        from impdar_b200 import migrationlib as ml
        path = ml.kirchhoff_last_path()
        kname = "kirch_table_kernel" if path == "table" else "kirch_general_kernel"
        kt = kernel_times([kname])
        kms, kcnt = kt.get(kname, (ms, self.args.steps))
        peak1 = 148 * 32 * 1.965e9 / 2.0 if path == "table" else 148 * 128 * 1.965e9 / 18.0
        # rank 0's kernel time per step (the exchange pipeline launches the kernel once per row chunk) with rank 0's
        # share of the pairs (ranges are balanced by pair count)
        kms_step = kms * kcnt / self.args.steps
        pairs_s = self.pairs / self.world / (kms_step * 1e-3)
        return {"bound": "l1_load_wavefronts" if path == "table" else "sm_issue", "achieved": pairs_s, "peak": peak1,
                "unit": "pair/s", "frac": pairs_s / peak1, "traffic": None, "kernel": kname,
                "kernel_ms": kms, "kernel_ms_per_step": kms_step, "kernel_launches": kcnt,
                "kernel_share_of_step": kms * kcnt / (ms * self.args.steps),
                "pairs_whole_image": self.pairs, "exact_fp64_pairs": self.exact_pairs,
                "step_pairs_per_s_all_ranks": self.pairs / (ms * 1e-3),
                "note": "achieved/peak are per GPU (rank 0's kernel, 1/world of the pairs); the step adds the exposed part "
                        "of the NCCL broadcast of the input and the all_gather of the output blocks (overlapped with "
                        "the kernels in bottom-up row chunks)"}


class StoltC5(Workload):
    """The north-star target shape for Stolt: 65536 traces x 8192 samples on one GPU (replicas for N > 1)."""
    name = "stolt_65536tr_x_8192smp"
    S, T = 8192, 65536

    def setup(self):
        import torch
        from impdar_b200 import synthetic
        self.tt, self.dist, self.trace_int = synthetic.geometry(self.S, self.T)
        self.x = synthetic.diffractor_radargram(self.S, self.T, seed=5 + self.rank, n_diffractors=256)
        self.out = torch.empty((1, 2 * (self.S // 2), self.T), dtype=torch.float32, device="cuda")
        self.units = self.S * self.T
        self.host = torch.empty((self.S, self.T), dtype=torch.float32).pin_memory()
        self.host.copy_(self.x)

    def l2_note(self):
        return "inputs (2 GiB) larger than L2; no flush needed"

    def step(self):
        from impdar_b200 import migrationlib as ml
        ml.stolt_device(self.x, 1e-8, 5.0, VEL_S, 10, 10, out=self.out)

    def e2e_step(self):
        import impdar_b200
        d = impdar_b200.RadarData(self.host.numpy(), dt=1e-8, travel_time=self.tt, dist=self.dist,
                                  trace_int=self.trace_int)
        with _quiet():
            d.migrate(mtype='stolt', vel=VEL_S, htaper=10, vtaper=10)
        return self.S * self.T * 4, d.data.nbytes, float(d.data[self.S // 2, self.T // 2])

    def roofline(self, ms, hbm_gbs, src):
        # 40 B per real sample = five sweeps that each read and write the (paired-trace complex)
        # image once: rowA fwd (+taper), rowB fwd (+transpose), col (FFT_t, remap, iFFT_t), rowB inv, rowA inv.
        # Every pass kernel therefore moves 8 B per real sample per launch (algorithmic bytes).
        per = float(self.units) * 8
        table = {"stolt_col_kernel": per, "stolt_rowA_kernel": per, "stolt_rowB_kernel": per,
                 "stolt_remap_paired_kernel": per, "stolt_remap_kernel": per}
        r = hbm_kernel_roofline(self, table, ms, 40, hbm_gbs)
        r["compulsory_8B_gbs"] = self.units * 8 / (ms * 1e-3) / 1e9
        return r

    def cpu_sample(self, n, xi=None):
        from oracle import migration as om
        S, T = 1024, 2048   # bounded sample: same per-cell cost (two FITPACK point evaluations per (kz, kx) cell)
        img = self.x if self.x.dim() == 2 else self.x[0]
        x64 = img[:S, :T].double().cpu().numpy()
        stride = max(1, 512 // n)
        rows = len(range(0, S // 2, stride))
        t0 = time.perf_counter()
        om.stolt_loops(x64, 1e-8, np.ones(T) * 5.0, np.arange(T) * 0.005, VEL_S, 10, 10, row_stride=stride)
        return time.perf_counter() - t0, S * T * rows / 512.0

    cpu_sample_desc = "oracle.migration.stolt_loops (FITPACK point evaluation per cell) on ~%d of the 512 kz rows of a 1024x2048 crop"
    cpu_default_n = 256


class StoltC4(StoltC5):
    """One configs[3] profile (8192 traces x 2048 samples) through Stolt."""
    name = "stolt_8192tr_x_2048smp"
    S, T = 2048, 8192

    def l2_note(self):
        return "L2 flushed between timed steps (256 MiB write)"


class PipelineC4(Workload):
    """configs[3]: vertical_band_pass(2,10) + hfilt(0,T) + Stolt over profiles of 8192 traces x 2048 samples,
    profiles sharded round-robin over GPUs; one step = `--profiles` profiles per GPU (device resident)."""
    name = "pipeline_vbp_hfilt_stolt_8192tr_x_2048smp"
    e2e_api = "impdar_b200.process.process(dats, vbp=(2,10), hfilt=(0,T), migrate=True) on host numpy profiles (pinned)"
    S, T = 2048, 8192

    def setup(self):
        import torch
        from impdar_b200 import synthetic
        self.P = self.args.profiles
        self.tt, self.dist, self.trace_int = synthetic.geometry(self.S, self.T)
        self.x = torch.stack([synthetic.diffractor_radargram(self.S, self.T, seed=4 + self.rank * self.P + p,
                                                             n_diffractors=16) for p in range(self.P)])
        self.units = self.P * self.S * self.T
        from scipy.signal import butter
        nyq = 0.5 / 1e-8
        self.b, self.a = butter(5, [2e6 / nyq, 10e6 / nyq], 'bandpass')
        self.taper = np.exp(-self.tt * 0.05) / np.exp(-self.tt[0] * 0.05)
        self.host = torch.empty((self.P, self.S, self.T), dtype=torch.float32).pin_memory()
        self.host.copy_(self.x)
        self.e2e_units = self.P * self.S * self.T

    def l2_note(self):
        return "inputs larger than L2" if self.P * self.S * self.T * 4 > 126e6 else Workload.l2_note(self)

    def step(self):
        from impdar_b200 import filtering as fl, migrationlib as ml
        y = fl.filtfilt_device(self.x, 'f32', self.b, self.a)
        y = fl.horizontalfilt_device(y, 'f32', self.taper, 0, self.T)
        self.out = ml.stolt_device(y, 1e-8, 5.0, VEL_S, 10, 10)

    def e2e_step(self):
        # the call a user of the reference makes for this configuration: process(dats, vbp=(2, 10), hfilt=(0, T),
        # migrate=True) (lib/process.py:151-193) on HOST arrays; impdar_b200.process.process uploads each profile
        # once, runs the three steps device resident and downloads once, several profiles in flight
        import impdar_b200
        h = self.host.numpy()
        dats = [impdar_b200.RadarData(h[p], dt=1e-8, travel_time=self.tt, dist=self.dist, trace_int=self.trace_int)
                for p in range(self.P)]
        with _quiet():
            impdar_b200.process.process(dats, vbp=(2, 10), hfilt=(0, self.T), migrate=True)
        return self.P * self.S * self.T * 4, sum(d.data.nbytes for d in dats), float(dats[-1].data[self.S // 2, self.T // 2])

    def roofline(self, ms, hbm_gbs, src):
        # 8 + 8 + 40 B/sample unfused; every kernel of the step moves 8 B/sample algorithmically
        per = float(self.units) * 8
        table = {"filtfilt_kernel": per, "hfilt_kernel": per, "stolt_col_kernel": per, "stolt_rowA_kernel": per,
                 "stolt_rowB_kernel": per, "stolt_remap_paired_kernel": per}
        return hbm_kernel_roofline(self, table, ms, 56, hbm_gbs)

    def cpu_sample(self, n, xi=None):
        from oracle import filtering as of
        x64 = self.x[0].double().cpu().numpy()
        t0 = time.perf_counter()
        y = of.vertical_band_pass(x64, 1e-8, 2, 10)
        of.horizontalfilt(y, self.tt, 0, self.T)
        t_f = time.perf_counter() - t0
        t_s, cells = StoltC5.cpu_sample(self, n)
        # Stolt cost scales with cells: extrapolate it to one full profile, add the measured filters, and report
        # the measured time with the equivalent number of fully processed samples
        t_full = t_f + t_s * (self.S * self.T) / cells
        elapsed = t_f + t_s
        return elapsed, self.S * self.T * elapsed / t_full

    cpu_sample_desc = ("oracle vertical_band_pass + horizontalfilt on one full profile (measured) + stolt_loops on ~%d of the "
                       "512 kz rows of a 1024x2048 crop extrapolated by cell count to the profile")
    cpu_default_n = 128


class PhshC3(Workload):
    """configs[2]: phase-shift migration, 16384 traces x 4096 samples; constant velocity (default) or layered."""
    name = "phsh_const_16384tr_x_4096smp"
    S, T = 4096, 16384
    layered = False

    def setup(self):
        import torch
        from impdar_b200 import synthetic, migrationlib as ml
        self.tt, self.dist, self.trace_int = synthetic.geometry(self.S, self.T)
        self.x = synthetic.diffractor_radargram(self.S, self.T, seed=3 + self.rank, n_diffractors=128)
        self.out = torch.empty((self.S, self.T), dtype=torch.float32, device="cuda")
        self.units = self.S * self.T
        self.host = torch.empty((self.S, self.T), dtype=torch.float32).pin_memory()
        self.host.copy_(self.x)
        self.vel_table = synthetic.layered_velocity(self.tt)
        d = type("D", (), {})()
        d.travel_time, d.snum, d.tnum, d.dist = self.tt, self.S, self.T, self.dist
        self.vmig = ml.getVelocityProfile(d, self.vel_table) if self.layered else VEL_K

    def l2_note(self):
        return "inputs (256 MiB) larger than L2; no flush needed"

    def step(self):
        from impdar_b200 import migrationlib as ml
        ml.phase_shift_device(self.x, 1e-8, 5.0, self.tt, self.vmig, 10, 10, out=self.out)

    def e2e_step(self):
        import impdar_b200
        from impdar_b200 import migrationlib as ml
        d = impdar_b200.RadarData(self.host.numpy(), dt=1e-8, travel_time=self.tt, dist=self.dist,
                                  trace_int=self.trace_int)
        with _quiet():
            ml.migrationPhaseShift(d, vel=self.vel_table if self.layered else VEL_K, htaper=10, vtaper=10)
        return self.S * self.T * 4, d.data.nbytes, float(d.data[self.S // 2, self.T // 2])

    def roofline(self, ms, hbm_gbs, src):
        nt = 1 << (self.S - 1).bit_length()
        K = self.T // 2 + 1
        macs = float(nt) * K * self.S   # complex MACs over (w, kx >= 0, tau)
        peak = 148 * 128 * 1.965e9 / 6.0   # 6 FP32 issue slots per complex multiply-accumulate
        if self.layered:
            peak = 148 * 16 * 1.965e9 / 3.0   # MUFU bound: rsqrt + sin + cos per (tau, w, k)
        kname = "phsh_layered_pair_kernel" if self.layered else "phsh_const_pair_kernel"
        kt = kernel_times([kname])
        kms, kcnt = kt.get(kname, (ms, self.args.steps))
        return {"bound": "fp32_simt" if not self.layered else "mufu", "achieved": macs / (kms * 1e-3), "peak": peak,
                "unit": "cmac/s", "frac": macs / (kms * 1e-3) / peak, "traffic": ncu_traffic(kname, kms),
                "kernel": kname, "kernel_ms": kms, "kernel_launches": kcnt,
                "kernel_share_of_step": kms * kcnt / (ms * self.args.steps),
                "cmacs_per_launch": macs}

    def cpu_sample(self, n, xi=None):
        from oracle import migration as om
        S, T = self.S, 256   # bounded: all frequencies, a 256-trace crop, n output taus
        x64 = self.x[:, :T].double().cpu().numpy()
        tap = om.phsh_taper(x64, 10, 10)
        nt, kx, ws, FK = om.phase_shift_spectrum(tap, 1e-8, np.ones(T) * 5.0, None)
        t0 = time.perf_counter()
        if self.layered:
            om.phase_shift_layered_tk(FK, kx, ws, 1e-8, self.tt, np.asarray(self.vmig), tau_end=n)
        else:
            om.phase_shift_const_tk(FK, kx, ws, 1e-8, S, VEL_K, tau_end=n)
        return time.perf_counter() - t0, n * T

    cpu_sample_desc = "oracle phase-shift recurrence (vectorised over (w,kx), sequential in tau): %d taus x 256 traces, all frequencies"
    cpu_default_n = 256


class PhshC3Layered(PhshC3):
    name = "phsh_layered_16384tr_x_4096smp"
    layered = True


WORKLOADS = {"kirchhoff": KirchhoffC2, "kirchhoff_c5": KirchhoffC5, "stolt": StoltC5, "stolt_c4": StoltC4,
             "pipeline": PipelineC4, "phsh": PhshC3, "phsh_layered": PhshC3Layered}


class _quiet(object):
    def __enter__(self):
        self._so = sys.stdout
        sys.stdout = open(os.devnull, "w")

    def __exit__(self, *a):
        sys.stdout.close()
        sys.stdout = self._so


# ------------------------------------------------------------------------------------- reference arm
_REF_WL = None


def _ref_job(xi):
    wl, n = _REF_WL
    return wl.cpu_sample(n, xi)


def cpu_only_setup(wl):
    """Build the workload's synthetic input on the host (no CUDA) for the CPU arms."""
    from impdar_b200 import synthetic
    wl.tt, wl.dist, wl.trace_int = synthetic.geometry(wl.S, wl.T)
    nd = {"kirchhoff": 64}.get(wl.args.workload, 64)
    S, T = wl.S, wl.T
    if S * T > (1 << 26):   # the CPU sample only touches a crop; keep generation bounded
        T = 4096
    x = synthetic.diffractor_radargram(S, T, seed=2, n_diffractors=nd, device="cpu")
    wl.x = x
    if T != wl.T:
        wl.T_full, wl.T = wl.T, T
        wl.tt, wl.dist, wl.trace_int = synthetic.geometry(wl.S, wl.T)
    if isinstance(wl, PipelineC4):
        wl.x = x[None]
    if isinstance(wl, PhshC3):
        from impdar_b200 import migrationlib as ml
        wl.vel_table = synthetic.layered_velocity(wl.tt)
        d = type("D", (), {})()
        d.travel_time, d.snum, d.tnum, d.dist = wl.tt, wl.S, wl.T, wl.dist
        wl.vmig = ml.getVelocityProfile(d, wl.vel_table) if wl.layered else VEL_K


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (the oracle's reference-cost port) on all host cores."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload](args, 0, 1)
    cpu_only_setup(wl)
    n = max(1, wl.cpu_default_n // 4)
    torch.set_num_threads(1)

    import multiprocessing as mp
    global _REF_WL
    _REF_WL = (wl, n)
    pool = mp.get_context("fork").Pool(cores)   # the reference is single-threaded Python: one process per core

    def one_step():
        t0 = time.perf_counter()
        res = pool.map(_ref_job, [(wl.T // 2 + 7 * i) % wl.T for i in range(cores)])
        return time.perf_counter() - t0, float(sum(u for _, u in res))

    for _ in range(min(args.warmup, 1)):
        one_step()
    tot_t, tot_u = 0.0, 0.0
    for _ in range(args.steps):
        t, u = one_step()
        tot_t += t
        tot_u += u
    value = tot_u / tot_t
    line = {"impl": "reference", "metric": "migrated samples/s", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3,
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": (wl.cpu_sample_desc % n) + "; %d concurrent samples (processes), one per core" % cores},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kirchhoff", choices=sorted(WORKLOADS))
    ap.add_argument("--profiles", type=int, default=8, help="profiles per GPU per step (pipeline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from impdar_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    hbm_gbs, peak_src, _ = load_peaks()

    wl = WORKLOADS[args.workload](args, rank, world)
    wl.setup()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    need_flush = "flushed" in wl.l2_note()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        wl.step()
    barrier()
    launches0 = lib.impdar_b200_launch_count()
    evs = []
    lib.impdar_b200_kernel_timer(1)   # CUDA-event brackets around the dominant kernels, on their launch stream
    barrier()
    sampler.collect = True
    torch.cuda.profiler.start()   # ncu --profile-from-start off captures only the timed region
    for _ in range(args.steps):
        if need_flush:
            flush.fill_(1)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        wl.step()
        b.record()
        evs.append((a, b))
    barrier()
    sampler.collect = False
    torch.cuda.profiler.stop()
    lib.impdar_b200_kernel_timer(0)   # stop recording; the records stay readable for roofline()
    launches = lib.impdar_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units_all = wl.units * world
    value = units_all / (ms_per_step * 1e-3)

    # ---- end to end through the plugin call with HOST buffers
    e2e = None
    if not args.no_e2e and wl.host is not None:
        for _ in range(2):
            wl.e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h2d, d2h, probe = wl.e2e_step()
        torch.cuda.synchronize()
        dt_e2e = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_units = getattr(wl, "e2e_units", wl.units)
        e2e = {"value": e2e_units * world / float(te.item()), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(te.item()) * 1e3,
               "api": getattr(wl, "e2e_api", "impdar_b200.RadarData hot-path methods on host numpy data (pinned input)")}

    if rank == 0:
        line = {"metric": "migrated samples/s", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": {"workload": wl.name, "snum": wl.S, "tnum": wl.T, "l2": wl.l2_note(),
                           "parallelism": "1 process per GPU, %s" % ("independent radargrams per GPU, no collective"
                                                                     if wl.scaling == "weak" else
                                                                     "output-trace ranges per GPU, NCCL broadcast + all_gather")},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
                "roofline": wl.roofline(ms_per_step, hbm_gbs, peak_src), "peak_source": peak_src}
        if world == 1 and not args.no_cpu_baseline:
            n = args.cpu_samples or wl.cpu_default_n
            secs, units = wl.cpu_sample(n)
            line["cpu_baseline"] = {"value": units / secs, "unit": "samples/s", "cores": 1, "kind": "port",
                                    "sample": (wl.cpu_sample_desc % n) + "; %.1f s" % secs,
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
