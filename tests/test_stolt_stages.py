"""The five-pass Stolt pipeline (impdar_b200/csrc/stolt_fft.cu), stage by stage.

CPU: the numpy stage model (tests/stolt_stage_model.py) reproduces the oracle, i.e. the decomposition is exact.
GPU: every intermediate buffer of the kernels matches the model, and the full result matches the oracle and the
cuFFT pipelines within the north-star tolerance (relative L2 <= 1e-5)."""
import numpy as np
import pytest

import stolt_stage_model as model
from oracle import migration as om


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


DT, DX, VEL = 1e-8, 5.0, 1.68e8


def _input(S, T, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((S, T)).astype(np.float32)


@pytest.mark.parametrize("S,T,N2", [(64, 128, 16), (32, 256, 32), (128, 64, 8)])
def test_model_matches_oracle(S, T, N2):
    x = _input(S, T, 1).astype(np.float64)
    tap, want = om.stolt(x, DT, np.ones(T) * DX, np.arange(T) * DX / 1e3, VEL, 5, 7)
    got = model.full(tap, DT, DX, VEL, N2)
    assert rel_l2(got, want) < 1e-12


def _run_stage(x, stage, htaper, vtaper):
    import torch
    from impdar_b200 import _lib, device, migrationlib as ml
    lib = _lib.load()
    S, T = x.shape
    xd = torch.from_numpy(x).cuda()
    out = torch.zeros((1, S, T), dtype=torch.float32, device="cuda")
    ml.set_stolt_pipeline(ml.STOLT_FIVE_PASS)
    _lib.check(lib.impdar_stolt_debug_stop_after(stage))
    try:
        ml.stolt_device(xd, DT, DX, VEL, htaper, vtaper, out=out)
        torch.cuda.synchronize()
        assert ml.stolt_last_pipeline() == 'five_pass'
    finally:
        _lib.check(lib.impdar_stolt_debug_stop_after(0))
        ml.set_stolt_pipeline(ml.STOLT_AUTO)
    ws = device.workspace(1)
    w1 = out[0].cpu().numpy().view(np.complex64).reshape(S, T // 2)
    n = S * T // 2
    w2 = ws[:n * 8].cpu().numpy().view(np.complex64).reshape(T // 2, S)
    return out[0].cpu().numpy(), w1, w2


@pytest.mark.gpu
@pytest.mark.parametrize("S,T", [(1024, 8192), (2048, 16384), (1024, 32768)])
def test_stages_match_model(S, T):
    x = _input(S, T, S + T)
    ht, vt = 10, 20
    tap = om.stolt_taper(x.astype(np.float64), ht, vt)
    bu = model.beta_unit(S, T, DT, DX, VEL)
    m1 = model.p1(tap)
    m2 = model.p2(m1, T)
    m3 = model.p3(m2, T, bu)
    m4 = model.p4(m3, T)
    m5 = model.p5(m4, T)
    errs = {}
    _, w1, _ = _run_stage(x, 1, ht, vt)
    errs[1] = rel_l2(w1, m1)
    _, _, w2 = _run_stage(x, 2, ht, vt)
    errs[2] = rel_l2(w2, m2)
    _, _, w2 = _run_stage(x, 3, ht, vt)
    errs[3] = rel_l2(w2, m3)
    _, w1, _ = _run_stage(x, 4, ht, vt)
    errs[4] = rel_l2(w1, m4)
    o, _, _ = _run_stage(x, 0, ht, vt)
    errs[5] = rel_l2(o, m5)
    print("stolt five-pass stage errors %dx%d:" % (S, T), {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 1e-5, "stage %d: %.3e" % (k, v)


@pytest.mark.gpu
@pytest.mark.parametrize("S,T", [(1024, 8192), (4096, 8192), (8192, 8192), (2048, 65536), (1024, 131072)])
def test_five_pass_vs_oracle(S, T):
    import torch
    from impdar_b200 import migrationlib as ml
    x = _input(S, T, 7 * S + T)
    _, want = om.stolt(x.astype(np.float64), DT, np.ones(T) * DX, np.arange(T) * DX / 1e3, VEL, 10, 20)
    xd = torch.from_numpy(x).cuda()
    got = ml.stolt_device(xd, DT, DX, VEL, 10, 20).cpu().numpy()
    assert ml.stolt_last_pipeline() == 'five_pass'
    e = rel_l2(got, want)
    ml.set_stolt_pipeline(ml.STOLT_CUFFT_PAIRED)
    try:
        ref = ml.stolt_device(xd, DT, DX, VEL, 10, 20).cpu().numpy()
        assert ml.stolt_last_pipeline() == 'cufft_paired'
    finally:
        ml.set_stolt_pipeline(ml.STOLT_AUTO)
    e2 = rel_l2(ref, want)
    print("stolt %dx%d: five-pass %.2e, cuFFT paired %.2e (rel L2 vs float64 oracle), max abs %.2e"
          % (S, T, e, e2, np.abs(got - want).max()))
    assert e < 1e-5


@pytest.mark.gpu
def test_five_pass_batched_equals_single():
    """A stack of profiles goes through the five passes in one set of launches (grid.z / column index over the
    batch); the result must be bit-identical to running the profiles one by one."""
    import torch
    from impdar_b200 import migrationlib as ml
    S, T, B = 1024, 8192, 3
    x = torch.from_numpy(np.stack([_input(S, T, 100 + b) for b in range(B)])).cuda()
    got = ml.stolt_device(x, DT, DX, VEL, 10, 20)
    assert ml.stolt_last_pipeline() == 'five_pass'
    for b in range(B):
        one = ml.stolt_device(x[b].contiguous(), DT, DX, VEL, 10, 20)
        assert torch.equal(got[b], one)
