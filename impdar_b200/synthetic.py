"""Seeded synthetic radargrams of the BASELINE.json shapes (SURVEY.md 8d "common synthetic geometry").

dt = 1e-8 s, travel_time[i] = i * 0.01 us, dist[x] = 0.005 x km, trace_int = 5 m, fp32 data:
point diffractors rendered as Ricker wavelets along t = sqrt(t0^2 + (2 dx / v)^2) plus white noise.
Generation uses torch (on the GPU when there is one) purely as an array library; it is bench/test input,
not part of the migration path.
"""
import numpy as np

DT = 1.0e-8
DX = 5.0
VEL = 1.69e8


def geometry(snum, tnum, dt=DT, dx=DX):
    """(travel_time [us], dist [km], trace_int [m])."""
    return np.arange(snum) * dt * 1e6, np.arange(tnum) * dx / 1e3, np.ones(tnum) * dx


def diffractor_radargram(snum, tnum, seed, n_diffractors=64, f0=5.0e6, noise=0.05, dt=DT, dx=DX, vel=VEL,
                         device=None):
    """(snum, tnum) float32 torch tensor on ``device`` (default: cuda if available)."""
    import torch
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    rng = np.random.default_rng(seed)
    ax = rng.uniform(0, tnum * dx, n_diffractors)
    at = rng.uniform(0.1, 0.9, n_diffractors) * snum * dt
    amp = rng.uniform(0.5, 1.5, n_diffractors) * rng.choice([-1.0, 1.0], n_diffractors)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    out = noise * torch.randn((snum, tnum), generator=g, device=device, dtype=torch.float32)
    t = (torch.arange(snum, device=device, dtype=torch.float32) * dt)[:, None]
    x = (torch.arange(tnum, device=device, dtype=torch.float32) * dx)[None, :]
    half = 1.5 / f0  # wavelet support
    for i in range(n_diffractors):
        # only the columns where the hyperbola is inside the record
        amax = np.sqrt(max((snum * dt) ** 2 - at[i] ** 2, 0.0)) * vel / 2.0
        c0 = max(int((ax[i] - amax) / dx), 0)
        c1 = min(int((ax[i] + amax) / dx) + 2, tnum)
        if c1 <= c0:
            continue
        th = torch.sqrt(at[i] ** 2 + (2.0 * (x[:, c0:c1] - ax[i]) / vel) ** 2)
        arg = (np.pi * f0) ** 2 * (t - th) ** 2
        w = torch.where((t - th).abs() < half, (1.0 - 2.0 * arg) * torch.exp(-arg), torch.zeros((), device=device))
        out[:, c0:c1] += float(amp[i]) * w
    return out


def layered_velocity(travel_time_us, n=40, v0=1.69e8, v1=2.3e8):
    """C3's firn-like (v, z) table: passes getVelocityProfile's coverage guards."""
    zmax = v1 * travel_time_us[-1] * 1e-6 / 2
    z = np.linspace(0, 1.05 * zmax, n)
    v = v0 + (v1 - v0) * np.exp(-z / (0.15 * zmax))
    return np.stack([v, z], 1)
