"""Shared helpers for the test-suite."""
import numpy as np

import impdar_b200


def dat_from_golden(g, dtype=None):
    """impdar_b200.RadarData built from a golden fixture's inputs."""
    data = g["data"] if dtype is None else g["data"].astype(dtype)
    d = impdar_b200.RadarData(data.copy(), dt=float(g["dt"]), travel_time=g["travel_time"].copy(),
                              dist=g["dist"].copy() if "dist" in g else None,
                              trace_int=g["trace_int"].copy() if "trace_int" in g else None)
    return d


def synthetic_dat(S, T, seed=0, dt=1e-8, dx=5.0, tt0_us=0.0, dtype=np.float32, kind="noise"):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((S, T)).astype(dtype)
    return impdar_b200.RadarData(data, dt=dt, travel_time=tt0_us + np.arange(S) * dt * 1e6,
                                 dist=np.arange(T) * dx / 1e3, trace_int=np.ones(T) * dx)
