"""Accuracy evidence for the tensor-core form of the constant-velocity phase shift.

TK[tau, k] = 1/S sum_w FK[w, k] z^(tau + 1),  z = exp(i phi(w, k))                      (mig_python.py:396-420)
With tau + 1 = t0 B + j + 1 the sum is, per kx, a dense complex matrix product over w:
    TK[t0 B + j, k] = sum_w A[t0, w] Bm[w, j],   A[t0, w] = FK[w, k] (z^B)^t0,   Bm[w, j] = z^(j + 1)
(M = S / B rows, N = B columns, K = nt frequencies).  This script evaluates that product the way tcgen05.mma
kind::tf32 would - operands rounded to TF32 (10-bit mantissa), products accumulated in float32 - plain and with the
3 x TF32 split (a = a_hi + a_lo; a_hi b_hi + a_hi b_lo + a_lo b_hi), and compares with the float64 oracle on a crop
the oracle can do.   python scripts/phsh_tf32_emulation.py [S T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import migration as om


def tf32(x):
    """Round float32 to TF32 (10 explicit mantissa bits), round-to-nearest-even on the dropped 13 bits."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0xFFF + ((b >> 13) & 1)) & ~np.uint64(0x1FFF)
    return b.astype(np.uint32).view(np.float32)


def mm32(a, b):
    return np.matmul(a.astype(np.float32), b.astype(np.float32))      # float32 accumulation


def cgemm(Ar, Ai, Br, Bi, mode):
    if mode == "fp32":
        return mm32(Ar, Br) - mm32(Ai, Bi), mm32(Ar, Bi) + mm32(Ai, Br)
    split = lambda x: (tf32(x), tf32(x - tf32(x)))
    (Arh, Arl), (Aih, Ail), (Brh, Brl), (Bih, Bil) = split(Ar), split(Ai), split(Br), split(Bi)
    def prod(ah, al, bh, bl):
        p = mm32(ah, bh)
        if mode == "tf32x3":
            p = p + (mm32(ah, bl) + mm32(al, bh))
        return p
    return (prod(Arh, Arl, Brh, Brl) - prod(Aih, Ail, Bih, Bil)), (prod(Arh, Arl, Bih, Bil) + prod(Aih, Ail, Brh, Brl))


S, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 64)
B = 32
rng = np.random.default_rng(31)
x = rng.standard_normal((S, T)).astype(np.float32).astype(np.float64)
dt, dx, vel = 1e-8, 5.0, 1.69e8
tt = np.arange(S) * dt * 1e6
tap = om.phsh_taper(x, 10, 10)
nt, kx, ws, FK = om.phase_shift_spectrum(tap, dt, np.ones(T) * dx, None)
TK = om.phase_shift_const_tk(FK, kx, ws, dt, S, vel)
want = np.fft.ifft(TK).real
w = ws.copy()
w[w == 0.0] = 1e-10 / dt
for mode in ("fp32", "tf32", "tf32x3"):
    got = np.zeros((S, T), dtype=np.complex128)
    for k in range(T):
        vkx2 = (vel * kx[k] / 2.) ** 2.
        prop = vkx2 < w ** 2.
        phi = np.where(prop, w * dt * np.sqrt(np.where(prop, 1.0 - vkx2 / w ** 2., 0.0)), 0.0)      # float64 seeds
        fk = np.where(prop, FK[:, k], 0.0)
        t0 = np.arange(S // B)
        A = fk[None, :] * np.exp(1j * phi[None, :] * (B * t0)[:, None])            # (S/B, nt)
        Bm = np.exp(1j * phi[:, None] * (np.arange(B) + 1)[None, :])               # (nt, B)
        cr, ci = cgemm(A.real, A.imag, Bm.real, Bm.imag, mode)
        got[:, k] = (cr + 1j * ci).reshape(-1) / S
    out = np.fft.ifft(got).real
    print("%-7s rel-L2 of the migrated image vs the float64 oracle: %.3e   (S = %d, nt = %d, T = %d, B = %d)"
          % (mode, np.linalg.norm(out - want) / np.linalg.norm(want), S, nt, T, B))
