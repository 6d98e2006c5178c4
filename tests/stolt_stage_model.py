"""numpy model of the five-pass Stolt pipeline of impdar_b200/csrc/stolt_fft.cu (TEST INFRASTRUCTURE).

It restates, stage by stage and with the same index conventions as the kernels, how the 2-D transform,
the w -> kz remap (mig_python.py:171-200) and the inverse transform are split into five HBM sweeps:

  P1  rows:    taper, z[s][j] = d[s][2j] + i d[s][2j+1] (j = n1*N2 + n2), FFT over n1, twiddle      -> W1[s][k1*N2+n2]
  P2  rows:    FFT over n2, untangle to the half spectrum D[s][kx], transposed store                -> Dt[c][s]
               (memory column c = k1*N2 + k2 holds kx = k1 + N1*k2; column 0 packs kx = 0 and kx = T/2)
  P3  columns: FFT over s, remap + obliquity on both frequency halves, inverse FFT over w (in place) -> Gt[c][s]
  P4  rows:    transposed load, tangle, inverse FFT over k2                                          -> W1[s][k1*N2+n2]
  P5  rows:    twiddle, inverse FFT over k1 -> real image out[s][2j], out[s][2j+1]

tests/test_stolt_stages.py checks the model against the oracle (CPU) and the kernels' intermediate buffers
against the model (GPU).
"""
import numpy as np


def split(T, N2=256):
    Th = T // 2
    assert Th % N2 == 0
    return Th, Th // N2, N2


def p1(tap, N2=256):
    S, T = tap.shape
    Th, N1, N2 = split(T, N2)
    z = tap[:, 0::2] + 1j * tap[:, 1::2]                 # (S, Th)
    z = z.reshape(S, N1, N2)                              # [s][n1][n2]
    Y = np.fft.fft(z, axis=1)                             # over n1 -> k1
    k1 = np.arange(N1)[:, None]
    n2 = np.arange(N2)[None, :]
    Y = Y * np.exp(-2j * np.pi * k1 * n2 / Th)[None]
    return Y.reshape(S, Th)                               # W1[s][k1*N2 + n2]


def col_kx(T, N2=256):
    Th, N1, N2 = split(T, N2)
    c = np.arange(Th)
    return (c // N2) + N1 * (c % N2)


def p2(W1, T, N2=256):
    S = W1.shape[0]
    Th, N1, N2 = split(T, N2)
    Z = np.fft.fft(W1.reshape(S, N1, N2), axis=2)         # [s][k1][k2] = Z[k1 + N1*k2]
    Z = Z.reshape(S, Th)                                   # memory column c = k1*N2 + k2
    kx = col_kx(T, N2)
    # partner column of c: the one holding Th - kx
    inv = np.empty(Th, dtype=np.int64)
    inv[kx] = np.arange(Th)
    pc = inv[(Th - kx) % Th]
    Zp = np.conj(Z[:, pc])
    E = 0.5 * (Z + Zp)
    O = -0.5j * (Z - Zp)
    D = E + np.exp(-2j * np.pi * kx / T)[None, :] * O      # D[s][kx], kx in [0, Th)
    # column 0: pack (D[0], D[Th]) = (Zr + Zi, Zr - Zi)
    z0 = Z[:, 0]
    D[:, 0] = (z0.real + z0.imag) + 1j * (z0.real - z0.imag)
    return np.ascontiguousarray(D.T)                       # Dt[c][s]


def remap_column(F, beta, S, norm):
    """F: full complex spectrum over w (length S) of one kx >= 0 column.  Returns Q (length S)."""
    nz = S // 2
    j = np.arange(1, nz)
    f = np.sqrt(j.astype(np.float64) ** 2 + beta * beta)
    fq = np.minimum(f, float(nz))
    i0 = np.minimum(fq.astype(np.int64), nz - 1)
    a = fq - i0
    sc = j / f * norm
    Q = np.zeros(S, dtype=np.complex128)
    Q[j] = sc * ((1 - a) * F[i0] + a * F[i0 + 1])
    Q[S - j] = sc * ((1 - a) * F[(S - i0) % S] + a * F[S - i0 - 1])
    return Q


def p3(Dt, T, beta_unit, N2=256):
    Th, S = Dt.shape
    kx = col_kx(T, N2)
    norm = 1.0 / (S * T)
    Gt = np.empty_like(Dt)
    for c in range(Th):
        F = np.fft.fft(Dt[c])
        if c == 0:
            Fm = np.conj(np.roll(F[::-1], 1))              # conj(F[-w])
            F0 = 0.5 * (F + Fm)
            FN = -0.5j * (F - Fm)
            Q = remap_column(F0, 0.0, S, norm) + 1j * remap_column(FN, beta_unit * (T // 2), S, norm)
        else:
            Q = remap_column(F, beta_unit * kx[c], S, norm)
        Gt[c] = np.fft.ifft(Q) * S                          # unnormalised inverse
    return Gt


def p4(Gt, T, N2=256):
    Th, S = Gt.shape
    Th, N1, N2 = split(T, N2)
    G = Gt.T                                                # [s][c]
    kx = col_kx(T, N2)
    inv = np.empty(Th, dtype=np.int64)
    inv[kx] = np.arange(Th)
    pc = inv[(Th - kx) % Th]
    Gp = np.conj(G[:, pc])
    e = G + Gp
    o = (G - Gp) * np.exp(+2j * np.pi * kx / T)[None, :]
    Zq = e + 1j * o
    g0 = G[:, 0]
    Zq[:, 0] = (g0.real + g0.imag) + 1j * (g0.real - g0.imag)
    Zq = Zq.reshape(S, N1, N2)                              # [s][k1][k2]
    Y = np.fft.ifft(Zq, axis=2) * N2                        # over k2 -> n2, unnormalised
    return Y.reshape(S, Th)                                 # W1[s][k1*N2 + n2]


def p5(W1, T, N2=256):
    S = W1.shape[0]
    Th, N1, N2 = split(T, N2)
    Y = W1.reshape(S, N1, N2)
    k1 = np.arange(N1)[:, None]
    n2 = np.arange(N2)[None, :]
    Y = Y * np.exp(+2j * np.pi * k1 * n2 / Th)[None]
    z = np.fft.ifft(Y, axis=1) * N1                         # over k1 -> n1
    z = z.reshape(S, Th)
    out = np.empty((S, T))
    out[:, 0::2] = z.real
    out[:, 1::2] = z.imag
    return out


def beta_unit(S, T, dt, dx, vel):
    return vel * S * dt / (2.0 * T * dx)


def full(tap, dt, dx, vel, N2=256):
    S, T = tap.shape
    W1 = p1(tap, N2)
    Dt = p2(W1, T, N2)
    Gt = p3(Dt, T, beta_unit(S, T, dt, dx, vel), N2)
    W1b = p4(Gt, T, N2)
    return p5(W1b, T, N2)
