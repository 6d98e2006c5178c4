"""Evidence for DESIGN.md 4.4: can vertical_band_pass's IIR recurrence (scipy.signal.filtfilt on the (b, a) of a
5th-order band-pass, _RadarDataFiltering.py:528-529) be parallelised in TIME by a chunked scan?

The scan splits the (odd-extended) trace into C chunks, runs every chunk from a zero state, and fixes the chunk's
initial state up with the state-transition matrix: z_in[c+1] = A^L z_in[c] + s_c.  That is exact in exact arithmetic.
In float64 it is not usable in the direct-form basis scipy's lfilter works in: the transposed-direct-form-II state
matrix of a 10th-order band-pass is a companion matrix with entries up to ~250 and strongly non-normal, |A^L| grows to
1e5 .. 1e12 before the poles' decay wins, so a rounding error of 1e-16 in a chunk's initial state comes back amplified
by that transient TWICE (once through A^L, once through the chunk's own recurrence).  The serial recurrence suffers the
transient once (that is the well-known ~1e-9 disagreement between float64 orderings of this filter).

Run on the CPU (no GPU needed):  python scripts/vbp_scan_numerics.py
Columns: relative L2 error of the chunked scan against scipy.signal.lfilter with the same zi (what filtfilt runs), and
of lfilter itself against an 80-bit long-double evaluation of the same recurrence."""
import numpy as np
from scipy.signal import butter, cheby1, lfilter, lfilter_zi


def run(b, a, x, z, dtype=np.float64):
    """Transposed direct form II, the recurrence of scipy.signal.lfilter."""
    N = len(a) - 1
    y = np.empty(len(x), dtype=dtype)
    z = z.astype(dtype).copy()
    b = b.astype(dtype)
    a = a.astype(dtype)
    for n in range(len(x)):
        yv = b[0] * x[n] + z[0]
        for i in range(N - 1):
            z[i] = b[i + 1] * x[n] - a[i + 1] * yv + z[i + 1]
        z[N - 1] = b[N] * x[n] - a[N] * yv
        y[n] = yv
    return y, z


def transition_power(b, a, L):
    """A^L column by column: the zero-input response of the recurrence itself, in long double (repeated squaring of the
    companion matrix in float64 is useless: catastrophic cancellation)."""
    N = len(a) - 1
    AL = np.zeros((N, N))
    zero = np.zeros(L, dtype=np.longdouble)
    for j in range(N):
        e = np.zeros(N, dtype=np.longdouble)
        e[j] = 1
        AL[:, j] = run(b * 0, a, zero, e, np.longdouble)[1].astype(np.float64)
    return AL


rng = np.random.default_rng(0)
filters = {"butter-5 2-10 MHz at dt = 1e-8 (BASELINE config 4)": butter(5, [0.04, 0.2], 'bandpass'),
           "butter-5 0.5-2 MHz": butter(5, [0.01, 0.04], 'bandpass'),
           "cheby1-5 rp = 5, 2-10 MHz": cheby1(5, 5, [0.04, 0.2], 'bandpass')}
for name, (b, a) in filters.items():
    x = rng.standard_normal(2048 + 66)          # snum + 2 padlen
    zi = lfilter_zi(b, a) * x[0]
    yref, _ = lfilter(b, a, x, zi=zi)
    ytrue, _ = run(b, a, x.astype(np.longdouble), zi, np.longdouble)
    print("%s\n    scipy lfilter (float64) vs long double: %.2e" % (name, np.linalg.norm(yref - ytrue) / np.linalg.norm(ytrue)))
    for C in (4, 8, 16, 32):
        L = -(-len(x) // C)
        AL = transition_power(b, a, L)
        s = [run(b, a, x[c * L:(c + 1) * L], np.zeros(len(a) - 1))[1] for c in range(C)]
        zin = [zi.copy()]
        for c in range(1, C):
            zin.append(AL @ zin[c - 1] + s[c - 1])
        y = np.concatenate([run(b, a, x[c * L:(c + 1) * L], zin[c])[0] for c in range(C)])
        with np.errstate(all='ignore'):
            err = np.linalg.norm(y - yref) / np.linalg.norm(yref)
        print("    %2d chunks of %4d samples: scan vs lfilter %.2e   max |A^L| %.1e" % (C, L, err, np.abs(AL).max()))
