// Stolt f-k migration as five hand-written HBM sweeps (reference: migrationlib/mig_python.py:126-208).
//
// The cuFFT pipeline of stolt.cu moves ~9 sweeps of the radargram through HBM (taper, three transform passes
// forward, remap, three back).  For power-of-two shapes this file does the whole job in five, which is the
// 40 B/sample model of SURVEY.md 8d:
//
//   P1 rows     taper (mig_python.py:152-157) fused into the load; adjacent traces paired into one complex signal
//               z[s][j] = d[s][2j] + i d[s][2j+1], j = n1*N2 + n2 (N2 = 256); FFT over n1; twiddle      -> W1[s][k1*N2+n2]
//   P2 rows     FFT over n2; untangle the paired transform into the half spectrum D[s][kx], kx in [0, T/2);
//               transposed store, so a wavenumber column becomes contiguous in time                     -> Dt[c][s]
//               (memory column c = k1*N2 + k2 holds kx = k1 + N1*k2; column 0 packs the real kx = 0 and T/2 series)
//   P3 columns  one CTA per wavenumber: FFT over s in shared memory, w -> kz remap with complex linear
//               interpolation and obliquity scaling (:171-200) on both frequency halves, inverse FFT       in place
//   P4 rows     transposed load, tangle, inverse FFT over k2                                               -> W1[s][k1*N2+n2]
//   P5 rows     twiddle, inverse FFT over k1, real image written in place of W1 (the caller's `out`)
//
// All transforms are fp32 Stockham / four-step FFTs with radix-16/32 butterflies in registers (fft_regs.cuh)
// and twiddle tables computed in fp64; the remap coordinate sqrt(j^2 + beta^2) is fp64 (one Newton step on
// the fp32 square root).  tests/stolt_stage_model.py restates every stage in numpy with the same indexing.
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "fft_regs.cuh"

namespace impdar {
namespace sfft {

using namespace fftr;

constexpr int N2 = 256;      // contiguous sub-transform length of a row
constexpr int NC2 = 16;      // n2 values per P1/P5 tile (128-byte runs)
constexpr int TILE = 8192;   // complex elements per P1/P5 tile (64 KB)

// ------------------------------------------------------------------------------------------- tables
__global__ void twiddle_table_kernel(cf *__restrict__ tab, int n, int count) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)n, &s, &c);
    tab[m] = mk((float)c, (float)(-s));  // forward twiddle w_n^m = e^{-2 pi i m / n}
}

// ------------------------------------------------------------------ two-step FFT along the slow axis of a tile
// Element (slot, col) lives at sm[slot * PITCH + (SK ? slot >> 4 : 0) + col]; N = RA * RB slots.  Lanes run over
// `col`, so every access is conflict free.  In place; afterwards slot p holds X[k] with k = p / RB + RA * (p % RB).
template <int PITCH, int SK>
FFTR_DI int slot_addr(int slot) { return slot * PITCH + (SK ? (slot >> 4) : 0); }
template <int RA, int RB>
FFTR_DI int slot_to_k(int p) { return (p / RB) + RA * (p % RB); }
template <int RA, int RB>
FFTR_DI int k_to_slot(int k) { return (k % RA) * RB + (k / RA); }

template <int RA, int RB, int DIR, int PITCH, int SK, int NTHREADS>
FFTR_DI void tile_fft(cf *__restrict__ sm, const cf *__restrict__ tw, int tw_stride, int ncols, int tid) {
    // step 1: radix RA over q for every (r, col); twiddle w_N^{r ka}
    for (int item = tid; item < RB * ncols; item += NTHREADS) {
        const int r = item / ncols, col = item - r * ncols;
        cf v[RA];
#pragma unroll
        for (int q = 0; q < RA; ++q) v[q] = sm[slot_addr<PITCH, SK>(q * RB + r) + col];
        fft_reg<RA, DIR>(v);
        if (RB > 1) {
#pragma unroll
            for (int ka = 1; ka < RA; ++ka) v[ka] = cmul_dir<DIR>(v[ka], tw[(r * ka) * tw_stride]);
        }
#pragma unroll
        for (int ka = 0; ka < RA; ++ka) sm[slot_addr<PITCH, SK>(ka * RB + r) + col] = v[ka];
    }
    __syncthreads();
    if (RB > 1) {
        for (int item = tid; item < RA * ncols; item += NTHREADS) {
            const int ka = item / ncols, col = item - ka * ncols;
            cf v[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) v[r] = sm[slot_addr<PITCH, SK>(ka * RB + r) + col];
            fft_reg<RB, DIR>(v);
#pragma unroll
            for (int kb = 0; kb < RB; ++kb) sm[slot_addr<PITCH, SK>(ka * RB + kb) + col] = v[kb];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- taper
// min(i, n-1-i) / len clipped to 1 (mig_python.py:152-156).  `iceil` = ceil(len) as an int (INT_MAX when len is
// not finite), so the interior - where the weight is exactly 1 - is decided without any fp64 work.
FFTR_DI double taper_w(int i, int n, double len, int iceil) {
    const int m = min(i, n - 1 - i);
    if (m > 0 && m >= iceil) return 1.0;
    double w = (double)m / len;  // 0/0 -> NaN, like numpy
    if (w > 1.0) w = 1.0;
    return w;
}

// ------------------------------------------------------------------------------ P1 / P5: strided row sub-transform
struct RowAParams {
    const float *data;  // P1 input (S, T) real
    cf *W1;             // P1 output / P5 input (S, Th) complex; P5 writes the real image in place
    int S, T, Th, N1, rows_per_cta;
    double htaper, vtaper;
    int hceil, vceil;   // ceil(htaper), ceil(vtaper)
    int trunc_int;
    const cf *twTh;     // w_Th^m, m < Th
};

template <int N1, int RA, int RB, int DIR>
__global__ void __launch_bounds__(256) stolt_rowA_kernel(const __grid_constant__ RowAParams p) {
    static_assert(RA * RB == N1, "radix split");
    constexpr int NS = TILE / (N1 * NC2);  // rows per tile
    constexpr int COLS = NS * NC2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *tile = reinterpret_cast<cf *>(smem_raw);
    cf *tab = tile + TILE;         // [N1][NC2]: w_Th^{k1 n2}
    cf *tw = tab + N1 * NC2;       // [N1]: w_N1^m
    const int tid = threadIdx.x;
    const int n2b = blockIdx.x * NC2;
    for (int i = tid; i < N1 * NC2; i += 256) {
        const int k1 = i / NC2, n2 = n2b + (i % NC2);
        tab[i] = p.twTh[(int)(((long long)k1 * n2) % p.Th)];
    }
    for (int i = tid; i < N1; i += 256) tw[i] = p.twTh[i * N2];
    __syncthreads();
    const int row_begin = blockIdx.y * p.rows_per_cta;
    const int row_end = min(p.S, row_begin + p.rows_per_cta);
    for (int s0 = row_begin; s0 < row_end; s0 += NS) {
        // ---- load
#pragma unroll 4
        for (int i = 0; i < TILE / 256; ++i) {
            const int e = i * 256 + tid;
            const int n2l = e % NC2, sl = (e / NC2) % NS, n1 = e / COLS;
            const int s = s0 + sl;
            const int j = n1 * N2 + n2b + n2l;
            cf val = mk(0.f, 0.f);
            if (s < p.S) {
                if (DIR < 0) {
                    const float2 d = *reinterpret_cast<const float2 *>(p.data + (size_t)s * p.T + 2 * j);
                    const double v = taper_w(s, p.S, p.vtaper, p.vceil);
                    const double h0 = taper_w(2 * j, p.T, p.htaper, p.hceil);
                    const double h1 = taper_w(2 * j + 1, p.T, p.htaper, p.hceil);
                    if (v == 1.0 && h0 == 1.0 && h1 == 1.0 && !p.trunc_int) {
                        val = d;
                    } else {
                        double a = (double)d.x * h0 * v, b = (double)d.y * h1 * v;  // (data * H) * V, mig_python.py:157
                        if (p.trunc_int) {
                            a = trunc(a);
                            b = trunc(b);
                        }
                        val = mk((float)a, (float)b);
                    }
                } else {
                    val = cmul_dir<1>(p.W1[(size_t)s * p.Th + j], tab[n1 * NC2 + n2l]);
                }
            }
            tile[e] = val;
        }
        __syncthreads();
        tile_fft<RA, RB, DIR, COLS, 0, 256>(tile, tw, 1, COLS, tid);
        // ---- store
#pragma unroll 4
        for (int i = 0; i < TILE / 256; ++i) {
            const int e = i * 256 + tid;
            const int n2l = e % NC2, sl = (e / NC2) % NS, pslot = e / COLS;
            const int s = s0 + sl;
            const int k = slot_to_k<RA, RB>(pslot);
            if (s < p.S) {
                const cf val = tile[e];
                const size_t o = (size_t)s * p.Th + k * N2 + n2b + n2l;
                if (DIR < 0) p.W1[o] = cmul(val, tab[k * NC2 + n2l]);
                else p.W1[o] = val;  // (re, im) = (out[s][2j], out[s][2j+1])
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------- P2 / P4: contiguous row sub-transform + (un)tangle + transpose
struct RowBParams {
    cf *W1;   // (S, Th)
    cf *Dt;   // (Th, S)
    int S, T, Th, N1;
    const cf *tw512;  // w_512^m
};

constexpr int RB_PITCH = 33;
constexpr int RB_TILE = N2 * RB_PITCH + 16;

template <int DIR>
__global__ void __launch_bounds__(256) stolt_rowB_kernel(const __grid_constant__ RowBParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *tile = reinterpret_cast<cf *>(smem_raw);
    cf *tw = tile + RB_TILE;  // [512]
    const int tid = threadIdx.x;
    const int k1a = blockIdx.x, k1b = (p.N1 - k1a) % p.N1;
    const int nk = (k1a == k1b) ? 1 : 2;
    const int ncols = nk * 16;
    const int s0 = blockIdx.y * 16;
    for (int i = tid; i < 512; i += 256) tw[i] = p.tw512[i];
    double sn, cs;
    sincospi(2.0 * (double)k1a / (double)p.T, &sn, &cs);
    const cf wA = mk((float)cs, (float)(-sn));  // w_T^{k1a}
    const size_t S = (size_t)p.S;

    if (DIR < 0) {
        for (int e = tid; e < nk * 4096; e += 256) {
            const int n2 = e & 255, sl = (e >> 8) & 15, kidx = e >> 12;
            const int k1 = kidx ? k1b : k1a;
            tile[slot_addr<RB_PITCH, 1>(n2) + kidx * 16 + sl] = p.W1[(size_t)(s0 + sl) * p.Th + k1 * N2 + n2];
        }
    } else {
        for (int e = tid; e < nk * 4096; e += 256) {
            const int sl = e & 15, k2 = (e >> 4) & 255, kidx = e >> 12;
            const int k1 = kidx ? k1b : k1a;
            tile[slot_addr<RB_PITCH, 1>(k2) + kidx * 16 + sl] = p.Dt[((size_t)k1 * N2 + k2) * S + s0 + sl];
        }
    }
    __syncthreads();
    if (DIR < 0) tile_fft<16, 16, DIR, RB_PITCH, 1, 256>(tile, tw, 2, ncols, tid);

    // pairs (k1a, k2) <-> (k1b, k2p): the columns holding kx and T/2 - kx
    const int npair_k2 = (nk == 2) ? 256 : (k1a == 0 ? 129 : 128);
    for (int q = tid; q < npair_k2 * 16; q += 256) {
        const int sl = q & 15, k2 = q >> 4;
        const int k2p = (k1a == 0) ? ((256 - k2) & 255) : (255 - k2);
        const int colA = sl, colB = (nk == 2 ? 16 : 0) + sl;
        // forward: the transform output sits in digit-transposed slots; inverse: natural slots (input of the IFFT)
        const int sA = DIR < 0 ? k_to_slot<16, 16>(k2) : k2;
        const int sB = DIR < 0 ? k_to_slot<16, 16>(k2p) : k2p;
        const int aA = slot_addr<RB_PITCH, 1>(sA) + colA, aB = slot_addr<RB_PITCH, 1>(sB) + colB;
        const cf zA = tile[aA], zB = tile[aB];
        const bool self = (nk == 1) && (k2p == k2);
        const cf wk = cmul(wA, tw[k2]);  // w_T^{kx}, kx = k1a + N1 k2
        if (DIR < 0) {
            const size_t oA = ((size_t)k1a * N2 + k2) * S + s0 + sl;
            if (k1a == 0 && k2 == 0) {
                p.Dt[oA] = mk(zA.x + zA.y, zA.x - zA.y);  // (D[0], D[T/2]), both real
            } else {
                const cf e = mk(0.5f * (zA.x + zB.x), 0.5f * (zA.y - zB.y));    // (zA + conj zB) / 2
                const cf o = mk(0.5f * (zA.y + zB.y), -0.5f * (zA.x - zB.x));   // (zA - conj zB) / (2i)
                p.Dt[oA] = cadd(e, cmul(wk, o));
                if (!self) {
                    const cf wkb = mk(-wk.x, wk.y);  // w_T^{T/2 - kx} = -conj(w_T^{kx})
                    const size_t oB = ((size_t)k1b * N2 + k2p) * S + s0 + sl;
                    p.Dt[oB] = cadd(cconj(e), cmul(wkb, cconj(o)));
                }
            }
        } else {
            if (k1a == 0 && k2 == 0) {
                tile[aA] = mk(zA.x + zA.y, zA.x - zA.y);
            } else {
                // Zq[k] = (G[k] + conj G[k']) + i (G[k] - conj G[k']) conj(w_T^k)
                const cf eA = mk(zA.x + zB.x, zA.y - zB.y);
                const cf dA = cmul_dir<1>(mk(zA.x - zB.x, zA.y + zB.y), wk);
                tile[aA] = mk(eA.x - dA.y, eA.y + dA.x);
                if (!self) {
                    const cf wkb = mk(-wk.x, wk.y);
                    const cf eB = mk(zB.x + zA.x, zB.y - zA.y);
                    const cf dB = cmul_dir<1>(mk(zB.x - zA.x, zB.y + zA.y), wkb);
                    tile[aB] = mk(eB.x - dB.y, eB.y + dB.x);
                }
            }
        }
    }
    if (DIR > 0) {
        __syncthreads();
        tile_fft<16, 16, DIR, RB_PITCH, 1, 256>(tile, tw, 2, ncols, tid);
        for (int e = tid; e < nk * 4096; e += 256) {
            const int n2 = e & 255, sl = (e >> 8) & 15, kidx = e >> 12;
            const int k1 = kidx ? k1b : k1a;
            p.W1[(size_t)(s0 + sl) * p.Th + k1 * N2 + n2] =
                tile[slot_addr<RB_PITCH, 1>(k_to_slot<16, 16>(n2)) + kidx * 16 + sl];
        }
    }
}

// -------------------------------------------------------------------- P3: column transform + remap + inverse
struct ColParams {
    cf *Dt;  // (Th, S), in place
    int Th, N1, T;
    const cf *twS;  // w_S^k, k < S / R3
    double beta_unit;
    float norm;
};

template <int SH>
FFTR_DI int padi(int a) { return a + (a >> SH); }

// One Stockham step of radix R on the shared-memory sequence: item j reads x[j + t S/R], multiplies by
// w_{Ls R}^{k t} (k = j mod Ls), transforms, and writes y[(j - k) R + k + t Ls].
template <int S, int R, int DIR, int NT, int SH, bool TO_GLOBAL>
FFTR_DI void stockham_step(cf *__restrict__ buf, const cf *__restrict__ tw, int Ls, int tw_stride, int tid,
                           cf *__restrict__ gdst) {
    constexpr int ITEMS = S / R / NT;
    static_assert(ITEMS >= 1, "too many threads");
    cf v[ITEMS][R];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int j = tid + it * NT;
#pragma unroll
        for (int t = 0; t < R; ++t) v[it][t] = buf[padi<SH>(j + t * (S / R))];
    }
    if (!TO_GLOBAL) __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int j = tid + it * NT;
        const int k = j & (Ls - 1);
        apply_powers<R, DIR>(v[it], tw[k * tw_stride]);
        fft_reg<R, DIR>(v[it]);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) {
            if (TO_GLOBAL) gdst[base + t * Ls] = v[it][t];
            else buf[padi<SH>(base + t * Ls)] = v[it][t];
        }
    }
    if (!TO_GLOBAL) __syncthreads();
}

// remap coordinate of output row jj (1 <= jj < S/2) for beta^2 = b2: source row i0, weight a, scale sc
FFTR_DI void remap_coord(int jj, double b2, int nz, float norm, int &i0, float &a, float &sc) {
    const double q = fma((double)jj, (double)jj, b2);
    const float r0 = sqrtf((float)q);
    const double r = (double)r0;
    const double f = fma(fma(-r, r, q), (double)(0.5f / r0), r);  // one Newton step: |rel err| ~ 1e-14
    const double fq = fmin(f, (double)nz);
    i0 = min((int)fq, nz - 1);
    a = (float)(fq - (double)i0);
    sc = ((float)jj / (float)f) * norm;
}

template <int S, int R1, int R2, int R3, int NT>
__global__ void __launch_bounds__(NT, (NT >= 256 ? 2 : 4)) stolt_col_kernel(const __grid_constant__ ColParams p) {
    constexpr int SH = (R1 == 32) ? 5 : 4;
    constexpr int I1 = S / R1 / NT;
    static_assert(I1 >= 1 && R1 * R2 * R3 == S, "factorisation");
    constexpr int NZ = S / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *buf = reinterpret_cast<cf *>(smem_raw);
    cf *tw = buf + (S + (S >> SH));
    const int tid = threadIdx.x;
    for (int i = tid; i < S / R3; i += NT) tw[i] = p.twS[i];

    for (int c = blockIdx.x; c < p.Th; c += gridDim.x) {
        cf *col = p.Dt + (size_t)c * S;
        // ---- forward step 1 (radix R1, no twiddle) straight from global memory
        {
            cf v[I1][R1];
#pragma unroll
            for (int it = 0; it < I1; ++it) {
                const int j = tid + it * NT;
#pragma unroll
                for (int t = 0; t < R1; ++t) v[it][t] = __ldcs(col + j + t * (S / R1));
            }
#pragma unroll
            for (int it = 0; it < I1; ++it) fft_reg<R1, -1>(v[it]);
            __syncthreads();  // the previous column's last shared-memory reads are done
#pragma unroll
            for (int it = 0; it < I1; ++it) {
                const int j = tid + it * NT;
#pragma unroll
                for (int t = 0; t < R1; ++t) buf[padi<SH>(j * R1 + t)] = v[it][t];
            }
            __syncthreads();
        }
        stockham_step<S, R2, -1, NT, SH, false>(buf, tw, R1, S / (R1 * R2), tid, nullptr);
        stockham_step<S, R3, -1, NT, SH, false>(buf, tw, R1 * R2, 1, tid, nullptr);
        // buf now holds F[w], natural order

        // ---- remap fused with inverse step 1 (radix R1, no twiddle): item j needs Q[j + t S/R1]
        {
            const int kx = (c / N2) + p.N1 * (c % N2);
            const double beta = p.beta_unit * (double)(c == 0 ? p.T / 2 : kx);
            const double b2 = beta * beta;
            cf v[I1][R1];
#pragma unroll
            for (int it = 0; it < I1; ++it) {
                const int j = tid + it * NT;
#pragma unroll
                for (int t = 0; t < R1; ++t) {
                    const int w = j + t * (S / R1);
                    cf q = mk(0.f, 0.f);
                    if (w != 0 && w != NZ) {
                        const bool neg = w > NZ;
                        const int jj = neg ? S - w : w;
                        int i0;
                        float a, sc;
                        remap_coord(jj, b2, NZ, p.norm, i0, a, sc);
                        const float w0 = (1.f - a) * sc, w1 = a * sc;
                        if (c != 0) {
                            const int ia = neg ? ((S - i0) & (S - 1)) : i0;
                            const int ib = neg ? (S - i0 - 1) : (i0 + 1);
                            const cf f0 = buf[padi<SH>(ia)], f1 = buf[padi<SH>(ib)];
                            q = mk(fmaf(f0.x, w0, f1.x * w1), fmaf(f0.y, w0, f1.y * w1));
                        } else {
                            // packed column: Z = F0 + i FN with F0 (kx = 0) and FN (kx = T/2) Hermitian in w
                            const cf zj = buf[padi<SH>(jj)], zjm = buf[padi<SH>(S - jj)];
                            const cf f0 = mk(0.5f * (zj.x + zjm.x), 0.5f * (zj.y - zjm.y));  // F0[jj]
                            const cf z0 = buf[padi<SH>(i0)], z0m = buf[padi<SH>((S - i0) & (S - 1))];
                            const cf z1 = buf[padi<SH>(i0 + 1)], z1m = buf[padi<SH>(S - i0 - 1)];
                            const cf n0 = mk(0.5f * (z0.y + z0m.y), -0.5f * (z0.x - z0m.x));  // FN[i0]
                            const cf n1 = mk(0.5f * (z1.y + z1m.y), -0.5f * (z1.x - z1m.x));  // FN[i0 + 1]
                            cf q0 = cscale(f0, p.norm);  // kx = 0: beta = 0, w' = w, scale 1
                            cf qn = mk(fmaf(n0.x, w0, n1.x * w1), fmaf(n0.y, w0, n1.y * w1));
                            if (neg) {
                                q0 = cconj(q0);
                                qn = cconj(qn);
                            }
                            q = mk(q0.x - qn.y, q0.y + qn.x);  // Q0 + i QN
                        }
                    }
                    v[it][t] = q;
                }
            }
#pragma unroll
            for (int it = 0; it < I1; ++it) fft_reg<R1, 1>(v[it]);
            __syncthreads();
#pragma unroll
            for (int it = 0; it < I1; ++it) {
                const int j = tid + it * NT;
#pragma unroll
                for (int t = 0; t < R1; ++t) buf[padi<SH>(j * R1 + t)] = v[it][t];
            }
            __syncthreads();
        }
        stockham_step<S, R2, 1, NT, SH, false>(buf, tw, R1, S / (R1 * R2), tid, nullptr);
        stockham_step<S, R3, 1, NT, SH, true>(buf, tw, R1 * R2, 1, tid, col);
    }
}

// ---------------------------------------------------------------------------------------- host side
struct Tables {
    cf *twTh = nullptr, *tw512 = nullptr, *twS = nullptr;
};
static std::map<std::tuple<int, int, int>, Tables> g_tables;  // (device, S, T)
static std::mutex g_tables_mu;

static int col_r3(int S) { return S >= 4096 ? 16 : S / 256; }

static int get_tables(int S, int T, cudaStream_t st, Tables &out) {
    int dev = 0;
    IMPDAR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tables_mu);
    auto key = std::make_tuple(dev, S, T);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        out = it->second;
        return IMPDAR_B200_OK;
    }
    const int Th = T / 2;
    const int nS = S / col_r3(S);
    cf *mem = nullptr;
    IMPDAR_CUDA(cudaMalloc((void **)&mem, (size_t)(Th + 512 + nS) * sizeof(cf)));
    Tables t;
    t.twTh = mem;
    t.tw512 = mem + Th;
    t.twS = mem + Th + 512;
    twiddle_table_kernel<<<(Th + 255) / 256, 256, 0, st>>>(t.twTh, Th, Th);
    IMPDAR_LAUNCH_CHECK();
    twiddle_table_kernel<<<2, 256, 0, st>>>(t.tw512, 512, 512);
    IMPDAR_LAUNCH_CHECK();
    twiddle_table_kernel<<<(nS + 255) / 256, 256, 0, st>>>(t.twS, S, nS);
    IMPDAR_LAUNCH_CHECK();
    g_tables[key] = t;
    out = t;
    return IMPDAR_B200_OK;
}

template <int N1, int RA, int RB, int DIR>
static int launch_rowA(const RowAParams &p, cudaStream_t st) {
    const size_t smem = (size_t)(TILE + N1 * NC2 + N1) * sizeof(cf);
    static bool attr_done = false;
    if (!attr_done) {
        IMPDAR_CUDA(cudaFuncSetAttribute(stolt_rowA_kernel<N1, RA, RB, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        attr_done = true;
    }
    dim3 grid(N2 / NC2, (p.S + p.rows_per_cta - 1) / p.rows_per_cta);
    stolt_rowA_kernel<N1, RA, RB, DIR><<<grid, 256, smem, st>>>(p);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <int DIR>
static int dispatch_rowA(const RowAParams &p, cudaStream_t st) {
    switch (p.N1) {
        case 16: return launch_rowA<16, 16, 1, DIR>(p, st);
        case 32: return launch_rowA<32, 8, 4, DIR>(p, st);
        case 64: return launch_rowA<64, 16, 4, DIR>(p, st);
        case 128: return launch_rowA<128, 16, 8, DIR>(p, st);
        case 256: return launch_rowA<256, 16, 16, DIR>(p, st);
    }
    set_error("stolt: unsupported row split N1 = %d", p.N1);
    return IMPDAR_B200_EINVAL;
}

template <int DIR>
static int launch_rowB(const RowBParams &p, cudaStream_t st) {
    const size_t smem = (size_t)(RB_TILE + 512) * sizeof(cf);
    static bool attr_done = false;
    if (!attr_done) {
        IMPDAR_CUDA(cudaFuncSetAttribute(stolt_rowB_kernel<DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(p.N1 / 2 + 1, p.S / 16);
    stolt_rowB_kernel<DIR><<<grid, 256, smem, st>>>(p);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <int S, int R1, int R2, int R3, int NT>
static int launch_col(const ColParams &p, cudaStream_t st) {
    constexpr int SH = (R1 == 32) ? 5 : 4;
    const size_t smem = (size_t)(S + (S >> SH) + S / R3) * sizeof(cf);
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        IMPDAR_CUDA(cudaFuncSetAttribute(stolt_col_kernel<S, R1, R2, R3, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        int n = 0;
        IMPDAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stolt_col_kernel<S, R1, R2, R3, NT>, NT, smem));
        ctas_per_sm = n > 0 ? n : 1;
    }
    int grid = num_sms() * ctas_per_sm;
    if (grid > p.Th) grid = p.Th;
    stolt_col_kernel<S, R1, R2, R3, NT><<<grid, NT, smem, st>>>(p);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

static int dispatch_col(int S, const ColParams &p, cudaStream_t st) {
    switch (S) {
        case 512: return launch_col<512, 16, 16, 2, 32>(p, st);
        case 1024: return launch_col<1024, 16, 16, 4, 64>(p, st);
        case 2048: return launch_col<2048, 16, 16, 8, 128>(p, st);
        case 4096: return launch_col<4096, 16, 16, 16, 256>(p, st);
        case 8192: return launch_col<8192, 32, 16, 16, 256>(p, st);
    }
    set_error("stolt: unsupported column length %d", S);
    return IMPDAR_B200_EINVAL;
}

}  // namespace sfft

bool stolt_fft_supported(int S, int T) {
    const bool s_ok = (S == 512 || S == 1024 || S == 2048 || S == 4096 || S == 8192);
    const bool t_ok = (T == 8192 || T == 16384 || T == 32768 || T == 65536 || T == 131072);
    return s_ok && t_ok;
}

size_t stolt_fft_workspace_bytes(int S, int T) { return (size_t)S * (size_t)T * sizeof(float) + 256; }

// Runs passes P1..P5 (stop_after in 1..5 leaves the intermediate buffers in place for the stage tests).
int stolt_fft_run(const float *data, float *out, int S, int T, double dt, double dx, double vel, double htaper,
                  double vtaper, int trunc_int, void *workspace, int stop_after, cudaStream_t st) {
    using namespace sfft;
    Tables tb;
    int rc = get_tables(S, T, st, tb);
    if (rc) return rc;
    const int Th = T / 2, N1 = Th / N2;
    cf *W1 = reinterpret_cast<cf *>(out);
    cf *W2 = reinterpret_cast<cf *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);

    RowAParams pa;
    pa.data = data; pa.W1 = W1; pa.S = S; pa.T = T; pa.Th = Th; pa.N1 = N1;
    pa.rows_per_cta = 64;
    pa.htaper = htaper; pa.vtaper = vtaper; pa.trunc_int = trunc_int; pa.twTh = tb.twTh;
    pa.hceil = (htaper == htaper && htaper < 2.0e9) ? (int)ceil(htaper) : 0x7fffffff;
    pa.vceil = (vtaper == vtaper && vtaper < 2.0e9) ? (int)ceil(vtaper) : 0x7fffffff;
    RowBParams pb;
    pb.W1 = W1; pb.Dt = W2; pb.S = S; pb.T = T; pb.Th = Th; pb.N1 = N1; pb.tw512 = tb.tw512;
    ColParams pc;
    pc.Dt = W2; pc.Th = Th; pc.N1 = N1; pc.T = T; pc.twS = tb.twS;
    pc.beta_unit = vel * (double)S * dt / (2.0 * (double)T * dx);
    pc.norm = (float)(1.0 / ((double)S * (double)T));

    if ((rc = dispatch_rowA<-1>(pa, st))) return rc;
    if (stop_after == 1) return IMPDAR_B200_OK;
    if ((rc = launch_rowB<-1>(pb, st))) return rc;
    if (stop_after == 2) return IMPDAR_B200_OK;
    if ((rc = dispatch_col(S, pc, st))) return rc;
    if (stop_after == 3) return IMPDAR_B200_OK;
    if ((rc = launch_rowB<1>(pb, st))) return rc;
    if (stop_after == 4) return IMPDAR_B200_OK;
    return dispatch_rowA<1>(pa, st);
}

}  // namespace impdar
