// Uniform-geometry Kirchhoff, shared-memory staged form (north-star (a): output traces tiled per CTA, input trace
// windows staged through shared memory by TMA).  Included by kirchhoff.cu after KirchTabParams.
//
// The table form of the sum (kirchhoff.cu) is out[ti, x] = sum_m w[ti, m] (g[k(ti, m), x - m] + g[k(ti, m), x + m]):
// one 4-byte gather per pair.  Gathered from global memory every pair costs an L1 request (misaligned: ~1.3 data-pipe
// wavefronts) and 40 % of them miss to L2 (ncu, profiles/r01g_ncu_full_kirch_table.txt).  Here a CTA owns KT_Q
// consecutive output rows x KT_X output traces.  Rows of one block walk nearly the same hyperbola, so for a source
// row k the offsets m used by ANY row of the block form a short interval [mlo, mhi] (a device-built schedule, exact,
// from the table itself): the two row segments g[k, x0 - mhi .. x0 + X - mlo) and g[k, x0 + mlo .. x0 + X + mhi) are
// bulk-copied (cp.async.bulk, mbarrier complete_tx) ONCE into a ring of shared-memory stages by a producer warp and
// serve every (row, m) of the block that picks row k: each pair is one conflict-free shared-memory read (any
// alignment is one wavefront), L2 traffic drops from ~2 B to ~0.7 B per pair.
//
// Summation order per output sample: m ascending, partial sums flushed into a second accumulator at the end of every
// stage; stages are windows of KT_W source rows at ABSOLUTE multiples of KT_W, so the order - and therefore the
// result, bit for bit - does not depend on how the image is cut into row chunks, trace ranges or CTAs.
#pragma once

namespace impdar {

constexpr int KT_NW = 16;            // consumer warps per CTA
constexpr int KT_RW = 2;             // output rows per consumer warp
constexpr int KT_Q = KT_NW * KT_RW;  // output rows per CTA
constexpr int KT_TR = 8;             // output traces per lane (stride 32)
constexpr int KT_X = 32 * KT_TR;     // output traces per CTA
constexpr int KT_W = 16;             // source rows per stage (<= 32: one producer lane per row)
constexpr int KT_NST = 4;            // stages in the ring
constexpr int KT_SPAN = 120;         // widest [mlo, mhi] interval the staged segments hold
constexpr int KT_SEGW = KT_X + KT_SPAN + 8;                   // floats per staged segment (alignment shift <= 3)
constexpr int KT_STAGE_FLOATS = KT_W * 2 * KT_SEGW;
constexpr size_t KT_SMEM = (size_t)KT_NST * KT_STAGE_FLOATS * sizeof(float) + KT_NST * KT_W * sizeof(int2) +
                           2 * KT_NST * sizeof(unsigned long long) + 16;
constexpr int KT_THREADS = (KT_NW + 1) * 32;

struct KirchTileParams {
    const int2 *seg;     // [nblocks][S]: {mlo, mhi} per source row (mlo > mhi: not used by the block)
    const int *nstage;   // [nblocks]: stages the block walks (from its first row's absolute window)
    int *sched_flags;    // [0] != 0: some interval is wider than KT_SPAN -> this kernel stands down (table kernel runs)
};

__device__ __forceinline__ unsigned kt_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kt_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kt_mbar_arrive_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kt_mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned sb = kt_smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KT_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KT_WAIT_DONE;\n"
        "bra KT_WAIT_LOOP;\n"
        "KT_WAIT_DONE:\n"
        "}\n" ::"r"(sb),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void kt_bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     kt_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(kt_smem_u32(bar))
                 : "memory");
}

// Schedule: for every block of KT_Q output rows of this launch and every source row k, the interval of offsets m with
// k(ti, m) == k for some row ti of the block.  One CTA per block; the table rows are walked once (S x A1 entries in all).
__global__ void __launch_bounds__(256) kirch_tile_sched_kernel(const int2 *__restrict__ tab, const int *__restrict__ nm,
                                                               int S, int A1, int s_begin, int s_end, int2 *__restrict__ seg,
                                                               int *__restrict__ nstage, int *__restrict__ sched_flags) {
    __shared__ int s_kmax, s_wide;
    const int b = blockIdx.x;
    const int t0 = s_begin + b * KT_Q, t1 = min(t0 + KT_Q, s_end);
    int2 *__restrict__ sb = seg + (size_t)b * S;
    if (threadIdx.x == 0) { s_kmax = -1; s_wide = 0; }
    for (int k = t0 + threadIdx.x; k < S; k += blockDim.x) sb[k] = make_int2(0x7fffffff, -1);
    __syncthreads();
    int kmax = -1;
    for (int ti = t0; ti < t1; ++ti) {
        const int n = nm[ti];
        const int2 *__restrict__ row = tab + (size_t)ti * A1;
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            const int k = row[m].x;
            if (k >= 0) {
                atomicMin(&sb[k].x, m);
                atomicMax(&sb[k].y, m);
                kmax = max(kmax, k);
            }
        }
    }
    if (kmax >= 0) atomicMax(&s_kmax, kmax);
    __syncthreads();
    int wide = 0;
    for (int k = t0 + threadIdx.x; k < S; k += blockDim.x) {
        const int2 e = sb[k];
        if (e.y >= e.x && e.y - e.x > KT_SPAN) wide = 1;
    }
    if (wide) atomicOr(&s_wide, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int kbase = (t0 / KT_W) * KT_W;
        nstage[b] = s_kmax < 0 ? 0 : (s_kmax - kbase) / KT_W + 1;
        if (s_wide) atomicOr(sched_flags, 1);
    }
}

// (output sample, input trace) pairs of the table rows [s_begin, s_end) for output traces [x_begin, x_end): what the
// STATS variant of the table kernel counts pair by pair, in closed form per table entry.
__global__ void __launch_bounds__(256) kirch_table_count_kernel(const int2 *__restrict__ tab, const int *__restrict__ nm,
                                                                int A1, int T, int x_begin, int x_end, int s_begin,
                                                                unsigned long long *__restrict__ stats) {
    const int ti = s_begin + blockIdx.x;
    const int n = nm[ti];
    const int2 *__restrict__ row = tab + (size_t)ti * A1;
    unsigned long long c = 0;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        if (row[m].x < 0) continue;
        c += (unsigned long long)max(0, x_end - max(x_begin, m));                       // x - m >= 0
        if (m) c += (unsigned long long)max(0, min(x_end, T - m) - x_begin);            // x + m < T
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&stats[0], c);
}

__global__ void __launch_bounds__(KT_THREADS, 1) kirch_tile_kernel(const __grid_constant__ KirchTabParams p,
                                                                   const __grid_constant__ KirchTileParams q) {
    extern __shared__ __align__(128) unsigned char kt_smem_raw[];
    if (q.sched_flags[0] != 0 || p.flags[0] != 0) return;   // wide intervals or non-finite input: the table kernel runs
    float *buf = reinterpret_cast<float *>(kt_smem_raw);
    int2 *hdr = reinterpret_cast<int2 *>(buf + (size_t)KT_NST * KT_STAGE_FLOATS);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(hdr + KT_NST * KT_W);
    unsigned long long *empty = full + KT_NST;

    const int b = blockIdx.x;                       // row block (fast index: concurrent CTAs share the column window in L2)
    const int t0 = p.s_begin + b * KT_Q;
    const int x0 = p.x_begin + blockIdx.y * KT_X;
    const int nst = q.nstage[b];
    const int kbase = (t0 / KT_W) * KT_W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < KT_NST; ++i) {
            kt_mbar_init(&full[i], 1);
            kt_mbar_init(&empty[i], KT_NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == KT_NW) {
        // ---------------------------------------------------------------- producer: one lane per source row of the stage
        const int2 *__restrict__ sb = q.seg + (size_t)b * p.S;
        for (int s = 0; s < nst; ++s) {
            const int slot = s % KT_NST;
            if (s >= KT_NST) kt_mbar_wait(&empty[slot], (unsigned)((s / KT_NST - 1) & 1));
            float *sbuf = buf + (size_t)slot * KT_STAGE_FLOATS;
            unsigned bytes = 0, nl = 0, nr = 0;
            const float *gl = nullptr, *gr = nullptr;
            float *dl = nullptr, *dr = nullptr;
            if (lane < KT_W) {
                const int k = kbase + s * KT_W + lane;
                int2 e = make_int2(1, 0);
                if (k >= t0 && k < p.S) e = sb[k];
                if (e.y >= e.x) {
                    const int span = e.y - e.x;
                    const long long rowoff = (long long)k * p.Tp + p.Apad + x0;
                    const long long il = rowoff - e.y, ir = rowoff + e.x;          // first element of the left / right segment
                    const int shl = (int)(il & 3), shr = (int)(ir & 3);
                    nl = (unsigned)((shl + KT_X + span + 3) & ~3);
                    nr = (unsigned)((shr + KT_X + span + 3) & ~3);
                    gl = p.gP + (il - shl);
                    gr = p.gP + (ir - shr);
                    dl = sbuf + (size_t)lane * 2 * KT_SEGW;
                    dr = dl + KT_SEGW;
                    // value of output trace x0 + j at offset m: left  buf[offL + j - m], right buf[offR + j + m]
                    hdr[slot * KT_W + lane] = make_int2((int)(dl - buf) + shl + e.y, (int)(dr - buf) + shr - e.x);
                    bytes = (nl + nr) * 4u;
                }
            }
            const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
            __syncwarp();
            if (lane == 0) kt_mbar_arrive_tx(&full[slot], total);
            __syncwarp();
            if (bytes) {
                kt_bulk_load(dl, gl, nl * 4u, &full[slot]);
                kt_bulk_load(dr, gr, nr * 4u, &full[slot]);
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers: warp = KT_RW output rows x KT_X traces
    float acc[KT_RW][KT_TR], tot[KT_RW][KT_TR];
    int mcur[KT_RW], mend[KT_RW];
    const int2 *trow[KT_RW];
#pragma unroll
    for (int w = 0; w < KT_RW; ++w) {
        const int ti = t0 + warp + w * KT_NW;
        const bool live = ti < p.s_end && (x0 < p.x_end);
        mcur[w] = 0;
        mend[w] = live ? p.nm[ti] : 0;
        trow[w] = p.tab + (size_t)(live ? ti : t0) * p.A1;
#pragma unroll
        for (int r = 0; r < KT_TR; ++r) acc[w][r] = tot[w][r] = 0.f;
    }
    const float *__restrict__ lbuf = buf + lane;
    int2 ecur[KT_RW];
#pragma unroll
    for (int w = 0; w < KT_RW; ++w) ecur[w] = mend[w] > 0 ? __ldg(trow[w]) : make_int2(0x7fffffff, 0);
    for (int s = 0; s < nst; ++s) {
        const int slot = s % KT_NST;
        kt_mbar_wait(&full[slot], (unsigned)((s / KT_NST) & 1));
        const int k0 = kbase + s * KT_W, kend = k0 + KT_W;
        const int2 *__restrict__ h = hdr + slot * KT_W - k0;
#pragma unroll
        for (int w = 0; w < KT_RW; ++w) {
            int m = mcur[w];
            const int n1 = mend[w] - 1;
            const int2 *__restrict__ tp = trow[w];
            int2 e = ecur[w];
            while (e.x < kend) {                       // the sentinel ends the row
                ++tp;
                int2 en = make_int2(0x7fffffff, 0);
                if (m < n1) en = __ldg(tp);
                if (e.x >= 0) {
                    const int2 o = h[e.x];
                    const float wgt = __int_as_float(e.y);
                    const float *__restrict__ pl = lbuf + (o.x - m);
                    const float *__restrict__ pr = lbuf + (o.y + m);
                    if (m != 0) {
#pragma unroll
                        for (int r = 0; r < KT_TR; ++r) acc[w][r] = fmaf(wgt, pl[32 * r] + pr[32 * r], acc[w][r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < KT_TR; ++r) acc[w][r] = fmaf(wgt, pl[32 * r], acc[w][r]);
                    }
                }
                e = en;
                ++m;
            }
            mcur[w] = m;
            trow[w] = tp;
            ecur[w] = e;
#pragma unroll
            for (int r = 0; r < KT_TR; ++r) {
                tot[w][r] += acc[w][r];
                acc[w][r] = 0.f;
            }
        }
        __syncwarp();
        if (lane == 0) kt_mbar_arrive(&empty[slot]);
    }
#pragma unroll
    for (int w = 0; w < KT_RW; ++w) {
        const int ti = t0 + warp + w * KT_NW;
        if (ti >= p.s_end) continue;
#pragma unroll
        for (int r = 0; r < KT_TR; ++r) {
            const int x = x0 + lane + 32 * r;
            if (x < p.x_end) p.out[(size_t)ti * p.ldo + (x - p.x_begin)] = tot[w][r];
        }
    }
}

}  // namespace impdar
