// Constant-velocity phase shift as a per-kx complex matrix product on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulator in TMEM).  Reference: migrationlib/mig_python.py:396-420.
//
//   TK[tau, k] = 1/S sum_w [ (v kx/2)^2 < w^2 ] FK[w, k] z^(tau + 1),   z = exp(+i phi(w, k))
// With tau + 1 = 32 t0 + j + 1 the sum over w is, for one kx, a dense complex product:
//   TK[32 t0 + j, k] = sum_w A[t0, w] Bm[w, j],   A[t0, w] = FK[w, k] (z^32)^t0  (M = 128 rows t0),
//                                                  Bm[w, j] = z^(j + 1)         (N = 32 columns j),  K = the
// propagating frequencies of this kx.  Nothing is shared BETWEEN kx columns (the matrix depends on kx), but WITHIN a
// column it is a true dense contraction: 128 x 32 x nt complex MACs against (128 + 32) nt generated operands.
//
// One CTA per kx.  Generator warps build the operands of 32 frequencies at a time straight into shared memory in the
// UMMA canonical K-major layout (8-row x 16-byte core matrices, no swizzle): float64 phases (numpy's own operation
// sequence, written by phsh_tc_phase_kernel) give float32 seeds z^r, z^8, (z^32)^r, (z^32)^8, short float32
// recurrences fill the rows.  One elected thread issues the MMAs:
//   [C_re | C_im] (128 x 64, fp32, TMEM) += A_re [B_re | B_im]  +  A_im [-B_im | B_re]
// each as THREE TF32 products (3xTF32: a = a_hi + a_lo, a_hi b_hi + a_lo b_hi + a_hi b_lo), which keeps float32-level
// accuracy (scripts/phsh_tf32_emulation.py: 2.7e-7 relative L2 against the float64 oracle, plain TF32 2.7e-4).
// tcgen05.commit hands the stage back to the generators; the epilogue reads the accumulator with tcgen05.ld.
#include <math.h>

#include "common.cuh"

namespace impdar {

constexpr int TC_M = 128;             // rows t0 of one accumulator pass: 128 x 32 = 4096 taus
constexpr int TC_B = 32;              // taus per row
constexpr int TC_KC = 32;             // frequencies per stage
constexpr int TC_NST = 2;             // stages
#ifndef TC_CFG_GEN_WARPS
#define TC_CFG_GEN_WARPS 8
#endif
constexpr int TC_GEN_WARPS = TC_CFG_GEN_WARPS;   // 8 or 16: warp w generates core-matrix column w % 8 (frequencies 4 c .. 4 c + 3 of
                                                 // the stage); with 16 warps the rows of a column are split between two warps
constexpr int TC_HALVES = TC_GEN_WARPS / 8;
static_assert(TC_GEN_WARPS == 8 || TC_GEN_WARPS == 16, "8 or 16 generator warps");
constexpr int TC_MMA_WARP = TC_GEN_WARPS;          // one elected thread issues the MMAs
constexpr int TC_DRAIN_WARP0 = TC_GEN_WARPS + 1;   // four warps drain the accumulators (TMEM lane quarter = warp % 4)
constexpr int TC_THREADS = (TC_GEN_WARPS + 5) * 32;
// The tensor core adds into its fp32 accumulator with truncation: a chain of n accumulations shrinks the sum by
// ~n 2^-25.  With one accumulator per kx (3072 accumulations at nt = 4096) the migrated image was 5.5e-5 off the
// oracle, growing linearly with nt (r02 diag, profiles/); so every TC_DRAIN stages (24 MMAs each) the accumulator is
// moved to registers - summed there in round-to-nearest fp32 - and the next stage starts a fresh one (two TMEM
// accumulators, ping-pong, drained by four dedicated warps while the MMAs of the next stages run).
constexpr int TC_DRAIN = 1;
constexpr int TC_A_FLOATS = TC_M * TC_KC;          // one A operand array of a stage
constexpr int TC_B_FLOATS = 64 * TC_KC;            // one B operand array of a stage (N = 64 rows: re | im)
constexpr int TC_STAGE_FLOATS = 4 * TC_A_FLOATS + 4 * TC_B_FLOATS;
constexpr size_t TC_SMEM = (size_t)TC_NST * TC_STAGE_FLOATS * sizeof(float) + 128;
// canonical K-major, no swizzle: element (row, kk) of an operand with R rows lives at
//   ((kk / 4) * (R / 8) + row / 8) * 32 + (row % 8) * 4 + kk % 4     [floats]
// leading-dimension byte offset (core matrices adjacent in K) = (R / 8) * 128, stride byte offset (8-row groups) = 128
constexpr unsigned TC_LBO_A = (TC_M / 8) * 128, TC_LBO_B = (64 / 8) * 128, TC_SBO = 128;
constexpr unsigned TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct PhshTcParams {
    const float2 *FK;       // (nt, K)
    float2 *TK;             // (S, K)
    const double *turns;    // (K, nt): phi / (2 pi) per (kx, frequency bin), < 0 where evanescent
    const int *smin;        // (K): first propagating positive-frequency bin
    int nt, K, S, T;
    float inv_s;
};

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long tc_desc(const void *p, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((tc_smem_u32(p) >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(tc_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// e^{2 pi i u n}: the product in float64, reduced to one turn, then one float32 sincospi
__device__ __forceinline__ float2 tc_cis(double turns, int n) {
    double u = turns * (double)n;
    u -= rint(u);
    float s, c;
    sincospif(2.0f * (float)u, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ float2 tc_cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// 3xTF32 split: hi = the 19 bits the tensor core reads, lo = the exact remainder (its own truncation costs 2^-22)
__device__ __forceinline__ void tc_split(float v, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = v - hi;
}

// phi(w, kx) / 2 pi for every (kx, frequency bin) with numpy's float64 operation sequence (phaseshift.cu: ps_omega,
// ps_kx, ps_vkx2, ps_phi - the propagating / evanescent decision has exact ties on power-of-two geometries), and the
// first propagating positive-frequency bin of every kx.  -1 marks evanescent bins.
__global__ void __launch_bounds__(256) phsh_tc_phase_kernel(double *__restrict__ turns, int *__restrict__ smin, int nt, int K,
                                                            int T, double dt, double dx, double vel) {
    const int k = blockIdx.x;
    const double twopi = 6.283185307179586;
    const double valk = __ddiv_rn(1.0, __dmul_rn((double)T, dx));
    const double kx = __dmul_rn(twopi, __dmul_rn((double)k, valk));
    const double hk = __ddiv_rn(__dmul_rn(vel, kx), 2.0);
    const double vkx2 = __dmul_rn(hk, hk);
    const double valw = __ddiv_rn(1.0, __dmul_rn((double)nt, dt));
    int first = nt;
    for (int iw = threadIdx.x; iw < nt; iw += blockDim.x) {
        const int fi = (iw < (nt + 1) / 2) ? iw : iw - nt;
        const double w = (fi == 0) ? 1e-10 / dt : __dmul_rn(twopi, __dmul_rn((double)fi, valw));
        const double w2 = __dmul_rn(w, w);
        double u = -2.0;                       // evanescent marker (real phases are within half a turn)
        if (vkx2 < w2) {
            // the rotation per tau step is +w dt sqrt(.) (cp = conj(exp(i phase)), phase = -w dt sqrt(.), :415-416): signed,
            // negative frequencies rotate the other way
            const double phi = __dmul_rn(__dmul_rn(w, dt), __dsqrt_rn(__dsub_rn(1.0, __ddiv_rn(vkx2, w2))));
            u = phi * 0.15915494309189535;
            if (iw < nt / 2 && iw < first) first = iw;
        }
        turns[(size_t)k * nt + iw] = u;
    }
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = nt;
    __syncthreads();
    if (first < nt) atomicMin(&s_first, first);
    __syncthreads();
    if (threadIdx.x == 0) smin[k] = s_first;
}

__global__ void __launch_bounds__(TC_THREADS, 1) phsh_const_tc_kernel(const __grid_constant__ PhshTcParams p) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    float *stage0 = reinterpret_cast<float *>(tc_smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(stage0 + (size_t)TC_NST * TC_STAGE_FLOATS);
    unsigned long long *empty = full + TC_NST;
    unsigned long long *tfull = empty + TC_NST;      // [2] accumulator a holds a finished group of stages
    unsigned long long *tempty = tfull + 2;          // [2] accumulator a has been drained
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(tempty + 2);

    const int k = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = p.nt / 2;
    // propagating bins of this kx, in list order: positive frequencies [smin, nh), then negative (nh, nt - max(smin, 1)]
    const int smin = min(p.smin[k], nh);
    const int n_pos = nh - smin;
    const int n_neg = nh - max(smin, 1);
    const int n_list = n_pos + n_neg;
    const int n_stage = (n_list + TC_KC - 1) / TC_KC;
    const int n_pass = (p.S + TC_M * TC_B - 1) / (TC_M * TC_B);

    // zero the operand buffers once: rows / frequencies that are never generated must read as 0
    for (int i = threadIdx.x; i < TC_NST * TC_STAGE_FLOATS; i += TC_THREADS) stage0[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_NST; ++i) {
            tc_mbar_init(&full[i], TC_GEN_WARPS);
            tc_mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(&tfull[i], 1);
            tc_mbar_init(&tempty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;

    const int rows_needed = (p.S + TC_B - 1) / TC_B;       // t0 rows that hold real taus (over all passes)
    const int n_group = (n_stage + TC_DRAIN - 1) / TC_DRAIN;   // accumulator groups per pass
    int it = 0;                                             // running stage counter over passes (smem ring position)
    int ig = 0;                                             // running group counter over passes (accumulator ping-pong)
    for (int pass = 0; pass < n_pass; ++pass) {
        const int t0_base = pass * TC_M;
        if (warp < TC_GEN_WARPS) {
            // ------------------------------------------------------------------ generators
            const int r = lane >> 2, kq = lane & 3;        // row within a core matrix, frequency within the 16-byte chunk
            const int cc = warp & 7, half = warp >> 3;     // core-matrix column of this warp, and which part of its rows
            const int kk = 4 * cc + kq;                    // frequency within the stage
            const int groups = min(TC_M / 8, (rows_needed - t0_base + 7) / 8);   // 8-row groups with real taus
            const int ia0 = half * (TC_M / 8 / TC_HALVES), ia1 = min(groups, ia0 + TC_M / 8 / TC_HALVES);
            const int ib0 = half * (4 / TC_HALVES), ib1 = ib0 + 4 / TC_HALVES;
            // phase and spectrum value of this thread's frequency, fetched one stage ahead (the FK column is strided: L2 latency)
            auto fetch = [&](int s, double &u, float2 &fk) {
                const int pidx = s * TC_KC + kk;
                fk = make_float2(0.f, 0.f);
                u = 0.0;
                if (s < n_stage && pidx < n_list) {
                    const int iw = (pidx < n_pos) ? smin + pidx : nh + 1 + (pidx - n_pos);
                    u = p.turns[(size_t)k * p.nt + iw];
                    fk = p.FK[(size_t)iw * p.K + k];
                }
            };
            double u_next;
            float2 fk_next;
            fetch(0, u_next, fk_next);
            for (int s = 0; s < n_stage; ++s, ++it) {
                const int slot = it % TC_NST;
                double u = u_next;
                float2 fk = fk_next;
                fetch(s + 1, u_next, fk_next);
                if (!(u > -1.5)) {                         // (every listed bin propagates; the evanescent marker is a guard)
                    u = 0.0;
                    fk = make_float2(0.f, 0.f);
                }
                if (it >= TC_NST) tc_mbar_wait(&empty[slot], (unsigned)((it / TC_NST - 1) & 1));
                float *sa = stage0 + (size_t)slot * TC_STAGE_FLOATS;
                float *a_re_hi = sa, *a_re_lo = sa + TC_A_FLOATS, *a_im_hi = sa + 2 * TC_A_FLOATS, *a_im_lo = sa + 3 * TC_A_FLOATS;
                float *sb = sa + 4 * TC_A_FLOATS;
                float *b1_hi = sb, *b1_lo = sb + TC_B_FLOATS, *b2_hi = sb + 2 * TC_B_FLOATS, *b2_lo = sb + 3 * TC_B_FLOATS;
                // A: rows t0 = t0_base + 8 i + r  ->  FK (z^32)^t0 ; seed at i = 0, then multiply by (z^32)^8
                {
                    float2 a = tc_cmul(fk, tc_cis(u, TC_B * (t0_base + 8 * ia0 + r)));
                    const float2 step = tc_cis(u, TC_B * 8);
                    const int base = (cc * (TC_M / 8)) * 32 + r * 4 + kq;
                    for (int i = ia0; i < ia1; ++i) {
                        float h, l;
                        tc_split(a.x, h, l);
                        a_re_hi[base + i * 32] = h;
                        a_re_lo[base + i * 32] = l;
                        tc_split(a.y, h, l);
                        a_im_hi[base + i * 32] = h;
                        a_im_lo[base + i * 32] = l;
                        a = tc_cmul(a, step);
                    }
                }
                // B: rows n = 8 i + r (i < 4) hold z^(n + 1): [B_re | B_im] and [-B_im | B_re]
                {
                    float2 b = tc_cis(u, 8 * ib0 + r + 1);
                    const float2 step = tc_cis(u, 8);
                    const int base = (cc * 8) * 32 + r * 4 + kq;
#pragma unroll
                    for (int i = ib0; i < ib1; ++i) {
                        float h, l;
                        tc_split(b.x, h, l);
                        b1_hi[base + i * 32] = h;           // rows 0..31: re
                        b1_lo[base + i * 32] = l;
                        b2_hi[base + (i + 4) * 32] = h;     // rows 32..63 of the second operand: +re
                        b2_lo[base + (i + 4) * 32] = l;
                        tc_split(b.y, h, l);
                        b1_hi[base + (i + 4) * 32] = h;     // rows 32..63: im
                        b1_lo[base + (i + 4) * 32] = l;
                        b2_hi[base + i * 32] = -h;          // rows 0..31 of the second operand: -im
                        b2_lo[base + i * 32] = -l;
                        b = tc_cmul(b, step);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> the tensor core's reads
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&full[slot]);
            }
        } else if (warp == TC_MMA_WARP) {
            // ------------------------------------------------------------------ MMA issuer (one elected thread)
            for (int s = 0; s < n_stage; ++s, ++it) {
                const int slot = it % TC_NST;
                const int g = s / TC_DRAIN;                 // accumulator group of this pass
                const int acc = (ig + g) & 1;
                if (s % TC_DRAIN == 0 && ig + g >= 2)       // the accumulator must have been drained (group ig + g - 2)
                    tc_mbar_wait(&tempty[acc], (unsigned)((((ig + g) >> 1) - 1) & 1));
                tc_mbar_wait(&full[slot], (unsigned)((it / TC_NST) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const unsigned td = tmem_d + 64u * (unsigned)acc;
                    const float *sa = stage0 + (size_t)slot * TC_STAGE_FLOATS;
                    const float *sb = sa + 4 * TC_A_FLOATS;
#pragma unroll
                    for (int k8 = 0; k8 < TC_KC / 8; ++k8) {
                        const unsigned offa = k8 * 2 * TC_LBO_A, offb = k8 * 2 * TC_LBO_B;   // bytes: two core-matrix columns
                        const char *A0 = reinterpret_cast<const char *>(sa) + offa;
                        const char *B0 = reinterpret_cast<const char *>(sb) + offb;
                        const size_t ab = (size_t)TC_A_FLOATS * 4, bb = (size_t)TC_B_FLOATS * 4;
                        const unsigned long long d_are_hi = tc_desc(A0, TC_LBO_A, TC_SBO), d_are_lo = tc_desc(A0 + ab, TC_LBO_A, TC_SBO);
                        const unsigned long long d_aim_hi = tc_desc(A0 + 2 * ab, TC_LBO_A, TC_SBO), d_aim_lo = tc_desc(A0 + 3 * ab, TC_LBO_A, TC_SBO);
                        const unsigned long long d_b1_hi = tc_desc(B0, TC_LBO_B, TC_SBO), d_b1_lo = tc_desc(B0 + bb, TC_LBO_B, TC_SBO);
                        const unsigned long long d_b2_hi = tc_desc(B0 + 2 * bb, TC_LBO_B, TC_SBO), d_b2_lo = tc_desc(B0 + 3 * bb, TC_LBO_B, TC_SBO);
                        const unsigned acc0 = (s % TC_DRAIN != 0 || k8 > 0) ? 1u : 0u;     // a new group starts a new sum
                        // small terms first: the truncating accumulator loses less of them
                        tc_mma(td, d_are_lo, d_b1_hi, acc0);
                        tc_mma(td, d_are_hi, d_b1_lo, 1u);
                        tc_mma(td, d_aim_lo, d_b2_hi, 1u);
                        tc_mma(td, d_aim_hi, d_b2_lo, 1u);
                        tc_mma(td, d_are_hi, d_b1_hi, 1u);
                        tc_mma(td, d_aim_hi, d_b2_hi, 1u);
                    }
                    tc_commit(&empty[slot]);                 // the stage may be overwritten once these MMAs have read it
                    if (s % TC_DRAIN == TC_DRAIN - 1 || s == n_stage - 1) tc_commit(&tfull[acc]);   // group complete
                }
                __syncwarp();
            }
        } else {
            // ------------------------------------------------------------------ drain warps: TMEM -> registers -> TK
            const int q = warp & 3;                          // TMEM lane quarter of this warp
            const int t0 = t0_base + 32 * q + lane;
            float sum[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) sum[j] = 0.f;
            for (int g = 0; g < n_group; ++g) {
                const int acc = (ig + g) & 1;
                tc_mbar_wait(&tfull[acc], (unsigned)(((ig + g) >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned taddr = tmem_d + 64u * (unsigned)acc + ((unsigned)(32 * q) << 16);
#pragma unroll
                for (int c16 = 0; c16 < 4; ++c16) {
                    unsigned v[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr + 16u * (unsigned)c16));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[16 * c16 + j] += __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&tempty[acc]);          // the accumulator may be overwritten
            }

#pragma unroll
            for (int j = 0; j < TC_B; ++j) {
                const int tau = TC_B * t0 + j;
                if (tau < p.S) p.TK[(size_t)tau * p.K + k] = make_float2(sum[j] * p.inv_s, sum[32 + j] * p.inv_s);
            }
        }
        ig += n_group;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(128) : "memory");
}

// The unpaired Nyquist frequency bin contributes FK[nt/2, k] cos((tau + 1) phi) (phaseshift.cu, header): added after
// the tensor-core pass.
__global__ void __launch_bounds__(256) phsh_tc_nyq_kernel(const __grid_constant__ PhshTcParams p) {
    const int k = blockIdx.x;
    const double u = p.turns[(size_t)k * p.nt + p.nt / 2];
    if (!(u > -1.5)) return;
    const float2 f = p.FK[(size_t)(p.nt / 2) * p.K + k];
    for (int tau = threadIdx.x; tau < p.S; tau += blockDim.x) {
        double a = u * (double)(tau + 1);
        a -= rint(a);
        const float cn = (float)cospi(2.0 * a);
        float2 *dst = p.TK + (size_t)tau * p.K + k;
        float2 o = *dst;
        o.x = fmaf(f.x * p.inv_s, cn, o.x);
        o.y = fmaf(f.y * p.inv_s, cn, o.y);
        *dst = o;
    }
}

size_t phsh_tc_workspace_bytes(int nt, int K) { return (size_t)K * nt * sizeof(double) + (size_t)K * sizeof(int) + 256; }

// TK (S, K) <- FK (nt, K); `ws` holds phsh_tc_workspace_bytes(nt, K).
int phsh_const_tc_launch(const float2 *FK, float2 *TK, int nt, int K, int S, int T, double dt, double dx, double vel,
                         float inv_s, void *ws, cudaStream_t st) {
    IMPDAR_CHECK_ARG(nt >= 2 && (nt & (nt - 1)) == 0, "phsh tensor-core path: nt must be a power of two >= 2");
    double *turns = (double *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    int *smin = (int *)(turns + (size_t)K * nt);
    phsh_tc_phase_kernel<<<K, 256, 0, st>>>(turns, smin, nt, K, T, dt, dx, vel);
    IMPDAR_LAUNCH_CHECK();
    static bool attr_done[64];
    int dev = 0;
    IMPDAR_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        IMPDAR_CUDA(cudaFuncSetAttribute(phsh_const_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_done[dev] = true;
    }
    PhshTcParams p;
    p.FK = FK; p.TK = TK; p.turns = turns; p.smin = smin; p.nt = nt; p.K = K; p.S = S; p.T = T; p.inv_s = inv_s;
    ktimer_begin("phsh_const_tc_kernel", st);
    phsh_const_tc_kernel<<<K, TC_THREADS, TC_SMEM, st>>>(p);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    phsh_tc_nyq_kernel<<<K, 256, 0, st>>>(p);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // namespace impdar
