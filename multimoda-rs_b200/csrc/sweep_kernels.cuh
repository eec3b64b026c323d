// =============================================================================
// sweep_kernels.cuh — sm_100a device code of the Hausdorff rotation sweep.
//
//   K0  k_prep        f64 points -> centred FP32 staging layout (+ Rmax)
//   K1  k_sweep<TA>   FP32 sweep: one CTA = one unit x one tile of candidate
//                     angles, one warp = one candidate; reference contour
//                     staged in shared memory by a TMA bulk copy; test points
//                     rotated into registers; symmetric directed Hausdorff as
//                     a register-tiled min-over-M / max-over-N with packed
//                     f32x2 arithmetic (FADD2/FMUL2/FFMA2), 3-input FMNMX and
//                     warp REDUX; fused per-unit arg-min on a packed
//                     (distance, index) 64-bit key with atomicMin.
//   K2  k_shortlist   candidates within the FP32 error window of the minimum
//   K3  k_exact       the reference's f64 arithmetic, literally, on those
//   K4  k_select      leftmost f64 arg-min + tie count per unit
//
// Replaces (reference paths): search_range + hausdorff_distance +
// directed_hausdorff, src/intravascular/processing/process_utils.rs:33-121, and
// the cost closures align_within.rs:99-105/:200-206, align_between.rs:189-216.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cuda/std/type_traits>

#include "mmrs_internal.hpp"

namespace mmrs {

#ifndef MMRS_TRIP_UNROLL
#define MMRS_TRIP_UNROLL 1   // trips (four reference points each) per loop iteration of the sweep's main loop
#endif
#ifndef MMRS_EXP_NOTAIL   // timing experiments only (wrong results): skip the tail pass / the seed loads of exact tiling
#define MMRS_EXP_NOTAIL 0
#endif
#ifndef MMRS_EXP_NOSEED
#define MMRS_EXP_NOSEED 0
#endif
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

// ---- packed FP32 helpers (Blackwell-only PTX: *.f32x2, 3-input min.f32) --------
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

// =============================================================================
// K0: f64 (x,y) -> centred FP32 staging layout.
// A block (test points, the set that is rotated; register side of K1):
//   S = ceil(TA/2) float4 slots per lane and chunk; slot k < TA/2 at (c*S+k)*32 + l holds the
//   points i0 = c*32*TA + 2k*32 + l and i1 = i0 + 32 as (x0, x1, y0, y1); for odd TA the last
//   slot holds the single point i0 = c*32*TA + (TA-1)*32 + l as (x, x, y, y).
// B block (reference points; streamed from shared memory in K1):
//   float4 j = (bx0, by0, bx1, by1) = reference points 2j and 2j+1, j < m_pairs. The packed f32x2
//   instructions take these as scalar operands broadcast to both halves (one 32-bit register
//   read instead of a 64-bit pair: the sweep is register-file-bandwidth bound).
// Indices past the end repeat the last point (a duplicate never changes a
// min-over-points or a max-over-points).
// NB block (after the B block): float2 j = (|b_2j|^2, |b_2j+1|^2), the squared norms the expanded-form tier adds.
// Tail block (exact tiling, UnitDesc.n_tail > 0; after the NB block): the n mod 32 test points that do not fill a
// register slot, two per float4 (x0, y0, x1, y1). K1 scores them in a short pass of its own instead of a padded slot.
// =============================================================================
__global__ void k_prep(const UnitDesc* __restrict__ units, const double* __restrict__ test_xy,
                       const double* __restrict__ ref_xy, float4* __restrict__ lay, unsigned* __restrict__ rmax_bits) {
    const UnitDesc ud = units[blockIdx.x];
    if (ud.n <= 0 || ud.m <= 0) return;
    const int TA = ud.ta;
    const int half = (TA + 1) / 2, pairs = TA / 2;
    const int a_elems = ud.n_chunks * half * 32;
    // exact tiling (n_tail > 0): the register slots hold the first 32 TA points only, the rest is the tail block
    const int n_main = ud.n - ud.n_tail;
    // exact tiling: the B block is padded to 16 (TA + 1) float4 so that the tail pass can address its reference points
    // lane + 32 q without clamping (indices past the end repeat the last point)
    const int b_elems = ud.n_tail > 0 ? 16 * (TA + 1) : ud.m_pairs;
    const int nb_elems = (b_elems + 1) / 2;   // NB block: |b|^2 of the FP32 reference points, float2 per pair of points
    float4* A = lay + ud.lay_off;
    float4* B = A + a_elems;
    float2* NB = reinterpret_cast<float2*>(B + b_elems);
    float4* T = B + b_elems + nb_elems;
    float rmax = 0.f;
    for (int e = threadIdx.x; e < a_elems; e += blockDim.x) {
        int l = e & 31, ck = e >> 5;
        int c = ck / half, k = ck - c * half;
        int i0 = c * 32 * TA + 2 * k * 32 + l, i1 = (k < pairs) ? i0 + 32 : i0;
        i0 = min(i0, n_main - 1);
        i1 = min(i1, n_main - 1);
        const double* p0 = test_xy + 2 * (ud.test_off + i0);
        const double* p1 = test_xy + 2 * (ud.test_off + i1);
        float4 v;
        v.x = (float)(p0[0] - ud.cx);
        v.y = (float)(p1[0] - ud.cx);
        v.z = (float)(p0[1] - ud.cy);
        v.w = (float)(p1[1] - ud.cy);
        A[e] = v;
        rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int j = threadIdx.x; j < b_elems; j += blockDim.x) {
        const int j0 = min(2 * j, ud.m - 1), j1 = min(2 * j + 1, ud.m - 1);
        const double* p0 = ref_xy + 2 * (ud.ref_off + j0);
        const double* p1 = ref_xy + 2 * (ud.ref_off + j1);
        const float4 v = make_float4((float)(p0[0] - ud.cx), (float)(p0[1] - ud.cy), (float)(p1[0] - ud.cx),
                                     (float)(p1[1] - ud.cy));
        B[j] = v;
        // squared norms of the ROUNDED coordinates, formed in f64 and rounded once (expanded-form tier, K1x)
        NB[j] = make_float2((float)((double)v.x * v.x + (double)v.y * v.y), (float)((double)v.z * v.z + (double)v.w * v.w));
        rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (threadIdx.x == 0 && (b_elems & 1)) NB[b_elems] = make_float2(0.f, 0.f);   // padding of the odd float4
    // tail block: float4 t = tail points 2t and 2t+1 as (x0, y0, x1, y1) (an odd count repeats the last point)
    for (int t = threadIdx.x; t < (ud.n_tail + 1) / 2; t += blockDim.x) {
        const int i0 = n_main + min(2 * t, ud.n_tail - 1), i1 = n_main + min(2 * t + 1, ud.n_tail - 1);
        const double* p0 = test_xy + 2 * (ud.test_off + i0);
        const double* p1 = test_xy + 2 * (ud.test_off + i1);
        const float4 v = make_float4((float)(p0[0] - ud.cx), (float)(p0[1] - ud.cy), (float)(p1[0] - ud.cx),
                                     (float)(p1[1] - ud.cy));
        T[t] = v;
        rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    unsigned rb = __reduce_max_sync(0xffffffffu, __float_as_uint(rmax));
    if ((threadIdx.x & 31) == 0) atomicMax(&rmax_bits[blockIdx.x], rb);
}

// f64 (cos, sin) table -> FP32 table for K1.
__global__ void k_cs32(const double2* __restrict__ cs64, float2* __restrict__ cs32, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) cs32[i] = make_float2((float)cs64[i].x, (float)cs64[i].y);
}

// =============================================================================
// K1: the FP32 sweep.
// =============================================================================
// LIST = false: the dense sweep, work item = (unit, first candidate, candidates in this tile).
// LIST = true : exact re-scoring of a candidate list (after lower-bound pruning, or after the tensor-core prefilter of
//               tc_kernels.cuh). The list is ONE global array of (unit, candidate) items, contiguous per unit; CTA b owns
//               positions [b * l_chunk, (b + 1) * l_chunk) whatever units they belong to (it re-stages per run of equal
//               units), so the SMs stay busy however unevenly the survivors are spread over the units. Same arithmetic;
//               overwrites the dist32 entries and records the largest |d_old^2 - d_fp32^2| / Rmax^2 (the tensor-core
//               prefilter's error, checked against its window).
// TAILP = true : exact tiling for a test set of 32 TA + R points (0 < R < 32, single chunk). The R tail points would
//               waste most of a padded register slot (520 points: slot 17 carries 8 points and 24 duplicates, 4.4 % of
//               the launch), so per candidate the warp first runs a short pass with the roles swapped: every lane
//               holds ceil(M / 32) reference points in registers, the rotated tail points are broadcast two at a time;
//               the tail's row minima are finished with one REDUX each and its column minima — lane-local, no REDUX —
//               go to a per-warp shared-memory array from which the main loop seeds its column accumulators (one
//               broadcast LDS.64 per two reference points; the seed takes the free slot of the first 3-input minimum).
// XF = true    : K1x, the EXPANDED-FORM tier (dense launches only). |a - b|^2 = |b|^2 - 2 a.b + |a|^2: per packed pair of
//               test points and reference point two FFMA2 give r = |b|^2 - 2 a.b (what the row minima need: |a|^2 is
//               constant along a row and is added after the minimum) and one FADD2 gives r + |a|^2 (what the column
//               minima need) — 6 packed FP32 instructions per 2 x 2 block instead of 8 (-7 % per block measured,
//               profiles/r02_microbench_expanded_forms.txt). The price is cancellation: with Rn the largest point norm
//               the result carries an ABSOLUTE error of at most 13 u Rn^2 (u = 2^-24; one rounding each for |b|^2
//               (u), the two FMAs (3 u + 3 u: the partial sums stay below 3 Rn^2), the final add (4 u) and two for
//               |a|^2) <= 1.6e-6 Rmax^2, instead of the direct form's relative 3e-7. So K1x is a FILTER tier exactly like
//               K1 is a filter for the f64 recheck: every candidate is scored, the candidates within the error window
//               of the minimum are re-scored by the direct-form kernel (LIST), and K2/K3/K4 run unchanged on exact
//               FP32 values — the selection is bit-identical to the dense direct path. The tail pass and the odd
//               scalar slot stay in direct form (their values are exact, the seeds and minima mix freely).
template <int TA, bool MULTI, bool LIST, bool TAILP, bool XF>
__global__ void __launch_bounds__(kThreads, 2)
    k_sweep(const UnitDesc* __restrict__ units, const WorkItem* __restrict__ work, const float4* __restrict__ lay,
            const float2* __restrict__ cs32, float* __restrict__ dist32, unsigned long long* __restrict__ key,
            const int2* __restrict__ l_items, const unsigned* __restrict__ l_nitems, unsigned l_cap, int l_chunk,
            const unsigned* __restrict__ rmax_bits, unsigned* __restrict__ diag) {
    static_assert(TA >= 2 && TA <= 18, "register tile out of range");
    static_assert(!(TAILP && MULTI), "the tail pass is for single-chunk units");
    static_assert(!(XF && LIST), "the expanded form is a dense tier; lists are re-scored in direct form");
    constexpr int SB = TA + 1;         // TAILP: reference points per lane of the tail pass (M <= 32 SB)
    constexpr int H = TA / 2;          // packed pairs of test points per lane
    constexpr bool TAIL = (TA & 1);    // plus one unpaired point when TA is odd
    constexpr int S = H + (TAIL ? 1 : 0);
    constexpr int TU = MMRS_TRIP_UNROLL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw + 8);
    float4* sA = reinterpret_cast<float4*>(smem_raw + 16);

    // LIST: this CTA owns the global list positions [g_pos, g_hi); they may span several units (runs of equal .x).
    unsigned g_pos = 0, g_hi = 0;
    if (LIST) {
        const unsigned n_items = l_nitems ? min(*l_nitems, l_cap) : l_cap;
        g_pos = blockIdx.x * (unsigned)l_chunk;
        g_hi = min(g_pos + (unsigned)l_chunk, n_items);
        if (g_pos >= g_hi) return;  // nothing for this CTA (uniform: before any barrier)
    }
    const WorkItem w = LIST ? WorkItem{0, 0, 0, 0} : work[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for ptxas: the B-block and seed pointers live in uniform registers
    if (threadIdx.x == 0) mbar_init(bar, 1);
    uint32_t phase = 0;
    const float INF = __int_as_float(0x7f800000);
    float worst = 0.f;
    for (;;) {  // one pass for the dense sweep; one pass per run of equal units in LIST mode
    int unit = w.unit, total = w.count;
    const int2* my_items = nullptr;
    unsigned run_end = 0;
    if (LIST) {
        unit = l_items[g_pos].x;
        run_end = g_pos + 1;
        while (run_end < g_hi && l_items[run_end].x == unit) ++run_end;
        total = (int)(run_end - g_pos);
        my_items = l_items + g_pos;
        if (unit < 0) {  // neutralised entries of an overflowed reservation
            g_pos = run_end;
            if (g_pos >= g_hi) break;
            continue;
        }
    }
    const UnitDesc ud = units[unit];
    if (LIST && (ud.ta != TA || (ud.n_chunks > 1) != MULTI || (ud.n_tail > 0) != TAILP)) {
        // a run of another size class: the launch of that class scores it (uniform per CTA: before any barrier use)
        g_pos = run_end;
        if (g_pos >= g_hi) break;
        continue;
    }
    const int a_elems = ud.n_chunks * S * 32;
    const int b_elems = TAILP ? 16 * SB : ud.m_pairs;   // float4 per PAIR of reference points (exact tiling: padded)
    const int b_pts = 2 * ud.m_pairs;
    const int nb_elems = (b_elems + 1) / 2;                // float2 per PAIR of reference points: their squared norms
    const int t_elems = TAILP ? (ud.n_tail + 1) / 2 : 0;   // float4 per PAIR of tail points
    float4* sB = sA + a_elems;
    const float2* sNB = reinterpret_cast<const float2*>(sB + b_elems);
    const float4* sT = sB + b_elems + nb_elems;
    unsigned* s_col = reinterpret_cast<unsigned*>(sB + b_elems + nb_elems + t_elems);  // MULTI: [warp][b_pts rounded up to 4]; TAILP: [warp][32 SB]

    if (threadIdx.x == 0) {
        *s_key = ~0ull;
        const uint32_t bytes = (uint32_t)(a_elems + b_elems + nb_elems + t_elems) * 16u;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(sA, lay + ud.lay_off, bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, phase);
    phase ^= 1u;

    unsigned long long best = ~0ull;
    unsigned* my_col = s_col + wid * (TAILP ? 32 * SB : ((b_pts + 3) & ~3));   // 16-byte aligned rows: four seeds per LDS.128
    // LIST: squared give-up threshold from the unit's best exact distance so far (bits; 0xffffffff = none yet)
    unsigned give_up = 0xffffffffu;
    if (LIST) {
        const unsigned ub_bits = (unsigned)(key[unit] >> 32);
        if (ub_bits != 0xffffffffu) {
            const float t = __uint_as_float(ub_bits) * (1.0f + 4e-6f) + 4e-6f * __uint_as_float(rmax_bits[unit]);
            give_up = __float_as_uint(t * t * (1.0f + 1e-6f));
        }
    }

    for (int ci = wid; ci < total; ci += kWarpsPerCta) {
        const int c = LIST ? my_items[ci].y : w.begin + ci;
        const float2 cs = __ldg(&cs32[ud.cand_off + c]);
        const uint64_t C2 = pk(cs.x, cs.x), S2 = pk(cs.y, cs.y), NS2 = pk(-cs.y, -cs.y);
        unsigned rowmax = 0u, colmax = 0u;  // bit patterns of non-negative floats order like unsigned ints
        bool gave_up = false;

        if (TAILP) {
            // Tail pass, roles swapped: this lane's reference points lane, lane + 32, ... in registers (indices past the
            // end repeat the last stored point), the rotated tail points broadcast two at a time.
            const float2* sBp = reinterpret_cast<const float2*>(sB) + lane;
            float bx[SB], by[SB], tc[SB];
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                const float2 b = sBp[32 * q];   // one base register, immediate offsets (the B block is padded to 32 SB points)
                bx[q] = b.x;
                by[q] = b.y;
                tc[q] = INF;
            }
#if MMRS_EXP_NOTAIL
            for (int t = 0; t < 0; ++t) {
#else
#pragma unroll 1
            for (int t = 0; t < t_elems; ++t) {
#endif
                const float4 a = sT[t];  // (x0, y0, x1, y1): broadcast
                const uint64_t X2 = pk(a.x, a.z), Y2 = pk(a.y, a.w);
                const uint64_t PX = fma2(Y2, NS2, mul2(X2, C2)), PY = fma2(X2, S2, mul2(Y2, C2));
                float r0 = INF, r1 = INF;   // row minima of the two tail points over this lane's reference points
#pragma unroll
                for (int q = 0; q < SB; q += 2) {
                    const uint64_t dxa = sub2(PX, pk(bx[q], bx[q])), dya = sub2(PY, pk(by[q], by[q]));
                    const uint64_t da = fma2(dxa, dxa, mul2(dya, dya));   // (|t0 - b_q|^2, |t1 - b_q|^2)
                    float a0, a1;
                    upk(da, a0, a1);
                    tc[q] = min3(tc[q], a0, a1);
                    if (q + 1 < SB) {
                        const uint64_t dxb = sub2(PX, pk(bx[q + 1], bx[q + 1])), dyb = sub2(PY, pk(by[q + 1], by[q + 1]));
                        const uint64_t db = fma2(dxb, dxb, mul2(dyb, dyb));
                        float b0, b1;
                        upk(db, b0, b1);
                        tc[q + 1] = min3(tc[q + 1], b0, b1);
                        r0 = min3(r0, a0, b0);
                        r1 = min3(r1, a1, b1);
                    } else {
                        r0 = fminf(r0, a0);
                        r1 = fminf(r1, a1);
                    }
                }
                const unsigned m0 = __reduce_min_sync(0xffffffffu, __float_as_uint(r0));
                const unsigned m1 = __reduce_min_sync(0xffffffffu, __float_as_uint(r1));
                rowmax = max(rowmax, max(m0, m1));
            }
            __syncwarp();  // the previous candidate's main loop is done reading the seeds
#pragma unroll
            for (int q = 0; q < SB; ++q) my_col[lane + 32 * q] = __float_as_uint(tc[q]);
            __syncwarp();
        }

        for (int ch = 0; ch < ud.n_chunks; ++ch) {
            uint64_t AX[H], AY[H];           // XF: (-2 x', -2 y')
            uint64_t NA[XF ? H : 1];         // XF: |a'|^2 of the rotated FP32 points
            float row[TA];
            float tx = 0.f, ty = 0.f;
            if (TAIL) {
                const float4 a = sA[(ch * S + H) * 32 + lane];
                tx = fmaf(a.z, -cs.y, a.x * cs.x);
                ty = fmaf(a.x, cs.y, a.z * cs.x);
                row[TA - 1] = INF;
            }
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float4 a = sA[(ch * S + k) * 32 + lane];
                const uint64_t X2 = pk(a.x, a.y), Y2 = pk(a.z, a.w);
                AX[k] = fma2(Y2, NS2, mul2(X2, C2));  // x' = x cos - y sin
                AY[k] = fma2(X2, S2, mul2(Y2, C2));   // y' = x sin + y cos
                if (XF) {
                    NA[k] = fma2(AX[k], AX[k], mul2(AY[k], AY[k]));
                    const uint64_t M2 = pk(-2.f, -2.f);
                    AX[k] = mul2(AX[k], M2);          // exact scaling
                    AY[k] = mul2(AY[k], M2);
                }
                row[2 * k] = INF;
                row[2 * k + 1] = INF;
            }
            // One step = one float4 of the B block = two reference points against this lane's TA test points; c0 / c1 enter
            // holding the seeds of the two column minima (INF, the tail pass's minima, or the previous chunks' minima).
            auto step = [&](auto last_tag, const float4 B, float c0, float c1, unsigned* col_out) {
                constexpr bool LASTC = decltype(last_tag)::value;   // all test points seen after this chunk
                const uint64_t bx0 = pk(B.x, B.x), by0 = pk(B.y, B.y);
                const uint64_t bx1 = pk(B.z, B.z), by1 = pk(B.w, B.w);
                if (TAIL) {  // scalar FP32 ops for the unpaired point
                    const float ex0 = tx - B.x, ey0 = ty - B.y, ex1 = tx - B.z, ey1 = ty - B.w;
                    const float t0 = fmaf(ex0, ex0, ey0 * ey0), t1 = fmaf(ex1, ex1, ey1 * ey1);
                    row[TA - 1] = min3(row[TA - 1], t0, t1);
                    c0 = (MULTI || TAILP) ? fminf(c0, t0) : t0;  // plain single-chunk: these distances seed the column minima
                    c1 = (MULTI || TAILP) ? fminf(c1, t1) : t1;
                }
                if (XF) {
                    const float2 nb = sNB[(int)(col_out - my_col) >> 1];
                    const uint64_t n0 = pk(nb.x, nb.x), n1 = pk(nb.y, nb.y);
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const uint64_t r0 = fma2(AX[k], bx0, fma2(AY[k], by0, n0));  // |b0|^2 - 2 a.b0 for (a0, a1)
                        const uint64_t r1 = fma2(AX[k], bx1, fma2(AY[k], by1, n1));
                        const uint64_t d0 = add2(r0, NA[k]), d1 = add2(r1, NA[k]);   // + |a|^2: the squared distances
                        float r00, r10, r01, r11, d00, d10, d01, d11;
                        upk(r0, r00, r10);
                        upk(r1, r01, r11);
                        upk(d0, d00, d10);
                        upk(d1, d01, d11);
                        row[2 * k] = min3(row[2 * k], r00, r01);
                        row[2 * k + 1] = min3(row[2 * k + 1], r10, r11);
                        c0 = min3(c0, d00, d10);
                        c1 = min3(c1, d01, d11);
                    }
                    c0 = fmaxf(c0, 0.f);   // cancellation can leave a tiny negative value: the bit-pattern order needs >= 0
                    c1 = fmaxf(c1, 0.f);
                } else {
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const uint64_t dx0 = sub2(AX[k], bx0), dy0 = sub2(AY[k], by0);
                        const uint64_t dx1 = sub2(AX[k], bx1), dy1 = sub2(AY[k], by1);
                        const uint64_t d0 = fma2(dx0, dx0, mul2(dy0, dy0));  // (|a0-b0|^2, |a1-b0|^2)
                        const uint64_t d1 = fma2(dx1, dx1, mul2(dy1, dy1));  // (|a0-b1|^2, |a1-b1|^2)
                        float d00, d10, d01, d11;
                        upk(d0, d00, d10);
                        upk(d1, d01, d11);
                        row[2 * k] = min3(row[2 * k], d00, d01);
                        row[2 * k + 1] = min3(row[2 * k + 1], d10, d11);
                        c0 = min3(c0, d00, d10);
                        c1 = min3(c1, d01, d11);
                    }
                }
                const unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(c0));
                const unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(c1));
                if (LASTC) {
                    colmax = max(colmax, max(r0, r1));  // these are the column minima
                } else if (lane == 0) {
                    *reinterpret_cast<uint2*>(col_out) = make_uint2(r0, r1);
                }
            };
            // The B block is walked two float4 (four reference points) per trip with running pointers: one LDS.128 brings
            // the four seeds, and the addresses cost two pointer increments instead of an index computation per array.
            // Specialised on (last chunk, seeded) so that a chunked unit's loop carries no per-step branches.
            auto walk = [&](auto last_tag, auto seeded_tag) {
                constexpr bool SEEDED = decltype(seeded_tag)::value;   // MULTI: lane 0 alone carries the seeds
                const float4* pB = sB;
                unsigned* pC = my_col;
                const float4* const pB_end2 = sB + (ud.m_pairs & ~1);
#pragma unroll TU
                for (; pB != pB_end2; pB += 2, pC += 4) {
                    const float4 B0 = pB[0], B1 = pB[1];
                    float s0 = INF, s1 = INF, s2 = INF, s3 = INF;
                    if (SEEDED && (TAILP || lane == 0)) {
                        const uint4 sd = *reinterpret_cast<const uint4*>(pC);
                        s0 = __uint_as_float(sd.x), s1 = __uint_as_float(sd.y), s2 = __uint_as_float(sd.z), s3 = __uint_as_float(sd.w);
                    }
                    step(last_tag, B0, s0, s1, pC);
                    // LIST: a column minimum above the unit's best exact distance (+ window) already proves that this
                    // candidate cannot win: stop, and keep the proven lower bound (uniform across the warp)
                    if (LIST && colmax > give_up) {
                        gave_up = true;
                        break;
                    }
                    step(last_tag, B1, s2, s3, pC + 2);
                    if (LIST && colmax > give_up) {
                        gave_up = true;
                        break;
                    }
                }
                if ((ud.m_pairs & 1) && !(LIST && gave_up)) {   // the odd last float4
                    float s0 = INF, s1 = INF;
                    if (SEEDED && (TAILP || lane == 0)) {
                        const uint2 sd = *reinterpret_cast<const uint2*>(pC);
                        s0 = __uint_as_float(sd.x), s1 = __uint_as_float(sd.y);
                    }
                    step(last_tag, *pB, s0, s1, pC);
                    if (LIST && colmax > give_up) gave_up = true;
                }
            };
            using T_ = cuda::std::true_type;
            using F_ = cuda::std::false_type;
            if (!MULTI) {
                if (TAILP && !MMRS_EXP_NOSEED) walk(T_{}, T_{});
                else walk(T_{}, F_{});
            } else if (ud.n_chunks == 1) {
                walk(T_{}, F_{});
            } else if (ch == 0) {
                walk(F_{}, F_{});
            } else if (ch < ud.n_chunks - 1) {
                walk(F_{}, T_{});
            } else {
                walk(T_{}, T_{});
            }
            if (LIST && gave_up) break;  // the row minima are incomplete: only the column bound counts
            if (XF) {   // row minima of r -> squared distances: + |a|^2 (the odd scalar slot is already a distance)
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    float n0, n1;
                    upk(NA[k], n0, n1);
                    row[2 * k] += n0;
                    row[2 * k + 1] += n1;
                }
            }
            float rm = XF ? fmaxf(row[0], 0.f) : row[0];
#pragma unroll
            for (int k = 1; k < TA; ++k) rm = fmaxf(rm, row[k]);
            rowmax = max(rowmax, __float_as_uint(rm));
        }
        if (MULTI) __syncwarp();  // the next candidate reuses my_col
        const unsigned h2 = __reduce_max_sync(0xffffffffu, max(rowmax, colmax));
        const float d = sqrtf(__uint_as_float(h2));
        if (lane == 0) {
            // A candidate that gave up holds only a proven LOWER BOUND (above the unit's window): it is stored as such
            // (mmrs_b200.h: dist32 of pruned / given-up candidates is a bound) but enters neither the error diagnostic
            // nor the unit's key.
            if (LIST && !gave_up) {
                const float old = dist32[ud.dist_off + c];
                worst = fmaxf(worst, fabsf(d * d - old * old));
            }
            dist32[ud.dist_off + c] = d;
            if (!LIST || !gave_up) {
                const unsigned long long k64 = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)c;
                best = min(best, k64);
            }
        }
    }
    if (LIST && lane == 0 && worst > 0.f) {
        const float r = fmaxf(__uint_as_float(rmax_bits[unit]), 1e-30f);
        atomicMax(diag, __float_as_uint(worst / (r * r)));
        worst = 0.f;
    }
    if (lane == 0 && best != ~0ull) atomicMin(s_key, best);
    __syncthreads();
    if (threadIdx.x == 0 && *s_key != ~0ull) atomicMin(&key[unit], *s_key);
    if (!LIST) break;
    g_pos = run_end;
    if (g_pos >= g_hi) break;
    __syncthreads();  // every warp is done with the staged unit and s_key before the next run re-stages
    }
}

// =============================================================================
// K1b: the FP32 sweep for units that do not fit K1's shared-memory staging (more than ~4 000 points per set; the
// reference has no size limit, process_utils.rs:84-121). Same arithmetic as K1's chunked flavour (one warp = one
// candidate, TA = 16 test points per lane and chunk, packed f32x2, 3-input minima, one REDUX per reference point), but
//   * the REFERENCE set is streamed through shared memory in blocks of kBigBlock points (all warps of the CTA walk the
//     blocks together, each with its own candidate); the per-warp column-minimum array only spans one block;
//   * the TEST set is read chunk by chunk from the staging image in global memory (L2) and rotated again per block;
//   * the ROW minima of a candidate persist across the blocks in a per-warp scratch row in global memory
//     ([CTA][warp][slot][lane] floats, coalesced, L2-resident: 2 x 4 B per test point and block).
// Persistent grid: CTA b walks the work items b, b + gridDim.x, ... (its scratch rows are reused).
// =============================================================================
constexpr int kBigBlock = 1024;   // reference points per shared-memory block
constexpr int kBigTA = 16;

__global__ void __launch_bounds__(kThreads, 2)
    k_sweep_big(const UnitDesc* __restrict__ units, const WorkItem* __restrict__ work, int n_work,
                const float4* __restrict__ lay, const float2* __restrict__ cs32, float* __restrict__ dist32,
                unsigned long long* __restrict__ key, float* __restrict__ row_scratch, long long scratch_per_warp) {
    constexpr int TA = kBigTA, H = TA / 2;
    __shared__ __align__(16) float4 sB[kBigBlock / 2];
    __shared__ unsigned s_col[kWarpsPerCta][kBigBlock];
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for ptxas
    const float INF = __int_as_float(0x7f800000);
    float* my_rows = row_scratch + ((long long)blockIdx.x * kWarpsPerCta + wid) * scratch_per_warp;
    unsigned* my_col = s_col[wid];
    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const WorkItem w = work[wi];
        const UnitDesc ud = units[w.unit];
        const float4* gA = lay + ud.lay_off;
        const float4* gB = gA + (long long)ud.n_chunks * H * 32;
        const int n_blocks = (ud.m_pairs + kBigBlock / 2 - 1) / (kBigBlock / 2);
        for (int c0 = 0; c0 < w.count; c0 += kWarpsPerCta) {   // one candidate per warp and round; every warp walks the blocks
            const bool live = c0 + wid < w.count;
            const int c = w.begin + min(c0 + wid, w.count - 1);
            const float2 cs = __ldg(&cs32[ud.cand_off + c]);
            const uint64_t C2 = pk(cs.x, cs.x), S2 = pk(cs.y, cs.y), NS2 = pk(-cs.y, -cs.y);
            unsigned rowmax = 0u, colmax = 0u;
            for (int bb = 0; bb < n_blocks; ++bb) {
                const int j0 = bb * (kBigBlock / 2), jn = min(kBigBlock / 2, ud.m_pairs - j0);
                __syncthreads();   // every warp is done with the previous block
                for (int j = threadIdx.x; j < jn; j += kThreads) sB[j] = gB[j0 + j];
                __syncthreads();
                const bool last_block = bb == n_blocks - 1;
                for (int ch = 0; ch < ud.n_chunks; ++ch) {
                    uint64_t AX[H], AY[H];
                    float row[TA];
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const float4 a = __ldg(&gA[(ch * H + k) * 32 + lane]);
                        const uint64_t X2 = pk(a.x, a.y), Y2 = pk(a.z, a.w);
                        AX[k] = fma2(Y2, NS2, mul2(X2, C2));
                        AY[k] = fma2(X2, S2, mul2(Y2, C2));
                        if (bb == 0) {
                            row[2 * k] = INF;
                            row[2 * k + 1] = INF;
                        } else {   // the row minima over the previous blocks
                            const float2 r = *reinterpret_cast<const float2*>(&my_rows[((ch * H + k) * 32 + lane) * 2]);
                            row[2 * k] = r.x;
                            row[2 * k + 1] = r.y;
                        }
                    }
                    const bool last_chunk = ch == ud.n_chunks - 1;
#pragma unroll 2
                    for (int j = 0; j < jn; ++j) {
                        const float4 B = sB[j];
                        const uint64_t bx0 = pk(B.x, B.x), by0 = pk(B.y, B.y), bx1 = pk(B.z, B.z), by1 = pk(B.w, B.w);
                        float q0 = INF, q1 = INF;
                        if (ch > 0 && lane == 0) {
                            const uint2 prev = *reinterpret_cast<const uint2*>(&my_col[2 * j]);
                            q0 = __uint_as_float(prev.x);
                            q1 = __uint_as_float(prev.y);
                        }
#pragma unroll
                        for (int k = 0; k < H; ++k) {
                            const uint64_t dx0 = sub2(AX[k], bx0), dy0 = sub2(AY[k], by0);
                            const uint64_t dx1 = sub2(AX[k], bx1), dy1 = sub2(AY[k], by1);
                            const uint64_t d0 = fma2(dx0, dx0, mul2(dy0, dy0));
                            const uint64_t d1 = fma2(dx1, dx1, mul2(dy1, dy1));
                            float d00, d10, d01, d11;
                            upk(d0, d00, d10);
                            upk(d1, d01, d11);
                            row[2 * k] = min3(row[2 * k], d00, d01);
                            row[2 * k + 1] = min3(row[2 * k + 1], d10, d11);
                            q0 = min3(q0, d00, d10);
                            q1 = min3(q1, d01, d11);
                        }
                        const unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(q0));
                        const unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(q1));
                        if (last_chunk) colmax = max(colmax, max(r0, r1));
                        else if (lane == 0) *reinterpret_cast<uint2*>(&my_col[2 * j]) = make_uint2(r0, r1);
                    }
                    if (last_block) {
                        float rm = row[0];
#pragma unroll
                        for (int k = 1; k < TA; ++k) rm = fmaxf(rm, row[k]);
                        rowmax = max(rowmax, __float_as_uint(rm));
                    } else {
#pragma unroll
                        for (int k = 0; k < H; ++k)
                            *reinterpret_cast<float2*>(&my_rows[((ch * H + k) * 32 + lane) * 2]) = make_float2(row[2 * k], row[2 * k + 1]);
                    }
                    __syncwarp();   // the next chunk reads the column minima lane 0 just stored
                }
            }
            const unsigned h2 = __reduce_max_sync(0xffffffffu, max(rowmax, colmax));
            const float d = sqrtf(__uint_as_float(h2));
            if (lane == 0 && live) {
                dist32[ud.dist_off + c] = d;
                atomicMin(&key[w.unit], ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)c);
            }
        }
    }
}

// =============================================================================
// K1p: exact lower-bound pruning tier (opt-in, mmrs_sweep_opts.prune). For ANY subsets A' of the test points and B'
// of the reference points,
//     H(A, B) = max( max_{a in A} min_{b in B} |a - b| , max_{b in B} min_{a in A} |a - b| )
//            >= max( max_{a in A'} min_{b in B} |a - b| , max_{b in B'} min_{a in A} |a - b| ) =: LB,
// because a maximum over fewer rows can only be smaller while every row minimum still runs over ALL points of the
// other set. LB costs (|A'| M + |B'| N) pair distances instead of N M. A candidate whose LB exceeds the exact distance
// of any evaluated candidate (plus the FP32 window) cannot be the arg-min and is never scored; every other candidate is
// scored by the exact FP32 kernel (k_sweep LIST) and continues to K2/K3/K4 unchanged, so the selected candidate, its
// angle and its f64 distance are identical to the dense path's.
//   k_prep_lb   per unit two staging images: rows = R sampled test points / columns = all reference points, and
//               rows = R sampled reference points / columns = all test points (rotated by -theta instead). The sample
//               is, per window of n/R consecutive points, the point farthest from the rotation centre: the directed
//               maxima sit where a contour sticks out (measured: 4x fewer survivors than plain striding).
//   k_lb<RS,CB> rows-only sweep of those images, 32 RS rows x CB candidates per warp: row minima + max, no column minima,
//               no REDUX per column; result max-combined into dist32 with atomicMax on the (non-negative) float bits.
//               32 rows for sets below 1 024 points, 128 rows above (dense 2 000-point lumina of consecutive frames
//               differ so little that a coarser sample lets half of the candidates through).
//   k_lb_argmin the candidate with the smallest LB of each unit (scored first: its exact distance is the bound).
// =============================================================================
constexpr int kLbNegSin = 0x100;  // UnitDesc.flags of a lower-bound unit: rotate its rows by -theta

__global__ void k_prep_lb(const UnitDesc* __restrict__ units, const UnitDesc* __restrict__ lb_units, int n_units,
                          const double* __restrict__ test_xy, const double* __restrict__ ref_xy,
                          float4* __restrict__ lay, int R, int R_eff, int pick_mode) {
    const int u = blockIdx.x;
    const UnitDesc ud = units[u];
    if (ud.n <= 0 || ud.m <= 0) return;
    for (int pass = 0; pass < 2; ++pass) {
        const UnitDesc lb = lb_units[u + pass * n_units];
        const double* rows = pass == 0 ? test_xy + 2 * ud.test_off : ref_xy + 2 * ud.ref_off;
        const double* cols = pass == 0 ? ref_xy + 2 * ud.ref_off : test_xy + 2 * ud.test_off;
        const int nr = pass == 0 ? ud.n : ud.m, nc = pass == 0 ? ud.m : ud.n;
        float2* A = reinterpret_cast<float2*>(lay + lb.lay_off);   // R rows (x, y); lane l of k_lb owns row l
        float4* B = lay + lb.lay_off + R / 2;
        // R_eff < R (experiments): only R_eff distinct rows, repeated, to measure how the bound degrades
        auto pick = [&](int r) { return nr <= R ? min(r, nr - 1) : (int)(((long long)(r % R_eff) * nr) / R_eff); };
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            int i = pick(r);
            if (pick_mode == 3 && nr > R) {
                // half strided, half farthest: windows of 2 nr / R points; even rows take the window's first point,
                // odd rows its point farthest from the rotation centre
                const int w = r >> 1, W = R >> 1;
                const int lo = (int)(((long long)w * nr) / W), hi = (int)(((long long)(w + 1) * nr) / W);
                i = lo;
                if (r & 1) {
                    double best = -1.0;
                    for (int k = lo; k < hi; ++k) {
                        const double dx = rows[2 * k] - ud.cx, dy = rows[2 * k + 1] - ud.cy, rr = dx * dx + dy * dy;
                        if (rr > best) best = rr, i = k;
                    }
                }
            } else if (pick_mode != 0 && nr > R) {
                // Any subset gives a valid bound; the directed maxima sit where a contour sticks out or caves in, so
                // take, per index window, the point farthest from (even windows / mode 2: all) or nearest to (odd
                // windows) the rotation centre instead of the window's first point.
                const int lo = (int)(((long long)r * nr) / R), hi = (int)(((long long)(r + 1) * nr) / R);
                const bool want_max = pick_mode == 2 || (r & 1) == 0;
                double best = want_max ? -1.0 : 1e300;
                for (int k = lo; k < hi; ++k) {
                    const double dx = rows[2 * k] - ud.cx, dy = rows[2 * k + 1] - ud.cy, rr = dx * dx + dy * dy;
                    if (want_max ? rr > best : rr < best) best = rr, i = k;
                }
            }
            A[r] = make_float2((float)(rows[2 * i] - ud.cx), (float)(rows[2 * i + 1] - ud.cy));
        }
        for (int j = threadIdx.x; j < lb.m_pairs; j += blockDim.x) {
            const int j0 = min(2 * j, nc - 1), j1 = min(2 * j + 1, nc - 1);
            B[j] = make_float4((float)(cols[2 * j0] - ud.cx), (float)(cols[2 * j0 + 1] - ud.cy), (float)(cols[2 * j1] - ud.cx),
                               (float)(cols[2 * j1 + 1] - ud.cy));
        }
    }
}

// One warp = CB consecutive candidates x 32 RS sampled rows (lane l owns rows l, l + 32, ...): the packed f32x2 lanes
// carry TWO CANDIDATES of the same row, so one broadcast LDS.128 of two column points feeds RS CB/2 = 4 blocks of
// 8 packed FP32 + 2 FMNMX3 — the instruction mix of K1's inner loop without its column side.
template <int RS, int CB>
__global__ void __launch_bounds__(kThreads, 2)
    k_lb(const UnitDesc* __restrict__ lb_units, const WorkItem* __restrict__ work, const float4* __restrict__ lay,
         const float2* __restrict__ cs32, float* __restrict__ dist32) {
    constexpr int P = CB / 2;
    static_assert(RS * CB == 8 && CB >= 2, "rows per lane x candidates per warp");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float4* sA = reinterpret_cast<float4*>(smem_raw + 16);
    const WorkItem w = work[blockIdx.x];
    const UnitDesc ud = lb_units[w.unit];
    const int a_elems = 16 * RS, b_elems = ud.m_pairs;   // 32 RS rows x 8 B
    float4* sB = sA + a_elems;
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for ptxas
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const uint32_t bytes = (uint32_t)(a_elems + b_elems) * 16u;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(sA, lay + ud.lay_off, bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const float INF = __int_as_float(0x7f800000);
    const float sgn = (ud.flags & kLbNegSin) ? -1.f : 1.f;
    unsigned* out = reinterpret_cast<unsigned*>(dist32 + ud.dist_off);
    uint64_t X2[RS], Y2[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const float2 a = reinterpret_cast<const float2*>(sA)[lane + 32 * r];
        X2[r] = pk(a.x, a.x);
        Y2[r] = pk(a.y, a.y);
    }
    for (int g = wid * CB; g < w.count; g += kWarpsPerCta * CB) {
        uint64_t AX[RS][P], AY[RS][P];
        float row[RS][CB];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int c0 = w.begin + min(g + 2 * p, w.count - 1), c1 = w.begin + min(g + 2 * p + 1, w.count - 1);
            const float2 s0 = __ldg(&cs32[ud.cand_off + c0]), s1 = __ldg(&cs32[ud.cand_off + c1]);
            const uint64_t C2 = pk(s0.x, s1.x), S2 = pk(s0.y * sgn, s1.y * sgn), NS2 = pk(-s0.y * sgn, -s1.y * sgn);
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                AX[r][p] = fma2(Y2[r], NS2, mul2(X2[r], C2));   // (x cos0 - y sin0, x cos1 - y sin1)
                AY[r][p] = fma2(X2[r], S2, mul2(Y2[r], C2));
                row[r][2 * p] = INF;
                row[r][2 * p + 1] = INF;
            }
        }
#pragma unroll 2
        for (int j = 0; j < ud.m_pairs; ++j) {
            const float4 B = sB[j];
            const uint64_t bx0 = pk(B.x, B.x), by0 = pk(B.y, B.y), bx1 = pk(B.z, B.z), by1 = pk(B.w, B.w);
#pragma unroll
            for (int r = 0; r < RS; ++r) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint64_t dx0 = sub2(AX[r][p], bx0), dy0 = sub2(AY[r][p], by0);
                    const uint64_t dx1 = sub2(AX[r][p], bx1), dy1 = sub2(AY[r][p], by1);
                    const uint64_t d0 = fma2(dx0, dx0, mul2(dy0, dy0));  // column point 0: (candidate 2p, candidate 2p+1)
                    const uint64_t d1 = fma2(dx1, dx1, mul2(dy1, dy1));  // column point 1
                    float d00, d01, d10, d11;
                    upk(d0, d00, d01);
                    upk(d1, d10, d11);
                    row[r][2 * p] = min3(row[r][2 * p], d00, d10);
                    row[r][2 * p + 1] = min3(row[r][2 * p + 1], d01, d11);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            float rm = row[0][k];
#pragma unroll
            for (int r = 1; r < RS; ++r) rm = fmaxf(rm, row[r][k]);
            const unsigned h2 = __reduce_max_sync(0xffffffffu, __float_as_uint(rm));
            if (lane == k && g + k < w.count) atomicMax(&out[w.begin + g + k], __float_as_uint(sqrtf(__uint_as_float(h2))));
        }
    }
}

// One CTA per unit: the candidate with the smallest lower bound (ties -> lowest index) becomes the unit's one-item list.
__global__ void k_lb_argmin(const UnitDesc* __restrict__ units, const float* __restrict__ dist32, int* __restrict__ l_count,
                            unsigned* __restrict__ l_base, int2* __restrict__ l_items) {
    __shared__ unsigned long long s_best;
    const int u = blockIdx.x;
    const UnitDesc ud = units[u];
    if (threadIdx.x == 0) s_best = ~0ull;
    __syncthreads();
    if (ud.flags || ud.n_cand <= 0) {
        if (threadIdx.x == 0) l_count[u] = 0, l_base[u] = (unsigned)u, l_items[u] = make_int2(-1, 0);
        return;
    }
    unsigned long long best = ~0ull;
    const float* d = dist32 + ud.dist_off;
    for (int c = threadIdx.x; c < ud.n_cand; c += blockDim.x) {
        const unsigned long long k = ((unsigned long long)__float_as_uint(d[c]) << 32) | (unsigned)c;
        best = best < k ? best : k;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = best < other ? best : other;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) {
        l_items[u] = make_int2(u, (int)(s_best & 0xffffffffu));
        l_base[u] = (unsigned)u;
        l_count[u] = 1;
    }
}

// =============================================================================
// K2: shortlist = candidates whose FP32 distance lies inside the FP32 error window above the
// minimum. Two scans of the unit's FP32 row (L2-resident): count, reserve a contiguous range of
// the global item pool with one atomicAdd, then write (unit, candidate) items. A unit may take
// any number of items; only when the POOL is exhausted is it flagged for the dense f64 path.
// =============================================================================
// mode 0: the FP32 window  d <= dmin (1 + rel) + abs_scale Rmax  (tier 2 -> f64 recheck).
// mode 1: the tensor-core prefilter's window in SQUARED distance  d^2 <= dmin^2 + abs_scale Rmax^2  (tier 1 -> exact
//         FP32 re-scoring); `key` is then the prefilter's own minimum.
// prev_count (mode 0, optional): a unit whose tier-1 list overflowed its pool (< 0) was never re-scored in FP32; it is
// passed on as an overflow so the host rechecks all of its candidates in f64.
__global__ void k_shortlist(const UnitDesc* __restrict__ units, const float* __restrict__ dist32,
                            const unsigned long long* __restrict__ key, const unsigned* __restrict__ rmax_bits,
                            float rel, float abs_scale, unsigned pool_cap, int* __restrict__ sl_count,
                            unsigned* __restrict__ sl_base, int2* __restrict__ items, unsigned* __restrict__ n_items,
                            int mode, const int* __restrict__ prev_count) {
    __shared__ int s_n, s_pos;
    __shared__ unsigned s_base;
    const int u = blockIdx.x;
    const UnitDesc ud = units[u];
    if (ud.n_cand <= 0 || ud.n <= 0 || ud.m <= 0 || ud.c_hi <= ud.c_lo) {   // nothing of this unit is swept here
        if (threadIdx.x == 0) {
            sl_count[u] = 0;
            sl_base[u] = 0;
        }
        return;
    }
    if (prev_count && prev_count[u] < 0) {
        if (threadIdx.x == 0) {
            sl_count[u] = -ud.n_cand;
            sl_base[u] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) s_n = 0, s_pos = 0;
    __syncthreads();
    const float dmin = __uint_as_float((unsigned)(key[u] >> 32));
    const float rmax = __uint_as_float(rmax_bits[u]);
    // mode 2 (expanded-form tier): the window of the tier's own absolute error on d^2 (abs_scale Rmax^2), widened by the
    // whole FP32 window of the stage behind it (4e-6 relative + 4e-6 Rmax: twice what k_shortlist mode 0 uses), so that
    // every candidate the exact FP32 values would shortlist is among the re-scored ones whatever the distance is.
    const float thr = mode == 2   ? sqrtf(fmaf(dmin, dmin, abs_scale * rmax * rmax)) * (1.0f + rel) + 4e-6f * rmax
                      : mode == 1 ? sqrtf(fmaf(dmin, dmin, abs_scale * rmax * rmax)) * (1.0f + rel)
                                  : dmin * (1.0f + rel) + abs_scale * rmax;
    const float* d = dist32 + ud.dist_off;
    int mine = 0;
    for (int c = ud.c_lo + threadIdx.x; c < ud.c_hi; c += blockDim.x) mine += (d[c] <= thr) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_n, mine);
    __syncthreads();
    const int n = s_n;
    if (threadIdx.x == 0) s_base = atomicAdd(n_items, (unsigned)n);
    __syncthreads();
    const unsigned base = s_base;
    const bool fits = (unsigned long long)base + (unsigned)n <= pool_cap;
    if (threadIdx.x == 0) {
        sl_base[u] = base;
        sl_count[u] = fits ? n : -n;  // negative: pool exhausted, the host rechecks the whole unit in f64
    }
    if (!fits) {  // neutralise the part of the reservation that lies inside the pool
        for (unsigned k = base + threadIdx.x; k < pool_cap && k < base + (unsigned)n; k += blockDim.x)
            items[k] = make_int2(-1, 0);
        return;
    }
    for (int c = ud.c_lo + threadIdx.x; c < ud.c_hi; c += blockDim.x)
        if (d[c] <= thr) items[base + atomicAdd(&s_pos, 1)] = make_int2(u, c);
}

// =============================================================================
// K3: the reference's f64 arithmetic, literally (no FMA contraction: explicit
// __dmul_rn/__dadd_rn/__dsub_rn; IEEE sqrt). cos/sin come from the HOST's glibc
// (what Rust's f64::sin/cos call), passed as a table, because CUDA's double
// sin/cos are not bit-identical to glibc's.
//   rotate     : contour_point.rs:38-52 (mode 0, identity iff angle == 0.0) /
//                align_between.rs:193-209 (mode 1)
//   distances  : process_utils.rs:84-121, both directions, sqrt, max
// One CTA per (unit, candidate) item; grid-stride over the item list.
// =============================================================================
__device__ __forceinline__ double block_max(double v, double* s_red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = s_red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmax(r, s_red[i]);
    __syncthreads();
    return r;
}

// s_rot / s_ref: the CTA's staging of the rotated test set / the reference set — shared memory, or (units too large for
// it) a per-CTA scratch row in global memory for the rotated points and the caller's reference array itself (copy_ref = false).
__device__ __forceinline__ double exact_cost(const UnitDesc& ud, const double* __restrict__ test_xy,
                                             const double* __restrict__ ref_xy, double cosv, double sinv, bool identity,
                                             double2* s_rot, const double2* s_ref_in, double* s_red, bool copy_ref = true) {
    double2* s_ref_w = const_cast<double2*>(s_ref_in);
    const double2* s_ref = copy_ref ? s_ref_in : reinterpret_cast<const double2*>(ref_xy + 2 * ud.ref_off);
    for (int i = threadIdx.x; i < ud.n; i += blockDim.x) {
        const double px = test_xy[2 * (ud.test_off + i)], py = test_xy[2 * (ud.test_off + i) + 1];
        double rx = px, ry = py;
        if (!identity) {
            const double x = __dsub_rn(px, ud.cx), y = __dsub_rn(py, ud.cy);
            rx = __dadd_rn(__dsub_rn(__dmul_rn(x, cosv), __dmul_rn(y, sinv)), ud.cx);
            ry = __dadd_rn(__dadd_rn(__dmul_rn(x, sinv), __dmul_rn(y, cosv)), ud.cy);
        }
        s_rot[i] = make_double2(rx, ry);
    }
    if (copy_ref)
        for (int j = threadIdx.x; j < ud.m; j += blockDim.x)
            s_ref_w[j] = make_double2(ref_xy[2 * (ud.ref_off + j)], ref_xy[2 * (ud.ref_off + j) + 1]);
    __syncthreads();
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    // forward: reference -> rotated
    double lmax = 0.0;
    for (int i = threadIdx.x; i < ud.m; i += blockDim.x) {
        const double2 a = s_ref[i];
        double mn = INF;
        for (int j = 0; j < ud.n; ++j) {
            const double2 b = s_rot[j];
            const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < mn) mn = d2;
        }
        if (isfinite(mn) && mn > lmax) lmax = mn;
    }
    const double fwd = block_max(lmax, s_red);
    // backward: rotated -> reference
    lmax = 0.0;
    for (int i = threadIdx.x; i < ud.n; i += blockDim.x) {
        const double2 a = s_rot[i];
        double mn = INF;
        for (int j = 0; j < ud.m; ++j) {
            const double2 b = s_ref[j];
            const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < mn) mn = d2;
        }
        if (isfinite(mn) && mn > lmax) lmax = mn;
    }
    const double bwd = block_max(lmax, s_red);
    return fmax(__dsqrt_rn(fwd), __dsqrt_rn(bwd));
}

__global__ void __launch_bounds__(256)
    k_exact(const UnitDesc* __restrict__ units, const double* __restrict__ test_xy, const double* __restrict__ ref_xy,
            const double2* __restrict__ cs64, const unsigned char* __restrict__ zero_flag,
            const int2* __restrict__ items, const unsigned* __restrict__ n_items, unsigned pool_cap,
            double* __restrict__ sl_dist, int max_n, double2* __restrict__ g_scratch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_red = reinterpret_cast<double*>(smem_raw);
    // g_scratch != NULL: the point sets do not fit shared memory — rotated points in this CTA's global scratch row
    double2* s_rot = g_scratch ? g_scratch + (size_t)blockIdx.x * max_n : reinterpret_cast<double2*>(smem_raw + 64);
    double2* s_ref = g_scratch ? nullptr : s_rot + max_n;
    const unsigned total = min(*n_items, pool_cap);
    for (unsigned it = blockIdx.x; it < total; it += gridDim.x) {
        const int2 item = items[it];
        if (item.x < 0) continue;  // uniform per CTA
        const UnitDesc ud = units[item.x];
        const int c = item.y;
        const double2 cs = cs64[ud.cand_off + c];
        const bool identity = zero_flag[ud.cand_off + c] != 0;
        const double d = exact_cost(ud, test_xy, ref_xy, cs.x, cs.y, identity, s_rot, s_ref, s_red, g_scratch == nullptr);
        if (threadIdx.x == 0) sl_dist[it] = d;
        __syncthreads();
    }
}

// Dense variant: every candidate [c0, c0+count) of one unit (shortlist overflow
// path and mmrs_eval_exact).
__global__ void __launch_bounds__(256)
    k_exact_dense(const UnitDesc* __restrict__ units, int unit, const double* __restrict__ test_xy,
                  const double* __restrict__ ref_xy, const double2* __restrict__ cs64,
                  const unsigned char* __restrict__ zero_flag, long long cs_off, int count, double* __restrict__ out,
                  int max_n, double2* __restrict__ g_scratch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_red = reinterpret_cast<double*>(smem_raw);
    double2* s_rot = g_scratch ? g_scratch + (size_t)blockIdx.x * max_n : reinterpret_cast<double2*>(smem_raw + 64);
    double2* s_ref = g_scratch ? nullptr : s_rot + max_n;
    const UnitDesc ud = units[unit];
    for (int c = blockIdx.x; c < count; c += gridDim.x) {
        const double2 cs = cs64[cs_off + c];
        const bool identity = zero_flag[cs_off + c] != 0;
        const double d = exact_cost(ud, test_xy, ref_xy, cs.x, cs.y, identity, s_rot, s_ref, s_red, g_scratch == nullptr);
        if (threadIdx.x == 0) out[c] = d;
        __syncthreads();
    }
}

// =============================================================================
// K4: per unit, leftmost arg-min over the rechecked candidates in f64 (strict <,
// ties -> lowest candidate index: process_utils.rs:69-74) and the tie count.
// One warp per unit.
// =============================================================================
__global__ void k_select(const UnitDesc* __restrict__ units, int n_units, const double2* __restrict__ cs64,
                         const int2* __restrict__ items, const double* __restrict__ sl_dist,
                         const int* __restrict__ sl_count,
                         const unsigned* __restrict__ sl_base, const unsigned long long* __restrict__ key,
                         const unsigned* __restrict__ rmax_bits, double tie_margin, UnitResultDev* __restrict__ res) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n_units) return;
    const int n = sl_count[u];
    UnitResultDev r;
    if (units[u].flags & kFlagRemote) {   // another rank owns this unit: all-zero entry, the merge is a sum
        if (lane == 0) {
            r.best_idx = 0, r.best_dist = 0.0, r.best_d32 = 0.f, r.n_shortlist = 0, r.n_ties = 0, r.flags = 0;
            res[u] = r;
        }
        return;
    }
    r.best_idx = -1;
    r.best_dist = 0.0;
    r.best_d32 = __uint_as_float((unsigned)(key[u] >> 32));
    r.n_shortlist = n;
    r.n_ties = 0;
    r.flags = units[u].flags;
    if (n > 0) {
        const unsigned base = sl_base[u];
        double bd = __longlong_as_double(0x7ff0000000000000LL);
        int bi = 0x7fffffff;
        for (int k = lane; k < n; k += 32) {
            const double d = sl_dist[base + k];
            const int i = items[base + k].y;
            if (d < bd || (d == bd && i < bi)) {
                bd = d;
                bi = i;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) {
                bd = od;
                bi = oi;
            }
        }
        // Tie count: rechecked candidates within the margin whose ANGLE differs from the winner's. A candidate
        // with the winner's own angle (the -pi / +pi wrap duplicates of a +-180 deg grid) has the winner's cost
        // in any arithmetic, so the lowest index wins there regardless and it is not an ambiguity.
        const double lim = bd + tie_margin * fmax(1.0, (double)__uint_as_float(rmax_bits[u]));
        const long long co = units[u].cand_off;
        const double2 wcs = cs64[co + bi];
        int ties = 0;
        for (int k = lane; k < n; k += 32) {
            const int i = items[base + k].y;
            if (i == bi || sl_dist[base + k] > lim) continue;
            const double2 c = cs64[co + i];
            ties += (c.x != wcs.x || c.y != wcs.y) ? 1 : 0;
        }
        for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
        r.best_idx = bi;
        r.best_dist = bd;
        r.n_ties = ties + 1;
    }
    if (lane == 0) res[u] = r;
}

// =============================================================================
// Candidate-axis partition (mmrs_ctx_set_partition 2): every rank swept one contiguous candidate sub-range of every
// unit and holds its LOCAL leftmost f64 arg-min (k_select over its own shortlist, which was cut against the GLOBAL
// FP32 minimum: the packed keys were merged with all-reduce(MIN, uint64) before k_shortlist). res_all = the ranks'
// local results after the all-gather, [world][n_units].
//   k_merge_angle   global leftmost f64 arg-min (lowest distance, ties -> lowest candidate index: process_utils.rs:69-74)
//                   and THIS rank's share of the counts: its shortlist size, its candidates within tie_margin of the
//                   global winner whose angle differs from the winner's, and whether its pool overflowed.
//   k_finish_angle  after the all-reduce(SUM) of those counts.
// One warp per unit.
// =============================================================================
__global__ void k_merge_angle(const UnitDesc* __restrict__ units, int n_units, int world,
                              const UnitResultDev* __restrict__ res_all, const double2* __restrict__ cs64,
                              const int2* __restrict__ items, const double* __restrict__ sl_dist,
                              const int* __restrict__ sl_count, const unsigned* __restrict__ sl_base,
                              const unsigned long long* __restrict__ key, const unsigned* __restrict__ rmax_bits,
                              double tie_margin, UnitResultDev* __restrict__ res, int4* __restrict__ cnt) {
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n_units) return;
    long long bi = -1;
    double bd = 0.0;
    for (int r = 0; r < world; ++r) {   // rank order = candidate order, but compare explicitly
        const UnitResultDev x = res_all[(size_t)r * n_units + u];
        if (x.best_idx < 0) continue;
        if (bi < 0 || x.best_dist < bd || (x.best_dist == bd && x.best_idx < bi)) bi = x.best_idx, bd = x.best_dist;
    }
    const int n = sl_count[u];
    int ties = 0;
    if (bi >= 0 && n > 0) {
        const unsigned base = sl_base[u];
        const double lim = bd + tie_margin * fmax(1.0, (double)__uint_as_float(rmax_bits[u]));
        const long long co = units[u].cand_off;
        const double2 wcs = cs64[co + bi];
        for (int k = lane; k < n; k += 32) {
            const int i = items[base + k].y;
            if (i == bi || sl_dist[base + k] > lim) continue;
            const double2 c = cs64[co + i];
            ties += (c.x != wcs.x || c.y != wcs.y) ? 1 : 0;
        }
        for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
    }
    if (lane == 0) {
        UnitResultDev r;
        r.best_idx = bi;
        r.best_dist = bd;
        r.best_d32 = __uint_as_float((unsigned)(key[u] >> 32));
        r.n_shortlist = 0;
        r.n_ties = 0;
        r.flags = units[u].flags;
        res[u] = r;
        cnt[u] = make_int4(n > 0 ? n : 0, ties, n < 0 ? 1 : 0, 0);
    }
}

__global__ void k_finish_angle(int n_units, const int4* __restrict__ cnt, UnitResultDev* __restrict__ res) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const int4 c = cnt[u];
    // a pool overflow on any rank: every rank rechecks ALL candidates of the unit in f64 at download (negative count)
    res[u].n_shortlist = c.z ? -1 : c.x;
    res[u].n_ties = res[u].best_idx >= 0 ? c.y + 1 : 0;
}

// =============================================================================
// FP32 peak probe: 8 independent FFMA chains per thread, all SMs busy.
// =============================================================================
__global__ void __launch_bounds__(256) k_fp32_probe(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            x0 = fmaf(x0, a, b);
            x1 = fmaf(x1, a, b);
            x2 = fmaf(x2, a, b);
            x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b);
            x5 = fmaf(x5, a, b);
            x6 = fmaf(x6, a, b);
            x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace mmrs
