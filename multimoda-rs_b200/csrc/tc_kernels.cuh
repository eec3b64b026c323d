// =============================================================================
// tc_kernels.cuh — tensor-core PREFILTER tier of the Hausdorff rotation sweep (sm_100a: tcgen05 + TMEM).
//
//   K1t k_tc_sweep   every candidate of every unit is scored on the 5th-generation tensor cores:
//                    |a' - b|^2 = |a|^2 + |b|^2 - 2 a'.b  (a' = R(theta) a; |a'| = |a|), the cross term and ONE of
//                    the norms as a K = 16 bf16 contraction per point pair. FP32 coordinates are split
//                    into three bf16 terms (hi + mid + lo = 24 mantissa bits) and the six significant cross
//                    products per coordinate are laid out along K, so the products are exact in the FP32
//                    accumulator and the result carries FP32-level error (measured <= 6e-7 * Rmax^2 on d^2,
//                    profiles/r01_tc_probe.txt), not bf16-level error.
//                      P1: rows = rotated test points (changes per candidate), columns = reference points
//                          D1[i][j] = |b_j|^2 - 2 a'_i.b_j   -> row minima + |a_i|^2 -> max = h(A->B)^2
//                      P2: rows = reference points, columns = rotated test points
//                          D2[j][i] = |a_i|^2 - 2 b_j.a'_i   -> row minima + |b_j|^2 -> max = h(B->A)^2
//                    One tcgen05.mma (M = 128, N <= 128, K = 16, cta_group::1) per 128 x N tile, operands in the
//                    no-swizzle K-major canonical layout, accumulators in TMEM (4 stages of 128 columns), read
//                    back with tcgen05.ld.32x32b (one thread = one row) and folded with 3-input FMNMX.
//                    Warp roles: warps 2-5 rotate + split the test points of the
//                    next candidate into shared memory, warps 6-13 (two warpgroups, each with its own pair
//                    of TMEM stages and its own MMA-issuing thread in warps 0-1) are the min/max epilogue. The
//                    issuers read their operand descriptors from small tables built once per CTA: a single
//                    thread that recomputes them per tile is slower than the tensor core AND the epilogue.
//                    The epilogue's FMNMX rate is the bound.
//   k_sweep<.., LIST> (sweep_kernels.cuh) then re-scores, with the exact FP32 arithmetic of K1, only the
//                    candidates whose tensor-core distance lies inside the prefilter's error window of the unit's
//                    minimum; K2/K3/K4 (FP32 window -> reference f64 arithmetic -> leftmost arg-min) are unchanged,
//                    so the selected candidate and its f64 distance are bit-identical to the dense FP32 path.
//
// Replaces (reference): the evaluation loop of search_range, process_utils.rs:69-74, with the cost closures of
// align_within.rs:99-105/:200-206 and align_between.rs:189-216 -> hausdorff_distance, process_utils.rs:78-121.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sweep_kernels.cuh"

namespace mmrs {

constexpr int kTcThreads = 448;    // warps 0-1: MMA issuers, 2-5: producers, 6-13: epilogue (two warpgroups)
constexpr int kTcProducers = 128;
constexpr int kTcGroup = 384;  // bytes per 8 operand rows: [k0..7 | k8..15 variant 1 | k8..15 variant 2], 128 B each
constexpr int kTcMaxPts = 2048;
constexpr int kTcMinPts = 64;

struct TcGeom {  // row bookkeeping of one point set as an MMA operand
    int n;       // valid points
    int rows;    // operand rows (multiple of 128)
    int full;    // rows in full 128-row tiles = (n / 128) * 128
    int tail;    // valid rows in the tail tile (0 = none)
    int shift;   // row offset of the tail's valid rows inside the tail tile (spreads epilogue load over quarters)
    int tailw;   // N-side width of the tail tile (multiple of 16)
    __host__ __device__ static TcGeom make(int n, bool shifted) {
        TcGeom g;
        g.n = n;
        g.full = (n / 128) * 128;
        g.tail = n - g.full;
        g.rows = g.full + (g.tail ? 128 : 0);
        g.tailw = (g.tail + 15) / 16 * 16;
        g.shift = 0;
        if (shifted && g.tail) g.shift = g.tail <= 32 ? 32 : (g.tail <= 64 ? 64 : (g.tail <= 96 ? 32 : 0));
        if (g.shift + g.tailw > 128) g.shift = 128 - g.tailw;
        return g;
    }
    __host__ __device__ int m_tiles() const { return rows / 128; }
    __host__ __device__ int n_tiles() const { return full / 128 + (tail ? 1 : 0); }
    __device__ int src(int row) const {  // operand row -> point index (padding rows repeat the last point)
        if (row < full) return row;
        const int k = row - full - shift;
        return (k >= 0 && k < tail) ? full + k : n - 1;
    }
    __device__ int tile_row0(int t) const { return t * 128 < full ? t * 128 : full + shift; }  // N-side tile start row
    __device__ int tile_w(int t) const { return t * 128 < full ? 128 : tailw; }
    __device__ bool quarter_valid(int mt, int q) const {  // does lane quarter q of M-tile mt hold a non-padding row?
        if (mt * 128 < full) return true;
        const int lo = shift, hi = shift + tail;  // valid rows [lo, hi) of the tail tile
        return q * 32 < hi && q * 32 + 32 > lo;
    }
};

__host__ __device__ inline size_t tc_smem_bytes(int n, int m, int ndyn) {
    const TcGeom a = TcGeom::make(n, false), b = TcGeom::make(m, true);
    return (size_t)(b.rows / 8) * kTcGroup + (size_t)ndyn * (a.rows / 8) * kTcGroup + (size_t)a.rows * 8 + (size_t)a.rows * 4 +
           (size_t)b.rows * 4 + 1024;
}

// ---- tcgen05 / mbarrier helpers -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init_n(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo) {  // no swizzle, K-major, SBO = kTcGroup
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(kTcGroup >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t tc_idesc(int n) {  // kind::f16: bf16 x bf16 -> f32, K-major A and B, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bf16x3 splitting --------------------------------------------------------------------------
// cvt.rn.bf16x2.f32 d, a, b : d = {bf16(a) in the upper half, bf16(b) in the lower half}
__device__ __forceinline__ uint32_t bf2(float hi, float lo) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
struct Split3 {  // packed (x | y) bf16 pairs of the three terms: x in the upper half, y in the lower half
    uint32_t H, M, L;
};
__device__ __forceinline__ Split3 split_xy(float x, float y) {
    Split3 s;
    s.H = bf2(x, y);
    const float r1x = x - __uint_as_float(s.H & 0xffff0000u), r1y = y - __uint_as_float(s.H << 16);  // exact
    s.M = bf2(r1x, r1y);
    const float r2x = r1x - __uint_as_float(s.M & 0xffff0000u), r2y = r1y - __uint_as_float(s.M << 16);
    s.L = bf2(r2x, r2y);
    return s;
}
// three bf16 terms of a non-negative norm given in double: (n0 | n1 << 16, n2)
__device__ __forceinline__ uint2 split_norm(double v) {
    const uint32_t a = bf2(0.f, (float)v) & 0xffffu;
    const double r1 = v - (double)__uint_as_float(a << 16);
    const uint32_t b = bf2(0.f, (float)r1) & 0xffffu;
    const double r2 = r1 - (double)__uint_as_float(b << 16);
    const uint32_t c = bf2(0.f, (float)r2) & 0xffffu;
    return make_uint2(a | (b << 16), c);
}
constexpr uint32_t kBf16One2 = 0x3F803F80u, kBf16One1 = 0x00003F80u;

struct TcShared {  // small control block at the end of the dynamic shared memory
    uint64_t dyn_full[2], dyn_empty[2], tmem_full[4], tmem_empty[4];
    unsigned long long key;
    uint32_t tmem_base;
    unsigned cand_val[8], cand_cnt[8];
    // descriptor tables of the MMA issuers (low descriptor words; dynamic-buffer entries are relative to the buffer)
    uint2 job_a[32];   // per M-tile job: {A descriptor low word, 1 if A is the dynamic operand (P1)}
    uint2 col_s[16];   // P1 column tiles (static operand): {B descriptor low word, instruction descriptor}
    uint2 col_d[16];   // P2 column tiles (dynamic operand): {B descriptor low word (relative), instruction descriptor}
};

// =============================================================================
// K1t
// =============================================================================
__global__ void __launch_bounds__(kTcThreads, 1)
    k_tc_sweep(const UnitDesc* __restrict__ units, const WorkItem* __restrict__ work, const double* __restrict__ test_xy,
               const double* __restrict__ ref_xy, const float2* __restrict__ cs32, float* __restrict__ dist32,
               unsigned long long* __restrict__ key_tc, long long* __restrict__ trace) {
    // trace (experiments only, MMRS_TC_TRACE): clock64 stamps of CTA 0, [wg][kind][tile < 256];
    // kind 0 issuer after its stage wait, 1 issuer after commit, 2 epilogue (quarter 0) woken, 3 epilogue arrives
    const bool tracing = trace != nullptr && blockIdx.x == 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WorkItem w = work[blockIdx.x];
    const UnitDesc ud = units[w.unit];
    const TcGeom ga = TcGeom::make(ud.n, false), gb = TcGeom::make(ud.m, true);
    const int ndyn = ga.rows <= 1024 ? 2 : 1;
    unsigned char* s_static = smem_raw;
    unsigned char* s_dyn = s_static + (size_t)(gb.rows / 8) * kTcGroup;
    const uint32_t dyn_stride = (uint32_t)(ga.rows / 8) * kTcGroup;
    float2* s_pts = reinterpret_cast<float2*>(s_dyn + (size_t)ndyn * dyn_stride);
    float* s_na = reinterpret_cast<float*>(s_pts + ga.rows);
    float* s_nb = s_na + ga.rows;
    TcShared* sh = reinterpret_cast<TcShared*>(s_nb + gb.rows);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- setup -------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init_n(&sh->dyn_full[i], kTcProducers), mbar_init_n(&sh->dyn_empty[i], 2);
        for (int i = 0; i < 4; ++i) mbar_init_n(&sh->tmem_full[i], 1), mbar_init_n(&sh->tmem_empty[i], 4);
        sh->key = ~0ull;
        for (int i = 0; i < 8; ++i) sh->cand_val[i] = 0u, sh->cand_cnt[i] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // test points: centred FP32 coordinates, norms, and the constant halves of the operand rows
    for (int i = tid; i < ga.rows; i += kTcThreads) {
        const int s = ga.src(i);
        const float x = (float)(test_xy[2 * (ud.test_off + s)] - ud.cx), y = (float)(test_xy[2 * (ud.test_off + s) + 1] - ud.cy);
        s_pts[i] = make_float2(x, y);
        const double nd = (double)x * x + (double)y * y;
        s_na[i] = (float)nd;
        const uint2 ns = split_norm(nd);
        for (int b = 0; b < ndyn; ++b) {
            unsigned char* row = s_dyn + (size_t)b * dyn_stride + (size_t)(i >> 3) * kTcGroup + (i & 7) * 16;
            *reinterpret_cast<uint2*>(row + 128 + 8) = make_uint2(kBf16One2, kBf16One1);  // role A (P1 rows): x 1
            *reinterpret_cast<uint2*>(row + 256 + 8) = ns;                                  // role B (P2 columns): |a|^2
        }
    }
    // reference points: the whole static operand (-2 b split into bf16x3; |b|^2 for P1, ones for P2)
    for (int j = tid; j < gb.rows; j += kTcThreads) {
        const int s = gb.src(j);
        const float x = (float)(ref_xy[2 * (ud.ref_off + s)] - ud.cx), y = (float)(ref_xy[2 * (ud.ref_off + s) + 1] - ud.cy);
        const double nd = (double)x * x + (double)y * y;
        s_nb[j] = (float)nd;
        const Split3 t = split_xy(-2.f * x, -2.f * y);
        unsigned char* row = s_static + (size_t)(j >> 3) * kTcGroup + (j & 7) * 16;
        const uint32_t xhm = __byte_perm(t.H, t.M, 0x7632), xlh = __byte_perm(t.L, t.H, 0x7632);
        const uint32_t yhm = __byte_perm(t.H, t.M, 0x5410), ylh = __byte_perm(t.L, t.H, 0x5410);
        *reinterpret_cast<uint4*>(row) = make_uint4(xhm, xhm, xlh, yhm);
        const uint2 ns = split_norm(nd);
        *reinterpret_cast<uint4*>(row + 128) = make_uint4(yhm, ylh, ns.x, ns.y);            // P1 columns: + |b|^2
        *reinterpret_cast<uint4*>(row + 256) = make_uint4(yhm, ylh, kBf16One2, kBf16One1);  // P2 rows: x 1
    }
    const int MA = ga.m_tiles(), MB = gb.m_tiles(), J = MA + MB;
    const int NTA = ga.n_tiles(), NTB = gb.n_tiles();
    {   // issuer tables
        const uint32_t a_static = smem_u32(s_static);
        for (int j = tid; j < J; j += kTcThreads) {
            const bool p1 = j < MA;
            const int mt = p1 ? j : j - MA;
            const uint32_t off = (uint32_t)mt * 16u * kTcGroup;
            sh->job_a[j] = p1 ? make_uint2((off >> 4) | ((128u >> 4) << 16), 1u)
                              : make_uint2(((a_static + off) >> 4) | ((256u >> 4) << 16), 0u);
        }
        for (int t = tid; t < NTB; t += kTcThreads)
            sh->col_s[t] = make_uint2(((a_static + (uint32_t)(gb.tile_row0(t) >> 3) * kTcGroup) >> 4) | ((128u >> 4) << 16),
                                      tc_idesc(gb.tile_w(t)));
        for (int t = tid; t < NTA; t += kTcThreads)
            sh->col_d[t] = make_uint2((((uint32_t)(ga.tile_row0(t) >> 3) * kTcGroup) >> 4) | ((256u >> 4) << 16),
                                      tc_idesc(ga.tile_w(t)));
    }
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;

    if (warp < 2) {
        // ===== MMA issuers: warp g feeds warpgroup g through TMEM stages {g, g + 2} (one thread each) =====
        if (lane == 0) {
            const int g = warp;
            const uint64_t hi = (uint64_t)((uint32_t)(kTcGroup >> 4) | (1u << 14)) << 32;  // SBO, descriptor version 1
            const uint32_t dyn0 = smem_u32(s_dyn) >> 4, dyn_step = dyn_stride >> 4;
            uint32_t cnt = 0;
            for (int ci = 0; ci < w.count; ++ci) {
                const int b = ci % ndyn;
                mbar_wait(&sh->dyn_full[b], (uint32_t)((ci / ndyn) & 1));
                tc_fence_after();
                const uint32_t dyn_lo = dyn0 + (uint32_t)b * dyn_step;
                for (int j = g; j < J; j += 2) {
                    const uint2 ja = sh->job_a[j];
                    const bool p1 = ja.y != 0u;
                    const uint64_t da = hi | (uint64_t)(ja.x + (p1 ? dyn_lo : 0u));
                    const uint2* col = p1 ? sh->col_s : sh->col_d;
                    const uint32_t add_b = p1 ? 0u : dyn_lo;
                    const int nt = p1 ? NTB : NTA;
                    for (int t = 0; t < nt; ++t) {
                        const uint2 cb = col[t];
                        const uint32_t st = (uint32_t)g + 2u * (cnt & 1u), use = cnt >> 1;
                        if (use > 0) {
                            mbar_wait(&sh->tmem_empty[st], (use - 1) & 1u);
                            tc_fence_after();
                        }
                        if (tracing && cnt < 256) trace[(g * 4 + 0) * 256 + cnt] = clock64();
                        tc_mma(tmem + st * 128u, da, hi | (uint64_t)(cb.x + add_b), cb.y);
                        tc_commit(&sh->tmem_full[st]);
                        if (tracing && cnt < 256) trace[(g * 4 + 1) * 256 + cnt] = clock64();
                        ++cnt;
                    }
                }
                tc_commit(&sh->dyn_empty[b]);  // arrives when every MMA this thread issued on the buffer is done
            }
        }
    } else if (warp < 6) {
        // ===== producers: rotate + split the test points of candidate ci into buffer ci % ndyn =====
        const int pid = tid - 64;
        for (int ci = 0; ci < w.count; ++ci) {
            const int b = ci % ndyn, use = ci / ndyn;
            if (use > 0) mbar_wait(&sh->dyn_empty[b], (uint32_t)((use - 1) & 1));
            const float2 cs = __ldg(&cs32[ud.cand_off + w.begin + ci]);
            unsigned char* base = s_dyn + (size_t)b * dyn_stride;
            for (int i = pid; i < ga.rows; i += kTcProducers) {
                const float2 a = s_pts[i];
                const float x = fmaf(a.y, -cs.y, a.x * cs.x), y = fmaf(a.x, cs.y, a.y * cs.x);  // as K1 rotates
                const Split3 t = split_xy(x, y);
                unsigned char* row = base + (size_t)(i >> 3) * kTcGroup + (i & 7) * 16;
                *reinterpret_cast<uint4*>(row) = make_uint4(__byte_perm(t.H, 0, 0x3232), __byte_perm(t.M, 0, 0x3232),
                                                            __byte_perm(t.H, t.L, 0x7632), __byte_perm(t.H, 0, 0x1010));
                const uint2 y2 = make_uint2(__byte_perm(t.M, 0, 0x1010), __byte_perm(t.H, t.L, 0x5410));
                *reinterpret_cast<uint2*>(row + 128) = y2;
                *reinterpret_cast<uint2*>(row + 256) = y2;
            }
            proxy_fence_async();
            mbar_arrive(&sh->dyn_full[b]);
        }
    } else {
        // ===== epilogue: row minima of every tile, + norm, max over rows; one value per candidate =====
        const int wg = (warp - 6) >> 2, q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float INF = __int_as_float(0x7f800000);
        uint32_t cnt = 0;
        unsigned long long best = ~0ull;
        for (int ci = 0; ci < w.count; ++ci) {
            float cmax = 0.f;
            for (int j = wg; j < J; j += 2) {
                const bool p1 = j < MA;
                const int mt = p1 ? j : j - MA;
                const uint2* col = p1 ? sh->col_s : sh->col_d;  // the column side
                const int nt = p1 ? NTB : NTA;
                const bool valid = (p1 ? ga : gb).quarter_valid(mt, q);
                float m0 = INF, m1 = INF, m2 = INF, m3 = INF;
                for (int t = 0; t < nt; ++t) {
                    const uint32_t st = (uint32_t)wg + 2u * (cnt & 1u), use = cnt >> 1;
                    mbar_wait(&sh->tmem_full[st], use & 1u);
                    tc_fence_after();
                    if (tracing && q == 0 && lane == 0 && cnt < 256) trace[(wg * 4 + 2) * 256 + cnt] = clock64();
                    if (valid) {
                        const int wdt = (int)((col[t].y >> 17) & 0x3fu) << 3;
                        const uint32_t taddr = tmem + st * 128u + lane_base;
                        uint32_t va[32], vb[32];
                        auto fold32 = [&](const uint32_t* v) {
#pragma unroll
                            for (int k = 0; k < 32; k += 8) {
                                m0 = min3(m0, __uint_as_float(v[k]), __uint_as_float(v[k + 1]));
                                m1 = min3(m1, __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
                                m2 = min3(m2, __uint_as_float(v[k + 4]), __uint_as_float(v[k + 5]));
                                m3 = min3(m3, __uint_as_float(v[k + 6]), __uint_as_float(v[k + 7]));
                            }
                        };
                        if (wdt == 128) {  // software-pipelined: the next 32 columns are in flight while 32 are folded
                            const bool tr2 = tracing && wg == 0 && q == 0 && lane == 0 && cnt < 64;
                            long long* t2 = trace + 2048 + cnt * 8;
                            if (tr2) t2[0] = clock64();
                            tmem_ld32(taddr, va);
                            tmem_wait_ld();
                            if (tr2) t2[1] = clock64();
                            tmem_ld32(taddr + 32, vb);
                            fold32(va);
                            if (tr2) t2[2] = clock64();
                            tmem_wait_ld();
                            if (tr2) t2[3] = clock64();
                            tmem_ld32(taddr + 64, va);
                            fold32(vb);
                            tmem_wait_ld();
                            tmem_ld32(taddr + 96, vb);
                            fold32(va);
                            tmem_wait_ld();
                            if (tr2) t2[4] = clock64();
                            fold32(vb);
                            if (tr2) t2[5] = clock64();
                        } else {
                            int c = 0;
                            for (; c + 32 <= wdt; c += 32) {
                                tmem_ld32(taddr + c, va);
                                tmem_wait_ld();
                                fold32(va);
                            }
                            if (c < wdt) {  // 16 columns left
                                tmem_ld16(taddr + c, va);
                                tmem_wait_ld();
#pragma unroll
                                for (int k = 0; k < 16; k += 8) {
                                    m0 = min3(m0, __uint_as_float(va[k]), __uint_as_float(va[k + 1]));
                                    m1 = min3(m1, __uint_as_float(va[k + 2]), __uint_as_float(va[k + 3]));
                                    m2 = min3(m2, __uint_as_float(va[k + 4]), __uint_as_float(va[k + 5]));
                                    m3 = min3(m3, __uint_as_float(va[k + 6]), __uint_as_float(va[k + 7]));
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    if (tracing && wg == 0 && q == 0 && lane == 0 && cnt < 64) trace[2048 + cnt * 8 + 6] = clock64();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh->tmem_empty[st]);
                    if (tracing && q == 0 && lane == 0 && cnt < 256) trace[(wg * 4 + 3) * 256 + cnt] = clock64();
                    ++cnt;
                }
                if (valid) {
                    const int row = mt * 128 + q * 32 + lane;
                    const float rowmin = fminf(fminf(m0, m1), fminf(m2, m3));
                    cmax = fmaxf(cmax, rowmin + (p1 ? s_na[row] : s_nb[row]));
                }
            }
            // one value per candidate: max over the 8 epilogue warps through a small ring of shared slots
            const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(cmax, 0.f)));
            if (lane == 0) {
                const int slot = ci & 7;
                atomicMax(&sh->cand_val[slot], bits);
                __threadfence_block();
                if (atomicAdd(&sh->cand_cnt[slot], 1u) == 7u) {
                    const unsigned h2 = atomicExch(&sh->cand_val[slot], 0u);
                    atomicExch(&sh->cand_cnt[slot], 0u);
                    const int c = w.begin + ci;
                    const float d = sqrtf(__uint_as_float(h2));
                    dist32[ud.dist_off + c] = d;
                    const unsigned long long k64 = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)c;
                    best = best < k64 ? best : k64;
                }
            }
        }
        if (lane == 0 && best != ~0ull) atomicMin(&sh->key, best);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && sh->key != ~0ull) atomicMin(&key_tc[w.unit], sh->key);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace mmrs
