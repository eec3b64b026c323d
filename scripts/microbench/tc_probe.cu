// tc_probe.cu — feasibility probe for a tensor-core (tcgen05) prefilter tier of the Hausdorff sweep.
//   (1) correctness + accuracy of  D[i][j] = nb_j - 2 a_i . b_j  computed by ONE tcgen05.mma (kind::f16, bf16
//       operands, K = 16) per 128 x N tile from bf16x3-split FP32 coordinates, operands in the no-swizzle K-major
//       canonical layout (8-row x 16-byte core matrices), for both readings of the LBO/SBO descriptor fields;
//   (2) throughput of the epilogue (tcgen05.ld 32x32b + 3-input FMNMX row minima) with 8 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t phase) {  // non-blocking test_wait polling
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// (1) one CTA, 128 threads. imgA / imgB: operand images already in the shared-memory byte layout.
// D (rows_a_pad x ldd floats) = for every 128-row tile of A and every n-tile of B one MMA.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_probe_mma(const uint4* __restrict__ imgA, int bytesA, const uint4* __restrict__ imgB,
                                                   int bytesB, float* __restrict__ D, int ldd, int m_tiles, int n_tiles,
                                                   int NT, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB,
                                                   uint32_t row_group_bytesA, uint32_t row_group_bytesB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* sA = smem;
    unsigned char* sB = smem + ((bytesA + 1023) / 1024) * 1024;
    for (int i = threadIdx.x; i < bytesA / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = imgA[i];
    for (int i = threadIdx.x; i < bytesB / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = imgB[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    for (int mt = 0; mt < m_tiles; ++mt) {
        for (int nt = 0; nt < n_tiles; ++nt) {
            if (threadIdx.x == 0) {
                // 128 rows = 16 groups of 8 rows; NT rows of B = NT/8 groups
                const uint64_t da = make_desc(smem_u32(sA) + (uint32_t)mt * 16u * row_group_bytesA, lboA, sboA);
                const uint64_t db = make_desc(smem_u32(sB) + (uint32_t)nt * (uint32_t)(NT / 8) * row_group_bytesB, lboB, sboB);
                mma_ss(tmem, da, db, make_idesc(128, NT), 0u);
                mma_commit(&bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
            tc_fence_after();
            const int row = mt * 128 + warp * 32 + lane;
            for (int c = 0; c < NT; c += 8) {
                uint32_t v[8];
                tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
                tmem_wait_ld();
                for (int k = 0; k < 8; ++k) D[(size_t)row * ldd + nt * NT + c + k] = __uint_as_float(v[k]);
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(256));
}

// ---------------------------------------------------------------------------------------------------
// (2) epilogue throughput: 8 warps / CTA, 1 CTA / SM; every warp reads `cols` columns of its lane
// quarter and folds them with 3-input minima. MODE 0: ld x32 + 16 min3; 1: ld only; 2: min3 only (same count).
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_probe_epi(float* out, int iters, int cols, long long* cyc) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint((float)(lane * 32 + k));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < cols; c += 32) {
            if (MODE != 2) {
                tmem_ld32(tmem + (uint32_t)c, v);
                tmem_wait_ld();
            }
            if (MODE != 1) {
#pragma unroll
                for (int k = 0; k < 32; k += 8) {
                    m0 = min3(m0, __uint_as_float(v[k]), __uint_as_float(v[k + 1]));
                    m1 = min3(m1, __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
                    m2 = min3(m2, __uint_as_float(v[k + 4]), __uint_as_float(v[k + 5]));
                    m3 = min3(m3, __uint_as_float(v[k + 6]), __uint_as_float(v[k + 7]));
                }
                if (MODE == 2) {  // keep the compiler from hoisting: perturb the inputs
#pragma unroll
                    for (int k = 0; k < 32; k += 8) v[k] ^= (uint32_t)it;
                }
            } else {
                m0 += __uint_as_float(v[it & 31]);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = m0 + m1 + m2 + m3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "n"(512));
}

// (3) the same epilogue loop (ld x32 + 16 FMNMX3 on columns 0..255) while a ninth warp keeps the tensor core busy:
// back-to-back M=128 N=128 K=16 MMAs into columns 256..511, `mma_per_round` of them per commit + wait.
__global__ void __launch_bounds__(288) k_probe_epi_mma(float* out, int iters, long long* cyc, int mma_per_round) {
    __shared__ uint32_t tmem_base_s;
    __shared__ uint64_t bar;
    __shared__ int stop;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f803f80u, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8) {
        if (lane == 0 && mma_per_round % 100 > 0) {
            const uint64_t da = make_desc(smem_u32(smem), 128, 256), db = make_desc(smem_u32(smem) + 8192, 128, 256);
            const uint32_t idesc = make_idesc(128, 128);
            uint32_t phase = 0;
            long long n = 0;
            while (*((volatile int*)&stop) == 0) {
                for (int k = 0; k < mma_per_round % 100; ++k) mma_ss(tmem_base_s + 256 + 128 * (k & 1), da, db, idesc, 0u);
                mma_commit(&bar);
                if (mma_per_round >= 100) mbar_spin(&bar, phase); else mbar_wait(&bar, phase);
                phase ^= 1;
                n += mma_per_round % 100;
            }
            cyc[gridDim.x + blockIdx.x] = n;
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
    uint32_t v[32];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < 128; c += 32) {
            tmem_ld32(tmem + (uint32_t)c, v);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
                m0 = min3(m0, __uint_as_float(v[k]), __uint_as_float(v[k + 1]));
                m1 = min3(m1, __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
                m2 = min3(m2, __uint_as_float(v[k + 4]), __uint_as_float(v[k + 5]));
                m3 = min3(m3, __uint_as_float(v[k + 6]), __uint_as_float(v[k + 7]));
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = m0 + m1 + m2 + m3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    // the eight epilogue warps meet here (named barrier 1), then the tensor-core warp is told to stop
    asm volatile("bar.sync 1, 256;");
    if (threadIdx.x == 0) *((volatile int*)&stop) = 1;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "n"(512));
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
static uint16_t bf16_rn(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}
static float bf16_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static void split3(float x, uint16_t s[3]) {
    s[0] = bf16_rn(x); float r = x - bf16_f(s[0]);
    s[1] = bf16_rn(r); r = r - bf16_f(s[1]);
    s[2] = bf16_rn(r);
}
// dynamic-side row (the rotated test point), role A: last slots (1,1,1,0); role B: (n splits, 0)
static void dyn_row(float x, float y, float nrm, bool with_norm, uint16_t k[16]) {
    uint16_t X[3], Y[3], Nn[3]; split3(x, X); split3(y, Y); split3(nrm, Nn);
    const uint16_t one = bf16_rn(1.0f);
    uint16_t r[16] = {X[0], X[0], X[1], X[1], X[0], X[2], Y[0], Y[0], Y[1], Y[1], Y[0], Y[2],
                      with_norm ? Nn[0] : one, with_norm ? Nn[1] : one, with_norm ? Nn[2] : one, 0};
    memcpy(k, r, 32);
}
// static-side row (reference point), scaled by -2
static void sta_row(float x, float y, float nrm, bool with_norm, uint16_t k[16]) {
    uint16_t X[3], Y[3], Nn[3]; split3(-2.f * x, X); split3(-2.f * y, Y); split3(nrm, Nn);
    const uint16_t one = bf16_rn(1.0f);
    uint16_t r[16] = {X[0], X[1], X[0], X[1], X[2], X[0], Y[0], Y[1], Y[0], Y[1], Y[2], Y[0],
                      with_norm ? Nn[0] : one, with_norm ? Nn[1] : one, with_norm ? Nn[2] : one, 0};
    memcpy(k, r, 32);
}
// image: rows x 16 bf16, group of 8 rows = `group_bytes`; chunk0 at +0, chunk1 at +chunk1_off
static std::vector<unsigned char> make_image(const std::vector<uint16_t>& rows16, int rows, int group_bytes, int chunk1_off) {
    std::vector<unsigned char> img((size_t)(rows / 8) * group_bytes, 0);
    for (int r = 0; r < rows; ++r) {
        unsigned char* g = img.data() + (size_t)(r / 8) * group_bytes + (r % 8) * 16;
        memcpy(g, &rows16[(size_t)r * 16], 16);
        memcpy(g + chunk1_off, &rows16[(size_t)r * 16 + 8], 16);
    }
    return img;
}

int main() {
    const int N = 520, M = 520, NT = 176, n_tiles = 3, m_tiles = 5;
    const int rowsA = m_tiles * 128, rowsB = n_tiles * NT;
    srand(7);
    auto contour = [&](int n, double rot, double r0, double e, std::vector<float>& x, std::vector<float>& y) {
        x.resize(n); y.resize(n);
        for (int i = 0; i < n; ++i) {
            double phi = 2 * M_PI * i / n;
            double r = r0 * (1 + e * cos(2 * phi) + 0.03 * cos(3 * phi + 0.4));
            x[i] = (float)(r * cos(phi + rot) + 0.01 * (rand() / (double)RAND_MAX - 0.5));
            y[i] = (float)(r * sin(phi + rot) + 0.01 * (rand() / (double)RAND_MAX - 0.5));
        }
    };
    std::vector<float> ax, ay, bx, by;
    contour(N, 0.3, 2.8, 0.3, ax, ay);
    contour(M, 0.0, 2.7, 0.28, bx, by);
    const float ang = 0.2931f, cs = cosf(ang), sn = sinf(ang);
    std::vector<float> rx(rowsA), ry(rowsA), na(rowsA), sx(rowsB), sy(rowsB), nb(rowsB);
    for (int i = 0; i < rowsA; ++i) {
        int s = i < N ? i : N - 1;
        rx[i] = fmaf(ay[s], -sn, ax[s] * cs);
        ry[i] = fmaf(ax[s], sn, ay[s] * cs);
        na[i] = fmaf(ax[s], ax[s], ay[s] * ay[s]);
    }
    for (int j = 0; j < rowsB; ++j) {
        int s = j < M ? j : M - 1;
        sx[j] = bx[s]; sy[j] = by[s]; nb[j] = fmaf(bx[s], bx[s], by[s] * by[s]);
    }
    std::vector<uint16_t> A16((size_t)rowsA * 16), B16((size_t)rowsB * 16);
    for (int i = 0; i < rowsA; ++i) dyn_row(rx[i], ry[i], na[i], false, &A16[(size_t)i * 16]);
    for (int j = 0; j < rowsB; ++j) sta_row(sx[j], sy[j], nb[j], true, &B16[(size_t)j * 16]);
    // exact values in double from the FP32 inputs
    std::vector<double> ref((size_t)rowsA * rowsB);
    for (int i = 0; i < rowsA; ++i)
        for (int j = 0; j < rowsB; ++j)
            ref[(size_t)i * rowsB + j] = (double)nb[j] - 2.0 * ((double)rx[i] * sx[j] + (double)ry[i] * sy[j]);

    float* dD; CK(cudaMalloc(&dD, (size_t)rowsA * rowsB * 4));
    std::vector<float> hD((size_t)rowsA * rowsB);
    struct Var { const char* name; int group; int c1off; uint32_t lbo, sbo; };
    const Var vars[] = {
        {"group256 LBO=128 SBO=256", 256, 128, 128, 256},
        {"group256 LBO=256 SBO=128 (fields swapped)", 256, 128, 256, 128},
        {"group384 LBO=128 SBO=384 (chunk1 version A)", 384, 128, 128, 384},
        {"group384 LBO=256 SBO=384 (chunk1 version B)", 384, 256, 256, 384},
    };
    CK(cudaFuncSetAttribute(k_probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (const Var& v : vars) {
        auto imgA = make_image(A16, rowsA, v.group, v.c1off);
        auto imgB = make_image(B16, rowsB, v.group, v.c1off);
        unsigned char *dA, *dB;
        CK(cudaMalloc(&dA, imgA.size())); CK(cudaMalloc(&dB, imgB.size()));
        CK(cudaMemcpy(dA, imgA.data(), imgA.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, imgB.data(), imgB.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, (size_t)rowsA * rowsB * 4));
        k_probe_mma<<<1, 128, 100 * 1024>>>((const uint4*)dA, (int)imgA.size(), (const uint4*)dB, (int)imgB.size(), dD, rowsB,
                                            m_tiles, n_tiles, NT, v.lbo, v.sbo, v.lbo, v.sbo, v.group, v.group);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[mma] %s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0, sum_err = 0; int nan = 0; double max_err_d2 = 0;
        for (int i = 0; i < rowsA; ++i)
            for (int j = 0; j < rowsB; ++j) {
                const double t = hD[(size_t)i * rowsB + j], r = ref[(size_t)i * rowsB + j];
                if (!(t == t)) { ++nan; continue; }
                max_err = fmax(max_err, fabs(t - r)); sum_err += fabs(t - r); max_ref = fmax(max_ref, fabs(r));
                // full squared distance with na added in FP32 like the epilogue will
                const float d2 = (float)t + na[i];
                const double dx = (double)rx[i] - sx[j], dy = (double)ry[i] - sy[j];
                max_err_d2 = fmax(max_err_d2, fabs((double)d2 - (dx * dx + dy * dy)));
            }
        printf("[mma] %-48s max|err| %.3e  mean|err| %.3e  max|ref| %.3f  max|err d2| %.3e (%.1f ulp of 2^-23*Rmax^2=%.3e)  nan %d\n",
               v.name, max_err, sum_err / ((double)rowsA * rowsB), max_ref, max_err_d2,
               max_err_d2 / (ldexp(1.0, -23) * 3.7 * 3.7), ldexp(1.0, -23) * 3.7 * 3.7, nan);
        printf("      sample D[0][0..3] = %.6f %.6f %.6f %.6f | ref %.6f %.6f %.6f %.6f\n", hD[0], hD[1], hD[2], hD[3],
               ref[0], ref[1], ref[2], ref[3]);
        printf("      sample D[300][200..203] = %.6f %.6f %.6f %.6f | ref %.6f %.6f %.6f %.6f\n", hD[300 * rowsB + 200],
               hD[300 * rowsB + 201], hD[300 * rowsB + 202], hD[300 * rowsB + 203], ref[300 * rowsB + 200],
               ref[300 * rowsB + 201], ref[300 * rowsB + 202], ref[300 * rowsB + 203]);
        fflush(stdout);
        cudaFree(dA); cudaFree(dB);
    }
    // (2) epilogue throughput
    {
        int dev_sms = 148;
        cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); dev_sms = p.multiProcessorCount;
        float* out; long long* cyc;
        CK(cudaMalloc(&out, dev_sms * 256 * 4)); CK(cudaMalloc(&cyc, dev_sms * 8));
        const int iters = 2000, cols = 512;
        std::vector<long long> h(dev_sms);
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k_probe_epi<0><<<dev_sms, 256>>>(out, iters, cols, cyc);
                if (mode == 1) k_probe_epi<1><<<dev_sms, 256>>>(out, iters, cols, cyc);
                if (mode == 2) k_probe_epi<2><<<dev_sms, 256>>>(out, iters, cols, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("[epi] mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            CK(cudaMemcpy(h.data(), cyc, dev_sms * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto c : h) avg += (double)c; avg /= dev_sms;
            // per SMSP: 2 warps, each iters * cols/32 chunks of 32 columns
            const double chunks_per_smsp = 2.0 * iters * (cols / 32);
            printf("[epi] mode %d (%s): %.0f cycles, %.2f cycles per 32-column chunk per SMSP, %.3f cycles per element-lane\n", mode,
                   mode == 0 ? "ld x32 + 16 FMNMX3" : mode == 1 ? "ld x32 only" : "16 FMNMX3 only", avg, avg / chunks_per_smsp,
                   avg / chunks_per_smsp / 32.0);
        }
    }
    // (3) epilogue throughput with the tensor core running beside it
    {
        cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
        const int sms = p.multiProcessorCount;
        float* out; long long* cyc;
        CK(cudaMalloc(&out, sms * 256 * 4)); CK(cudaMalloc(&cyc, 2 * sms * 8));
        CK(cudaFuncSetAttribute(k_probe_epi_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
        const int iters = 4000;
        std::vector<long long> h(2 * sms);
        for (int per_round : {0, 1, 4, 101, 104}) {
            CK(cudaMemset(cyc, 0, 2 * sms * 8));
            k_probe_epi_mma<<<sms, 288, 32 * 1024>>>(out, iters, cyc, per_round);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("[epi+mma] CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            CK(cudaMemcpy(h.data(), cyc, 2 * sms * 8, cudaMemcpyDeviceToHost));
            double avg = 0, mm = 0; for (int i = 0; i < sms; ++i) { avg += (double)h[i]; mm += (double)h[sms + i]; }
            avg /= sms; mm /= sms;
            const double chunks_per_smsp = 2.0 * iters * 4;
            printf("[epi+mma] %d MMAs (128x128x16) per commit: epilogue %.2f cycles per 32-column chunk per SMSP; %.0f MMAs in %.0f cycles = %.1f cycles per MMA\n",
                   per_round, avg / chunks_per_smsp, mm, avg, mm > 0 ? avg / mm : 0.0);
        }
    }
    printf("done\n");
    return 0;
}
