// Round-2 inner-loop study for k_sweep on sm_100a: direct vs EXPANDED squared distance
// (|a|^2 + |b|^2 - 2 a.b), a-packed vs b-packed operands, uniform vs lane-local origin.
// Every variant keeps the full reduction structure of K1: 16 test points per lane, row minima in
// registers, column minima + one CREDUX per reference point, reference points streamed from
// shared memory. Output: SMSP-cycles per j-iteration (2 reference points x 16 test points per
// lane = 32 point pairs per lane), 4 warps per SMSP like the shipped kernel (2 CTAs x 8 warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int H = 8;  // packed pairs of test points per lane (TA = 16)

// V0: the shipped direct form, a-packed.
template <int JU>
__global__ void __launch_bounds__(256, 2) k_direct(float* out, int m_pairs, int reps, float s) {
    __shared__ float4 sB[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sB[i] = make_float4(s * i, s * i + 1, s * i + 2, s * i + 3);
    __syncthreads();
    uint64_t AX[H], AY[H];
    float row[2 * H];
    for (int k = 0; k < H; ++k) {
        AX[k] = pk(threadIdx.x * 0.01f + k, threadIdx.x * 0.02f + k);
        AY[k] = pk(threadIdx.x * 0.03f + k, threadIdx.x * 0.04f + k);
        row[2 * k] = row[2 * k + 1] = 1e30f;
    }
    unsigned colmax = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll JU
        for (int j = 0; j < m_pairs; ++j) {
            const float4 B = sB[j];
            const uint64_t bx0 = pk(B.x, B.x), by0 = pk(B.y, B.y), bx1 = pk(B.z, B.z), by1 = pk(B.w, B.w);
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const uint64_t dx0 = sub2(AX[k], bx0), dy0 = sub2(AY[k], by0);
                const uint64_t dx1 = sub2(AX[k], bx1), dy1 = sub2(AY[k], by1);
                const uint64_t d0 = fma2(dx0, dx0, mul2(dy0, dy0));
                const uint64_t d1 = fma2(dx1, dx1, mul2(dy1, dy1));
                float d00, d10, d01, d11;
                upk(d0, d00, d10); upk(d1, d01, d11);
                row[2 * k] = min3(row[2 * k], d00, d01);
                row[2 * k + 1] = min3(row[2 * k + 1], d10, d11);
                c0 = min3(c0, d00, d10);
                c1 = min3(c1, d01, d11);
            }
            unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(c0));
            unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(c1));
            colmax = max(colmax, max(r0, r1));
        }
    }
    float acc = __uint_as_float(colmax);
    for (int k = 0; k < 2 * H; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// V1: expanded form, a-packed, ONE origin for the whole unit. Shared memory holds per reference point
// (-2 bx, -2 by, |b|^2, pad). LOCAL = 1 adds the lane-local origin: per (lane, reference point)
// b' = b - o_lane and |b'|^2 are formed in registers (1 FADD2 + FMUL + FFMA).
template <int JU, int LOCAL>
__global__ void __launch_bounds__(256, 2) k_expanded(float* out, int m_pairs, int reps, float s) {
    __shared__ float4 sB[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sB[i] = make_float4(s * i, s * i + 1, s * i + 2, s * i + 3);
    __syncthreads();
    uint64_t AX[H], AY[H], NA[H];   // (-2 a'x | a'x), (-2 a'y | a'y), |a'|^2
    float row[2 * H];
    for (int k = 0; k < H; ++k) {
        AX[k] = pk(threadIdx.x * 0.01f + k, threadIdx.x * 0.02f + k);
        AY[k] = pk(threadIdx.x * 0.03f + k, threadIdx.x * 0.04f + k);
        NA[k] = pk(threadIdx.x * 0.05f + k, threadIdx.x * 0.06f + k);
        row[2 * k] = row[2 * k + 1] = 1e30f;
    }
    const uint64_t O2 = pk(threadIdx.x * 0.001f, threadIdx.x * 0.002f);
    unsigned colmax = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll JU
        for (int j = 0; j < m_pairs; ++j) {
            const float4 B0 = sB[2 * j], B1 = sB[2 * j + 1];
            float bx0 = B0.x, by0 = B0.y, nb0 = B0.z, bx1 = B1.x, by1 = B1.y, nb1 = B1.z;
            if (LOCAL) {
                float x, y;
                upk(sub2(pk(B0.x, B0.y), O2), x, y);
                bx0 = x, by0 = y, nb0 = fmaf(x, x, y * y);
                upk(sub2(pk(B1.x, B1.y), O2), x, y);
                bx1 = x, by1 = y, nb1 = fmaf(x, x, y * y);
            }
            const uint64_t X0 = pk(bx0, bx0), Y0 = pk(by0, by0), N0 = pk(nb0, nb0);
            const uint64_t X1 = pk(bx1, bx1), Y1 = pk(by1, by1), N1 = pk(nb1, nb1);
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const uint64_t r0 = fma2(AX[k], X0, fma2(AY[k], Y0, N0));   // |b0|^2 - 2 a.b0  (a0 | a1)
                const uint64_t r1 = fma2(AX[k], X1, fma2(AY[k], Y1, N1));
                const uint64_t d0 = add2(r0, NA[k]), d1 = add2(r1, NA[k]);  // full squared distances
                float r00, r10, r01, r11, d00, d10, d01, d11;
                upk(r0, r00, r10); upk(r1, r01, r11); upk(d0, d00, d10); upk(d1, d01, d11);
                row[2 * k] = min3(row[2 * k], r00, r01);
                row[2 * k + 1] = min3(row[2 * k + 1], r10, r11);
                c0 = min3(c0, d00, d10);
                c1 = min3(c1, d01, d11);
            }
            // d^2 may be slightly negative in the expanded form: order as signed ints after a clamp at 0 is not needed
            // for timing; the real kernel clamps once per column (1 FMNMX per reference point).
            unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(c0, 0.f)));
            unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(c1, 0.f)));
            colmax = max(colmax, max(r0, r1));
        }
    }
    float acc = __uint_as_float(colmax);
    for (int k = 0; k < 2 * H; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// V3: expanded form, b-packed: one packed instruction = one test point (scalar operands) x two reference points
// (the packed B operands are identical for all 16 test points of the lane: operand-reuse friendly).
template <int JU, int LOCAL>
__global__ void __launch_bounds__(256, 2) k_expanded_bp(float* out, int m_pairs, int reps, float s) {
    __shared__ float4 sB[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sB[i] = make_float4(s * i, s * i + 1, s * i + 2, s * i + 3);
    __syncthreads();
    constexpr int TA = 2 * H;
    float ax[TA], ay[TA], na[TA], row[TA];
    for (int k = 0; k < TA; ++k) {
        ax[k] = threadIdx.x * 0.01f + k; ay[k] = threadIdx.x * 0.03f + k; na[k] = threadIdx.x * 0.05f + k; row[k] = 1e30f;
    }
    const float ox = threadIdx.x * 0.001f, oy = threadIdx.x * 0.002f;
    unsigned colmax = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll JU
        for (int j = 0; j < m_pairs; ++j) {
            const float4 Ba = sB[2 * j], Bb = sB[2 * j + 1];   // (bx0, bx1, by0, by1), (nb0, nb1, -, -)
            uint64_t BX = pk(Ba.x, Ba.y), BY = pk(Ba.z, Ba.w), NB = pk(Bb.x, Bb.y);
            if (LOCAL) {
                BX = sub2(BX, pk(ox, ox));
                BY = sub2(BY, pk(oy, oy));
                NB = fma2(BX, BX, mul2(BY, BY));
            }
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int k = 0; k < TA; k += 2) {
                const uint64_t ra = fma2(pk(ax[k], ax[k]), BX, fma2(pk(ay[k], ay[k]), BY, NB));
                const uint64_t rb = fma2(pk(ax[k + 1], ax[k + 1]), BX, fma2(pk(ay[k + 1], ay[k + 1]), BY, NB));
                const uint64_t da = add2(ra, pk(na[k], na[k])), db = add2(rb, pk(na[k + 1], na[k + 1]));
                float ra0, ra1, rb0, rb1, da0, da1, db0, db1;
                upk(ra, ra0, ra1); upk(rb, rb0, rb1); upk(da, da0, da1); upk(db, db0, db1);
                row[k] = min3(row[k], ra0, ra1);
                row[k + 1] = min3(row[k + 1], rb0, rb1);
                c0 = min3(c0, da0, db0);
                c1 = min3(c1, da1, db1);
            }
            unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(c0, 0.f)));
            unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(c1, 0.f)));
            colmax = max(colmax, max(r0, r1));
        }
    }
    float acc = __uint_as_float(colmax);
    for (int k = 0; k < TA; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// V4: expanded, a-packed, rows use r, columns use r too but with |a|^2 folded in by a second accumulator set:
// cols: min over a of (r + na) -- here tested as "mins only on r" to see what the FADD2 costs (NOT a valid kernel).
template <int JU>
__global__ void __launch_bounds__(256, 2) k_expanded_noadd(float* out, int m_pairs, int reps, float s) {
    __shared__ float4 sB[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sB[i] = make_float4(s * i, s * i + 1, s * i + 2, s * i + 3);
    __syncthreads();
    uint64_t AX[H], AY[H];
    float row[2 * H];
    for (int k = 0; k < H; ++k) {
        AX[k] = pk(threadIdx.x * 0.01f + k, threadIdx.x * 0.02f + k);
        AY[k] = pk(threadIdx.x * 0.03f + k, threadIdx.x * 0.04f + k);
        row[2 * k] = row[2 * k + 1] = 1e30f;
    }
    unsigned colmax = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll JU
        for (int j = 0; j < m_pairs; ++j) {
            const float4 B0 = sB[2 * j], B1 = sB[2 * j + 1];
            const uint64_t X0 = pk(B0.x, B0.x), Y0 = pk(B0.y, B0.y), N0 = pk(B0.z, B0.z);
            const uint64_t X1 = pk(B1.x, B1.x), Y1 = pk(B1.y, B1.y), N1 = pk(B1.z, B1.z);
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const uint64_t r0 = fma2(AX[k], X0, fma2(AY[k], Y0, N0));
                const uint64_t r1 = fma2(AX[k], X1, fma2(AY[k], Y1, N1));
                float r00, r10, r01, r11;
                upk(r0, r00, r10); upk(r1, r01, r11);
                row[2 * k] = min3(row[2 * k], r00, r01);
                row[2 * k + 1] = min3(row[2 * k + 1], r10, r11);
                c0 = min3(c0, r00, r10);
                c1 = min3(c1, r01, r11);
            }
            unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(c0));
            unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(c1));
            colmax = max(colmax, max(r0, r1));
        }
    }
    float acc = __uint_as_float(colmax);
    for (int k = 0; k < 2 * H; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
void run(const char* name, K kern) {
    float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
    const int m_pairs = 256, reps = 200;
    kern<<<148 * 2, 256>>>(out, m_pairs, 2, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int t = 0; t < 3; ++t) {
        cudaEventRecord(a);
        kern<<<148 * 2, 256>>>(out, m_pairs, reps, 1.0f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    const double jit = (double)reps * m_pairs * 4;  // j-iterations per SMSP (4 warps each)
    const double cyc = best * 1e-3 * 1.965e9 / jit;
    printf("%-58s %8.3f ms  %7.2f SMSP-cycles per j-iteration (32 pairs/lane)  = %.3f of the 5-cycle/pair reference (160)%s\n",
           name, best, cyc, 160.0 / cyc, e == cudaSuccess ? "" : "  CUDA ERROR");
    cudaFree(out);
}

int main() {
    run("V0 direct a-packed JU=2 (shipped form)", k_direct<2>);
    run("V0 direct a-packed JU=1", k_direct<1>);
    run("V1 expanded a-packed uniform origin JU=1", k_expanded<1, 0>);
    run("V1 expanded a-packed uniform origin JU=2", k_expanded<2, 0>);
    run("V2 expanded a-packed lane-local origin JU=1", k_expanded<1, 1>);
    run("V2 expanded a-packed lane-local origin JU=2", k_expanded<2, 1>);
    run("V3 expanded b-packed uniform origin JU=1", k_expanded_bp<1, 0>);
    run("V3 expanded b-packed uniform origin JU=2", k_expanded_bp<2, 0>);
    run("V3L expanded b-packed lane-local origin JU=1", k_expanded_bp<1, 1>);
    run("V3L expanded b-packed lane-local origin JU=2", k_expanded_bp<2, 1>);
    run("V4 expanded a-packed WITHOUT the |a|^2 add (timing only) JU=1", k_expanded_noadd<1>);
    run("V4 expanded a-packed WITHOUT the |a|^2 add (timing only) JU=2", k_expanded_noadd<2>);
    return 0;
}
