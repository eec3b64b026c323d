// Stripped-down inner loop of k_sweep to find its structural throughput limit on sm_100a.
// body(j): load (bx0,by0,bx1,by1) from shared memory; for k<H: 4 FADD2, 2 FMUL2, 2 FFMA2 [+ mins].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// MODE 0: full (row+col mins, min3)  1: FMA-pipe only (xor-fold results)  2: row mins only  3: col mins only
// 4: full but 2-input mins          5: full, no REDUX
template <int H, int MODE, int JU>
__global__ void __launch_bounds__(256, 2) kern(float* out, int m_pairs, int reps, float s) {
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
    uint64_t fold = 0;
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
                if (MODE == 1) { fold ^= d0 ^ d1; continue; }
                float d00, d10, d01, d11;
                upk(d0, d00, d10); upk(d1, d01, d11);
                if (MODE == 0 || MODE == 2 || MODE == 5) {
                    row[2 * k] = min3(row[2 * k], d00, d01);
                    row[2 * k + 1] = min3(row[2 * k + 1], d10, d11);
                }
                if (MODE == 0 || MODE == 3 || MODE == 5) {
                    c0 = min3(c0, d00, d10);
                    c1 = min3(c1, d01, d11);
                }
                if (MODE == 4) {
                    row[2 * k] = fminf(fminf(row[2 * k], d00), d01);
                    row[2 * k + 1] = fminf(fminf(row[2 * k + 1], d10), d11);
                    c0 = fminf(fminf(c0, d00), d10);
                    c1 = fminf(fminf(c1, d01), d11);
                }
                if (MODE == 2) fold ^= (uint64_t)__float_as_uint(d10) ;
            }
            if (MODE == 0 || MODE == 3 || MODE == 4) {
                unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(c0));
                unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(c1));
                colmax = max(colmax, max(r0, r1));
            } else if (MODE == 5) {
                colmax = max(colmax, max(__float_as_uint(c0), __float_as_uint(c1)));
            }
        }
    }
    float acc = __uint_as_float(colmax) + (float)(fold & 0xffff);
    for (int k = 0; k < 2 * H; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// b-packed variant: one packed instruction = one test point against TWO reference points; the reference pair
// (bx0,bx1)/(by0,by1) is the same operand for every test point of the lane (operand-reuse friendly).
template <int TA, int JU>
__global__ void __launch_bounds__(256, 2) kern_bp(float* out, int m_pairs, int reps, float s) {
    __shared__ float4 sB[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sB[i] = make_float4(s * i, s * i + 1, s * i + 2, s * i + 3);
    __syncthreads();
    float ax[TA], ay[TA], row[TA];
    for (int k = 0; k < TA; ++k) { ax[k] = threadIdx.x * 0.01f + k; ay[k] = threadIdx.x * 0.03f + k; row[k] = 1e30f; }
    unsigned colmax = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll JU
        for (int j = 0; j < m_pairs; ++j) {
            const float4 B = sB[j];  // (bx0, bx1, by0, by1)
            const uint64_t BX = pk(B.x, B.y), BY = pk(B.z, B.w);
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int k = 0; k < TA; k += 2) {
                const uint64_t dxa = sub2(pk(ax[k], ax[k]), BX), dya = sub2(pk(ay[k], ay[k]), BY);
                const uint64_t dxb = sub2(pk(ax[k + 1], ax[k + 1]), BX), dyb = sub2(pk(ay[k + 1], ay[k + 1]), BY);
                const uint64_t da = fma2(dxa, dxa, mul2(dya, dya));  // (|a-b0|^2, |a-b1|^2)
                const uint64_t db = fma2(dxb, dxb, mul2(dyb, dyb));
                float a0, a1, b0, b1;
                upk(da, a0, a1); upk(db, b0, b1);
                row[k] = min3(row[k], a0, a1);
                row[k + 1] = min3(row[k + 1], b0, b1);
                c0 = min3(c0, a0, b0);
                c1 = min3(c1, a1, b1);
            }
            unsigned r0 = __reduce_min_sync(0xffffffffu, __float_as_uint(c0));
            unsigned r1 = __reduce_min_sync(0xffffffffu, __float_as_uint(c1));
            colmax = max(colmax, max(r0, r1));
        }
    }
    float acc = __uint_as_float(colmax);
    for (int k = 0; k < TA; ++k) acc += row[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int TA, int JU>
void run_bp(const char* name) {
    float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
    const int m_pairs = 256, reps = 200;
    kern_bp<TA, JU><<<148 * 2, 256>>>(out, m_pairs, 2, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern_bp<TA, JU><<<148 * 2, 256>>>(out, m_pairs, reps, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double kiters = (double)reps * m_pairs * (TA / 2) * 4;
    double cyc = ms * 1e-3 * 1.965e9 / kiters;
    printf("TA=%d %-42s JU=%d %8.3f ms  %6.2f SMSP-cycles per 2x2 pair block\n", TA, name, JU, ms, cyc);
    cudaFree(out);
}

template <int H, int MODE, int JU>
void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
    const int m_pairs = 256, reps = 200;
    kern<H, MODE, JU><<<148 * 2, 256>>>(out, m_pairs, 2, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<H, MODE, JU><<<148 * 2, 256>>>(out, m_pairs, reps, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // SMSP cycles per k-iteration (8 FMA-pipe packed instrs): 4 warps per SMSP
    double kiters = (double)reps * m_pairs * H * 4;  // per SMSP
    double cyc = ms * 1e-3 * 1.965e9 / kiters;
    printf("H=%d %-44s JU=%d %8.3f ms  %6.2f SMSP-cycles per k-iteration (FMA-pipe floor 16.0) => FMA pipe %.1f%%\n", H, name, JU, ms, cyc, 1600.0 / cyc);
    cudaFree(out);
}

int main() {
    run_bp<16, 1>("b-packed full");
    run_bp<16, 2>("b-packed full");
    run_bp<18, 2>("b-packed full");
    run<8, 0, 2>("a-packed full (shipped form)");
    run<9, 0, 2>("a-packed full (shipped form)");
    return 0;
}
int main_old() {
    run<8, 1, 1>("FMA-pipe only");
    run<8, 2, 1>("FMA + row mins (2 FMNMX3)");
    run<8, 3, 1>("FMA + col mins (2 FMNMX3) + REDUX");
    run<8, 5, 1>("FMA + 4 FMNMX3, no REDUX");
    run<8, 0, 1>("full: 4 FMNMX3 + REDUX");
    run<8, 0, 2>("full: 4 FMNMX3 + REDUX");
    run<8, 4, 1>("full with 2-input FMNMX (8)");
    run<9, 0, 1>("full: 4 FMNMX3 + REDUX");
    run<9, 0, 2>("full: 4 FMNMX3 + REDUX");
    run<6, 0, 1>("full: 4 FMNMX3 + REDUX");
    run<4, 0, 1>("full: 4 FMNMX3 + REDUX");
    return 0;
}
