// Second round of sm_100a issue-model microbenchmarks: how packed FP32 instructions and FMNMX3 share the SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// MODE 0: 8x FFMA2(pair,pair,acc)                         1: 8x FFMA2(pair, scalar, acc)      2: 8x FFMA2(acc, scalar, scalar)
// MODE 3: 8x FFMA2(pair,pair,acc) + 4x FMNMX3             4: 8x FFMA2 + 8x FMNMX3             5: 4x FMNMX3 only     6: 8x FMNMX3 only
// MODE 7: 8x FFMA2 + 4x IADD3 (integer ALU)               8: 8x FFMA (scalar) + 4x FMNMX3     9: 16x FFMA scalar + 4 FMNMX3
template <int MODE>
__global__ void __launch_bounds__(256, 2) kern(float* out, int reps, float s) {
    uint64_t acc[8], x[8], y[8];
    float m[8], f[16];
    int ii[4] = {1, 2, 3, 4};
    const float sc = s * 1.0001f, sd = s * 0.5f;
    for (int k = 0; k < 8; ++k) {
        acc[k] = pk(threadIdx.x * 0.001f + k, k); x[k] = pk(1.0001f + k * 1e-4f, 0.9999f); y[k] = pk(1e-3f * k, 2e-3f);
        m[k] = 1e30f; f[2 * k] = k; f[2 * k + 1] = k + 0.5f;
    }
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 7)
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k]) : "l"(x[k]), "l"(y[k]));
            if (MODE == 1) { uint64_t sp = pk(sc, sc); asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k]) : "l"(x[k]), "l"(sp)); }
            if (MODE == 2) { uint64_t sp = pk(sc, sc), sq = pk(sd, sd); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(sp), "l"(sq)); }
            if (MODE == 8 || MODE == 9) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[2 * k]) : "f"(sc), "f"(sd));
                if (MODE == 9) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[2 * k + 1]) : "f"(sd), "f"(sc));
            }
            if ((MODE == 3 || MODE == 8 || MODE == 9 || MODE == 5) && (k & 1)) {
                float lo, hi; upk(acc[k], lo, hi);
                if (MODE == 8 || MODE == 9 || MODE == 5) { lo = f[2 * k]; hi = f[2 * k - 1]; }
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(lo), "f"(hi));
            }
            if (MODE == 4 || MODE == 6) {
                float lo, hi; upk(acc[k], lo, hi);
                if (MODE == 6) { lo = f[2 * k]; hi = f[2 * k + 1]; }
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(lo), "f"(hi));
            }
            if (MODE == 7 && (k & 1)) asm volatile("add.s32 %0, %0, %1;" : "+r"(ii[k >> 1]) : "r"(k + r));
        }
    }
    float a = 0;
    for (int k = 0; k < 8; ++k) { float lo, hi; upk(acc[k], lo, hi); a += lo + hi + m[k] + f[2 * k] + f[2 * k + 1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + ii[0] + ii[1] + ii[2] + ii[3];
}
template <int MODE>
void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
    const int reps = 20000;
    kern<MODE><<<148 * 2, 256>>>(out, 100, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<MODE><<<148 * 2, 256>>>(out, reps, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * 1.965e9 / ((double)reps * 4);   // SMSP-cycles per loop body (8 slots) per warp, 4 warps/SMSP
    printf("%-52s %8.3f ms  %6.2f SMSP-cycles per body\n", name, ms, cyc);
    cudaFree(out);
}
int main() {
    run<0>("8 FFMA2 (pair,pair,acc)");
    run<1>("8 FFMA2 (pair,scalar,acc)");
    run<2>("8 FFMA2 (acc,scalar,scalar)");
    run<5>("4 FMNMX3");
    run<6>("8 FMNMX3");
    run<3>("8 FFMA2 + 4 FMNMX3");
    run<4>("8 FFMA2 + 8 FMNMX3");
    run<7>("8 FFMA2 + 4 IADD");
    run<8>("8 FFMA + 4 FMNMX3");
    run<9>("16 FFMA + 4 FMNMX3");
    return 0;
}
