// Instruction-throughput microbenchmarks for the sweep kernel's instruction mix on sm_100a.
// Each test runs REP x (unrolled body) per thread with 16 warps/SM-resident CTAs; reports cycles per
// body per SMSP (4 warps per SMSP resident => divide accordingly). Build: nvcc -arch=sm_100a -O3 pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define PK(lo, hi, out) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(out) : "f"(lo), "f"(hi))

template <int MODE>
__global__ void __launch_bounds__(256, 2) kern(float* out, int reps, float s, long long* cyc) {
    // 8 independent "pairs-of-points" register sets
    uint64_t ax[8], ay[8], d[8];
    float row[16], c0 = 1e30f, c1 = 1e30f;
    float bx = s * 1.5f, by = s * 2.5f, bx1 = s * 3.5f, by1 = s * 0.5f;
    for (int k = 0; k < 8; ++k) {
        float a = threadIdx.x * 0.001f + k, b = a * 0.5f + s;
        PK(a, b, ax[k]); PK(b, a, ay[k]); d[k] = 0;
        row[2 * k] = 1e30f; row[2 * k + 1] = 1e30f;
    }
    uint64_t bxx, byy, bxx1, byy1;
    PK(bx, bx, bxx); PK(by, by, byy); PK(bx1, bx1, bxx1); PK(by1, by1, byy1);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) {  // FFMA2 only, 3 distinct regs: d = ax*ay + d
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[k]) : "l"(ax[k]), "l"(ay[k]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[k]) : "l"(ay[k]), "l"(ax[k]));
            } else if (MODE == 1) {  // FADD2 with scalar-broadcast b
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d[k]) : "l"(ax[k]), "l"(bxx));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(ax[k]) : "l"(d[k]), "l"(byy));
            } else if (MODE == 2) {  // FMNMX3 only
                float lo, hi;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(ax[k]));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(row[2 * k]) : "f"(lo), "f"(hi));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(row[2 * k + 1]) : "f"(hi), "f"(lo));
            } else if (MODE == 3 || MODE == 4 || MODE == 5) {
                // the sweep's k-iteration: 4 FADD2, 2 FMUL2, 2 FFMA2 (+ 4 FMNMX3 in MODE 3, + 8 FMNMX in MODE 5)
                uint64_t dx0, dy0, dx1, dy1, t0_, t1_, e0, e1;
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx0) : "l"(ax[k]), "l"(bxx));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy0) : "l"(ay[k]), "l"(byy));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx1) : "l"(ax[k]), "l"(bxx1));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy1) : "l"(ay[k]), "l"(byy1));
                asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(t0_) : "l"(dy0));
                asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(t1_) : "l"(dy1));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(e0) : "l"(dx0), "l"(t0_));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(e1) : "l"(dx1), "l"(t1_));
                float d00, d10, d01, d11;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d00), "=f"(d10) : "l"(e0));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d01), "=f"(d11) : "l"(e1));
                if (MODE == 3) {
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(row[2 * k]) : "f"(d00), "f"(d01));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(row[2 * k + 1]) : "f"(d10), "f"(d11));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(c0) : "f"(d00), "f"(d10));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(c1) : "f"(d01), "f"(d11));
                } else if (MODE == 5) {
                    row[2 * k] = fminf(fminf(row[2 * k], d00), d01);
                    row[2 * k + 1] = fminf(fminf(row[2 * k + 1], d10), d11);
                    c0 = fminf(fminf(c0, d00), d10);
                    c1 = fminf(fminf(c1, d01), d11);
                } else {
                    asm volatile("" ::"f"(d00), "f"(d10), "f"(d01), "f"(d11));
                    d[k] = e0 ^ e1;  // keep alive cheaply (LOP3 on ALU) - 2 ops
                }
            } else if (MODE == 6) {  // scalar FFMA 3 distinct regs
                float lo, hi, l2, h2;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(ax[k]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(l2), "=f"(h2) : "l"(ay[k]));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(row[2 * k]) : "f"(lo), "f"(l2));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(row[2 * k + 1]) : "f"(hi), "f"(h2));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(row[2 * k]) : "f"(h2), "f"(lo));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(row[2 * k + 1]) : "f"(l2), "f"(hi));
            } else if (MODE == 7) {  // FMUL2 self + FFMA2 (dx,dx,t): low register-read forms
                asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d[k]) : "l"(ax[k]));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(ax[k]) : "l"(ay[k]), "l"(d[k]));
            }
        }
    }
    long long t1 = clock64();
    float acc = c0 + c1;
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d[k] ^ ax[k]));
        acc += lo + hi + row[2 * k] + row[2 * k + 1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int fma_instr_per_body, int alu_instr_per_body) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 2 * 256 * 4); cudaMalloc(&cyc, 8);
    const int reps = 20000;
    kern<MODE><<<148 * 2, 256>>>(out, 100, 1.0f, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<MODE><<<148 * 2, 256>>>(out, reps, 1.0f, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: 4 warps resident (16 warps / SM); body x8 per rep per warp
    double cyc_per_body_per_warp = (double)h / reps / 8.0;
    double per_smsp = cyc_per_body_per_warp / 4.0;  // cycles of SMSP time per body (4 warps interleaved)
    printf("%-34s %8.3f ms  %7.2f cyc/body/warp  => %6.2f SMSP-cyc per body (%d FMA-pipe + %d ALU instrs): %.2f cyc per FMA-pipe instr\n",
           name, ms, cyc_per_body_per_warp, per_smsp, fma_instr_per_body, alu_instr_per_body,
           fma_instr_per_body ? per_smsp / fma_instr_per_body : 0.0);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<6>("FFMA scalar (3 regs) x4", 4, 0);
    run<0>("FFMA2 (3 distinct pairs) x2", 2, 0);
    run<1>("FADD2 (pair - scalar bcast) x2", 2, 0);
    run<7>("FMUL2 self + FFMA2(dx,dx,t)", 2, 0);
    run<2>("FMNMX3 x2", 0, 2);
    run<4>("sweep body w/o mins (8 FMA-pipe)", 8, 1);
    run<3>("sweep body (8 FMA-pipe + 4 FMNMX3)", 8, 4);
    run<5>("sweep body (8 FMA-pipe + 8 FMNMX)", 8, 8);
    return 0;
}
