// Micro-benchmark: issue rate of fma.rn.f32x2 vs fma.rn.f32 on sm_100a (one CTA per SM, 512 threads).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.0001f, c = 0.5f;
    unsigned long long p0, p1, p2, p3, pb, pc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
                a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pb), "l"(pc));
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    float q0, q1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(p0 ^ p1 ^ p2 ^ p3));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + q0 + q1;
}
int main() {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<148, 512>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        // per SM: 16 warps x iters x 32 FFMA warp-instructions = lane-flops 32 per instr
        printf("FFMA   : %lld cycles, %.2f warp-instr/cycle/SM, %.1f fp32 FMA lanes/cycle/SM\n", h, 16.0 * iters * 32 / h, 16.0 * iters * 32 * 32 / h);
        k<1><<<148, 512>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA2  : %lld cycles, %.2f warp-instr/cycle/SM, %.1f fp32 FMA lanes/cycle/SM\n", h, 16.0 * iters * 16 / h, 16.0 * iters * 16 * 64 / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
