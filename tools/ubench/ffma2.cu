// Micro-benchmark: does fma.rn.f32x2 (sm_100) halve the issue slots of paired FP32 FMAs?
// 16 one-warp CTAs per SM (the k_align2 shape), 8 independent FMA chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(32, 16) k_scalar(float *out, int iters, float a, float b) {
    float r[8];
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
    }
    float s = 0; for (int i = 0; i < 8; i++) s += r[i];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(32, 16) k_packed(float *out, int iters, float a, float b) {
    unsigned long long r[4], A, B;
    for (int i = 0; i < 4; i++) { float lo = threadIdx.x + 2 * i, hi = threadIdx.x + 2 * i + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(lo), "f"(hi)); }
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(A), "l"(B));
    }
    float s = 0;
    for (int i = 0; i < 4; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    out[blockIdx.x * 32 + threadIdx.x] = s;
}
// mixed: per 2 FMAs (or one packed) also 2 ALU-pipe ops (setp+selp-ish via max) -- the logadd mix
__global__ void __launch_bounds__(32, 16) k_scalar_mix(float *out, int iters, float a, float b) {
    float r[8], m[8];
    for (int i = 0; i < 8; i++) { r[i] = threadIdx.x + i; m[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(r[(i + 3) & 7]));
            }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += r[i] + m[i];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(32, 16) k_packed_mix(float *out, int iters, float a, float b) {
    unsigned long long r[4], A, B;
    float m[8];
    for (int i = 0; i < 8; i++) m[i] = i;
    for (int i = 0; i < 4; i++) { float lo = threadIdx.x + 2 * i, hi = threadIdx.x + 2 * i + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(lo), "f"(hi)); }
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(A), "l"(B));
                float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[(i + 1) & 3]));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(m[2 * i]) : "f"(lo));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(m[2 * i + 1]) : "f"(hi));
            }
    }
    float s = 0;
    for (int i = 0; i < 4; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    for (int i = 0; i < 8; i++) s += m[i];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    const int blocks = 148 * 16, iters = 20000;
    float *out; cudaMalloc(&out, blocks * 32 * sizeof(float));
    const double fmas = (double) blocks * 32 * iters * 64;
    float t;
    t = timeit([&] { k_scalar<<<blocks, 32>>>(out, iters, 0.999f, 0.001f); });      printf("scalar      %.3f ms  %.2f TFMA/s\n", t, fmas / t * 1e-9);
    t = timeit([&] { k_packed<<<blocks, 32>>>(out, iters, 0.999f, 0.001f); });      printf("packed      %.3f ms  %.2f TFMA/s\n", t, fmas / t * 1e-9);
    t = timeit([&] { k_scalar_mix<<<blocks, 32>>>(out, iters, 0.999f, 0.001f); });  printf("scalar+alu  %.3f ms  %.2f TFMA/s\n", t, fmas / t * 1e-9);
    t = timeit([&] { k_packed_mix<<<blocks, 32>>>(out, iters, 0.999f, 0.001f); });  printf("packed+alu  %.3f ms  %.2f TFMA/s\n", t, fmas / t * 1e-9);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
