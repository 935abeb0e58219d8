// Developer micro-benchmark: issue rate of packed FP32 (FFMA2 / FMUL2 / FADD2, sm_100) against scalar FFMA, alone and mixed
// with ALU-pipe work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned iters, float *sink) {
    float s[8]; unsigned long long v[8];
    for (int i = 0; i < 8; i++) { s[i] = threadIdx.x * 1e-3f + i; v[i] = ((unsigned long long)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 0.5f); }
    const float m = 1.000001f, c = 1e-7f;
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m), c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (unsigned it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __fmaf_rn(s[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = ffma2(v[i], m2, c2);
        }
    }
    float acc = 0; unsigned long long a2 = 0;
    for (int i = 0; i < 8; i++) { acc += s[i]; a2 ^= v[i]; }
    if (acc == 12345.6f || a2 == 42ull) sink[0] = acc;
}
int main() {
    float *sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned iters = 1 << 16, blocks = 148 * 8;
    for (int mode = 0; mode < 2; mode++) {
        float ms;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(iters, sink); else k<1><<<blocks, 256>>>(iters, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double inst = (double)blocks * 256 * iters * 8;
        printf("%s: %.3f ms, %.2f T lane-instr/s, %.2f TFLOP/s\n", mode == 0 ? "FFMA " : "FFMA2", ms, inst / ms / 1e9, inst * (mode ? 4 : 2) / ms / 1e9);
    }
    return 0;
}
