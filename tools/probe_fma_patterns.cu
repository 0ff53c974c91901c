// probe_fma_patterns.cu — operand-pattern micro-benchmarks for the FIR inner loop (sm_100a).
// Pattern = the real kernel's: NACC independent accumulators, each FMA reads a distinct sample register,
// a coefficient shared by G consecutive FMAs, and its accumulator.  Build & run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe tools/probe_fma_patterns.cu && ./probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int NACC, int G>
__global__ void __launch_bounds__(128) k_scalar(float* out, int iters, const float* in) {
    float x[NACC], acc[NACC], cf[NACC / G];
    for (int i = 0; i < NACC; ++i) { x[i] = in[threadIdx.x + i]; acc[i] = 0.f; }
    for (int i = 0; i < NACC / G; ++i) cf[i] = in[i + 64];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x[(i + rep) % NACC], cf[i / G], acc[i]);
    }
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int G>
__global__ void __launch_bounds__(128) k_packed(u64* out, int iters, const u64* in) {
    u64 x[NACC], acc[NACC], cf[NACC / G];
    for (int i = 0; i < NACC; ++i) { x[i] = in[threadIdx.x + i]; acc[i] = 0; }
    for (int i = 0; i < NACC / G; ++i) cf[i] = in[i + 64];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = ffma2(x[(i + rep) % NACC], cf[i / G], acc[i]);
    }
    u64 s = 0; for (int i = 0; i < NACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K, class B> double run(K k, B* out, const B* in, int blocks, double fma_per_thread_iter, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<blocks, 128>>>(out, iters / 4, in); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<blocks, 128>>>(out, iters, in); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return 2.0 * fma_per_thread_iter * iters * blocks * 128.0 / (ms * 1e-3) / 1e12;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
    for (int wps = 4; wps <= 8; wps += 4) {  // warps per SM sub-partition
        int blocks = sms * wps;
        printf("warps/SMSP=%d\n", wps);
        printf("  scalar FFMA  12 acc, coef shared by 12: %6.1f TFLOP/s\n", run(k_scalar<12, 12>, (float*)out, (const float*)in, blocks, 48, 4096));
        printf("  scalar FFMA  12 acc, coef shared by  6: %6.1f TFLOP/s\n", run(k_scalar<12, 6>, (float*)out, (const float*)in, blocks, 48, 4096));
        printf("  scalar FFMA  12 acc, coef shared by  1: %6.1f TFLOP/s\n", run(k_scalar<12, 1>, (float*)out, (const float*)in, blocks, 48, 4096));
        printf("  packed FFMA2 12 acc, coef shared by 12: %6.1f TFLOP/s\n", run(k_packed<12, 12>, (u64*)out, (const u64*)in, blocks, 96, 4096));
        printf("  packed FFMA2 12 acc, coef shared by  6: %6.1f TFLOP/s\n", run(k_packed<12, 6>, (u64*)out, (const u64*)in, blocks, 96, 4096));
        printf("  packed FFMA2 12 acc, coef shared by  1: %6.1f TFLOP/s\n", run(k_packed<12, 1>, (u64*)out, (const u64*)in, blocks, 96, 4096));
        printf("  packed FFMA2 24 acc, coef shared by 12: %6.1f TFLOP/s\n", run(k_packed<24, 12>, (u64*)out, (const u64*)in, blocks, 192, 2048));
    }
    return 0;
}
