// probe_fma_const.cu — can the FIR loop's coefficient operand come from the constant bank (kernel parameters) or a uniform
// register instead of a vector register? An FFMA2 with three register-file operands runs at ~2/3 rate (probe_fma_patterns3);
// with the coefficient elsewhere only two are read. Prints TFLOP/s; inspect the SASS (cuobjdump -sass) for `c[0x0]` / `UR` operands.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
struct Coefs { u64 c[960]; };   // 7.5 KB of kernel parameters (CUDA 12.1+: up to 32 KB)

// MODE 0: coefficients from shared memory into registers (the kernel today). MODE 1: from the parameter (constant) bank,
// dynamic uniform index. MODE 2: scalar FFMA with constant-bank coefficients.
template <int MODE>
__global__ void __launch_bounds__(128) k(u64* out, int iters, const u64* in, const __grid_constant__ Coefs P) {
    constexpr int NACC = 12;
    __shared__ u64 cs[960];
    for (int i = threadIdx.x; i < 960; i += 128) cs[i] = in[i];
    __syncthreads();
    u64 x[NACC + 2], acc[NACC];
    for (int i = 0; i < NACC + 2; ++i) x[i] = in[threadIdx.x + i];
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    for (int it = 0; it < iters; ++it) {
        const int base = (it & 127) * 4;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            u64 c0, c1;
            if (MODE == 0) { c0 = cs[base + q]; c1 = cs[base + 2 + q]; }
            else { c0 = P.c[base + q]; c1 = P.c[base + 2 + q]; }
#pragma unroll
            for (int r = 0; r < NACC; ++r) acc[r] = ffma2(x[(r + q) % (NACC + 2)], (r & 1) ? c1 : c0, acc[r]);
        }
    }
    u64 s = 0; for (int i = 0; i < NACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
__global__ void __launch_bounds__(128) ks(float* out, int iters, const float* in, const __grid_constant__ Coefs P) {
    constexpr int NACC = 24;
    float x[NACC + 4], acc[NACC];
    for (int i = 0; i < NACC + 4; ++i) x[i] = in[threadIdx.x + i];
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    const float* pc = reinterpret_cast<const float*>(P.c);
    for (int it = 0; it < iters; ++it) {
        const int base = (it & 127) * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float c0 = pc[base + q], c1 = pc[base + 4 + q];
#pragma unroll
            for (int r = 0; r < NACC; ++r) acc[r] = fmaf(x[(r + q) % (NACC + 4)], (r & 1) ? c1 : c0, acc[r]);
        }
    }
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
    Coefs P{}; 
    int blocks = sms * 4, iters = 8192;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
#define RUN(KERN, FL, label) { KERN<<<blocks, 128>>>((decltype(KERN == nullptr, (u64*)0))out, iters / 4, (const u64*)in, P); }
    k<0><<<blocks, 128>>>((u64*)out, iters / 4, (const u64*)in, P); cudaDeviceSynchronize();
    cudaEventRecord(a); k<0><<<blocks, 128>>>((u64*)out, iters, (const u64*)in, P); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("  FFMA2, coefficients via shared memory -> registers     %6.1f TFLOP/s\n", 96.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12);
    k<1><<<blocks, 128>>>((u64*)out, iters / 4, (const u64*)in, P); cudaDeviceSynchronize();
    cudaEventRecord(a); k<1><<<blocks, 128>>>((u64*)out, iters, (const u64*)in, P); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("  FFMA2, coefficients from the parameter (constant) bank %6.1f TFLOP/s\n", 96.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12);
    ks<0><<<blocks, 128>>>((float*)out, iters / 4, (const float*)in, P); cudaDeviceSynchronize();
    cudaEventRecord(a); ks<0><<<blocks, 128>>>((float*)out, iters, (const float*)in, P); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("  scalar FFMA, coefficients from the constant bank        %6.1f TFLOP/s\n", 96.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
