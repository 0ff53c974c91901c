// probe_fma_patterns2.cu — which FFMA2 operand orderings reach the FMA-pipe peak when a coefficient pair is
// shared by several accumulators (the FIR inner loop's pattern)?  sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 ffma2v(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MODE 0: blocks of 6 FFMA2 per coefficient (kernel today). MODE 1: round-robin over the NC coefficients.
// MODE 2: like 0 but volatile asm (program order kept by the front end). MODE 3: like 1, volatile.
template <int NC, int MODE>
__global__ void __launch_bounds__(128) k(u64* out, int iters, const u64* in) {
    constexpr int NACC = 12;
    u64 x[NACC + 4], acc[NACC], cf[NC];
    for (int i = 0; i < NACC + 4; ++i) x[i] = in[threadIdx.x + i];
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    for (int i = 0; i < NC; ++i) cf[i] = in[i + 64];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                int i, ci;
                if (MODE == 0 || MODE == 2) { i = j; ci = j / (NACC / NC); }          // coefficient-major
                else { ci = j % NC; i = (j % NC) * (NACC / NC) + j / NC; }              // round-robin
                if (MODE >= 2) acc[i] = ffma2v(x[i + rep], cf[ci], acc[i]);
                else acc[i] = ffma2(x[i + rep], cf[ci], acc[i]);
            }
        }
    }
    u64 s = 0; for (int i = 0; i < NACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> double run(K kern, u64* out, const u64* in, int blocks, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<blocks, 128>>>(out, iters / 4, in); cudaDeviceSynchronize();
    cudaEventRecord(a); kern<<<blocks, 128>>>(out, iters, in); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return 2.0 * 96.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
    int blocks = sms * 4;
#define RUN(NC, MODE, label) printf("  NC=%2d %-34s %6.1f TFLOP/s\n", NC, label, run(k<NC, MODE>, (u64*)out, (const u64*)in, blocks, 4096));
    RUN(1, 0, "one coefficient for all 12");
    RUN(2, 0, "coef-major (6 per coef)");
    RUN(2, 1, "round-robin over 2 coefs");
    RUN(2, 2, "coef-major, volatile");
    RUN(2, 3, "round-robin, volatile");
    RUN(4, 0, "coef-major (3 per coef)");
    RUN(4, 1, "round-robin over 4 coefs");
    RUN(4, 3, "round-robin over 4, volatile");
    RUN(6, 1, "round-robin over 6 coefs");
    RUN(12, 0, "all distinct");
    return 0;
}
