// probe_fma_patterns3.cu — the C4 kernel's exact FFMA2 pattern (12 accumulators, per 4-tap step two tap pairs q, window
// parity sh = r & 1 selects the filter copy: 4 coefficient registers c[sh][q] feed 24 FFMA2s) and variants that put more
// distinct coefficient registers into rotation. ptxas schedules freely, so each variant also prints how many FFMA2s in a row
// share their coefficient register in the SASS it got (cuobjdump), not just the source order.   sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MODE 0: kernel today (q outer, r inner; c[r&1][q]).  MODE 1: duplicated registers for r >= 6 (8 coefficient registers).
// MODE 2: three copies (r / 4).  MODE 3: every r its own register (24 registers: upper bound).
template <int MODE>
__global__ void __launch_bounds__(128) k(u64* out, int iters, const u64* in) {
    constexpr int NACC = 12;
    u64 x[NACC + 2], acc[NACC], cf[24];
    for (int i = 0; i < NACC + 2; ++i) x[i] = in[threadIdx.x + i];
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    for (int i = 0; i < 24; ++i) cf[i] = in[i + 64 + (threadIdx.x & 1)];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (MODE >= 4) {
                // MODE 4: c[sh][q] register-major (6 FFMA2s in a row per register).  MODE 5: 3 in a row.
                // MODE 6: round-robin over the four registers: position r uses tap pair q = ((r >> 1) + half) & 1.
                if (MODE == 4) {
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int sh = 0; sh < 2; ++sh)
#pragma unroll
                            for (int i = 0; i < 6; ++i) {
                                const int r = 2 * i + sh;
                                acc[r] = ffma2(x[(r + q + 2 * rep) % (NACC + 2)], cf[sh * 2 + q], acc[r]);
                            }
                } else if (MODE == 5) {
#pragma unroll
                    for (int rb = 0; rb < 12; rb += 6)
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int sh = 0; sh < 2; ++sh)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const int r = rb + 2 * i + sh;
                                    acc[r] = ffma2(x[(r + q + 2 * rep) % (NACC + 2)], cf[sh * 2 + q], acc[r]);
                                }
                } else {
#pragma unroll
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int r = 0; r < NACC; ++r) {
                            const int q = ((r >> 1) + half) & 1;
                            acc[r] = ffma2(x[(r + q + 2 * rep) % (NACC + 2)], cf[(r & 1) * 2 + q], acc[r]);
                        }
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int r = 0; r < NACC; ++r) {
                    int ci = (r & 1) * 2 + q;
                    if (MODE == 1) ci += (r / 6) * 4;
                    if (MODE == 2) ci += (r / 4) * 4;
                    if (MODE == 3) ci = r * 2 + q;
                    acc[r] = ffma2(x[(r + q + 2 * rep) % (NACC + 2)], cf[ci], acc[r]);  // circular window, compile-time index
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
    return 2.0 * 192.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
    int blocks = sms * 4;
#define RUN(MODE, label) printf("  %-52s %6.1f TFLOP/s\n", label, run(k<MODE>, (u64*)out, (const u64*)in, blocks, 4096));
    RUN(0, "4 coefficient registers c[sh][q] (kernel today)");
    RUN(1, "8 registers: duplicates for positions 6..11");
    RUN(2, "12 registers: a copy per 4 positions");
    RUN(3, "24 registers: one per FFMA2 (upper bound)");
    RUN(4, "4 registers, register-major (6 in a row)");
    RUN(5, "4 registers, 3 in a row");
    RUN(6, "4 registers, round-robin A B C D");
    return 0;
}
