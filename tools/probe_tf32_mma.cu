// mma.sync TF32 / BF16 probe (legacy tensor path, SASS HMMA) next to FFMA: informs whether a split-TF32 (3 x TF32 ~ fp32)
// FIR could beat the FFMA2 kernel. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_tf32 tools/probe_tf32_mma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void __launch_bounds__(256) ktf32(float* out, int iters) {
    float acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-3f + i + j;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f810000u, 0x3f820000u, 0x3f830000u}, b[2] = {0x3f000000u, 0x3f010000u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_tf32(acc[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4000;
    float* buf; cudaMalloc(&buf, (size_t)blocks * threads * 4);
    const double warps = (double)blocks * threads / 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int nacc : {4, 8}) {
        auto launch = [&] { if (nacc == 4) ktf32<4><<<blocks, threads>>>(buf, iters); else ktf32<8><<<blocks, threads>>>(buf, iters); };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mma.sync m16n8k8 tf32, %d acc: %8.1f TFLOP/s\n", nacc, 2.0 * 16 * 8 * 8 * nacc * iters * warps / (ms * 1e-3) / 1e12);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
