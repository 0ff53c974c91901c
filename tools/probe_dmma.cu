// DMMA probe: FP64 tensor-core throughput on this GPU (mma.sync.aligned.m8n8k4 / m16n8k4 / m16n8k8 .f64) next to the
// vector DFMA probe. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_dmma tools/probe_dmma.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k884(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3 + i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k1688(double* out, int iters, double a0, double b0) {
    double acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[4] = {a0, a0 * 1.1, a0 * 1.2, a0 * 1.3}, b[2] = {b0, b0 * 0.9};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k16816(double* out, int iters, double a0, double b0) {
    double acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = a0 * (1 + 0.1 * j);
    for (int j = 0; j < 4; ++j) b[j] = b0 * (1 - 0.1 * j);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) kdfma(double* out, int iters, double b, double c) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    double s = 0;
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3;
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4000;
    double* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    const double warps = (double)blocks * threads / 32;
    double t;
    t = timeit([&] { kdfma<<<blocks, threads>>>(buf, iters, 0.999999, 1e-7); });
    printf("DFMA vector            : %7.2f TFLOP/s\n", 2.0 * 64 * iters * blocks * threads / t / 1e12);
    t = timeit([&] { k884<4><<<blocks, threads>>>(buf, iters, 0.999, 1.0001); });
    printf("DMMA m8n8k4   4 acc    : %7.2f TFLOP/s\n", 2.0 * 8 * 8 * 4 * 4 * iters * warps / t / 1e12);
    t = timeit([&] { k884<8><<<blocks, threads>>>(buf, iters, 0.999, 1.0001); });
    printf("DMMA m8n8k4   8 acc    : %7.2f TFLOP/s\n", 2.0 * 8 * 8 * 4 * 8 * iters * warps / t / 1e12);
    t = timeit([&] { k1688<4><<<blocks, threads>>>(buf, iters, 0.999, 1.0001); });
    printf("DMMA m16n8k8  4 acc    : %7.2f TFLOP/s\n", 2.0 * 16 * 8 * 8 * 4 * iters * warps / t / 1e12);
    t = timeit([&] { k16816<4><<<blocks, threads>>>(buf, iters, 0.999, 1.0001); });
    printf("DMMA m16n8k16 4 acc    : %7.2f TFLOP/s\n", 2.0 * 16 * 8 * 16 * 4 * iters * warps / t / 1e12);
    t = timeit([&] { k16816<8><<<blocks, threads>>>(buf, iters, 0.999, 1.0001); });
    printf("DMMA m16n8k16 8 acc    : %7.2f TFLOP/s\n", 2.0 * 16 * 8 * 16 * 8 * iters * warps / t / 1e12);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
