// probe_dfma_const.cu — float64 version of probe_fma_const.cu: does DFMA take its coefficient from a uniform register
// (constant bank / kernel parameters), and what does the x2 FIR core's operand pattern gain from it? R = 6 positions x 2 phases,
// coefficient pair per tap shared by all lanes.
#include <cstdio>
#include <cuda_runtime.h>
struct Coefs { double c[2][256]; };
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int iters, const double* in, const __grid_constant__ Coefs P) {
    constexpr int R = 6;
    __shared__ double cs[2][256];
    for (int i = threadIdx.x; i < 512; i += 128) (&cs[0][0])[i] = in[i];
    __syncthreads();
    double x[R + 2], acc[R][2];
    for (int i = 0; i < R + 2; ++i) x[i] = in[threadIdx.x + i];
    for (int i = 0; i < R; ++i) acc[i][0] = acc[i][1] = 0;
    const double* c0 = MODE ? P.c[0] : cs[0];
    const double* c1 = MODE ? P.c[1] : cs[1];
    for (int it = 0; it < iters; ++it) {
        const double* p0 = c0 + (it & 63) * 2;
        const double* p1 = c1 + (it & 63) * 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double a = p0[i], b = p1[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r][0] = fma(x[(r + i) % (R + 2)], a, acc[r][0]);
                acc[r][1] = fma(x[(r + i) % (R + 2)], b, acc[r][1]);
            }
        }
    }
    double s = 0; for (int i = 0; i < R; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
    Coefs P{};
    int blocks = sms * 8, iters = 8192;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
    k<0><<<blocks, 128>>>((double*)out, iters / 4, (const double*)in, P); cudaDeviceSynchronize();
    cudaEventRecord(a); k<0><<<blocks, 128>>>((double*)out, iters, (const double*)in, P); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("  DFMA, coefficients via shared memory -> registers     %6.1f TFLOP/s\n", 2.0 * 24.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12);
    k<1><<<blocks, 128>>>((double*)out, iters / 4, (const double*)in, P); cudaDeviceSynchronize();
    cudaEventRecord(a); k<1><<<blocks, 128>>>((double*)out, iters, (const double*)in, P); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("  DFMA, coefficients from the parameter (constant) bank %6.1f TFLOP/s\n", 2.0 * 24.0 * iters * blocks * 128.0 / (ms * 1e-3) / 1e12);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
