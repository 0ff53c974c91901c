// kernels_misc.cu — QualityQuick cubic stage (cubic.go:33-90), carried-tail and cast kernels, the interleaved / integer-PCM
// boundary kernels (N1), the dependent-FMA probes for the roofline denominator, and the launch counter.
#include "device_common.cuh"
#include <atomic>

namespace gar {
namespace {
std::atomic<long long> g_launches{0};
}  // namespace
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

const char* tune_env(const char* name) {
    static const bool on = [] {
        const char* e = std::getenv("GAR_DEBUG_TUNING");
        return e && e[0] && e[0] != '0';
    }();
    return on ? std::getenv(name) : nullptr;
}

namespace {

// =============================================================================================
// Cubic (QualityQuick) stage: out[n] = poly(x_n) over in[idx_n-3 .. idx_n], evaluated in float64
// without contraction, exactly as cubic.go:73-85.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) cubic_kernel(const CubicCall c, const int n_tiles) {
    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    if (tile == n_tiles) {  // new tail = last 3 samples of hist(3) ++ in
        carry_row(hist, 3, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.n_in, 3);
        return;
    }
    const int n = tile * 256 + threadIdx.x;
    if (n >= c.n_out) return;
    const int i = c.idx[n];  // newest sample = in[i]; virtual index i+3 in hist(3) ++ in
    const double s2 = (double)vload(hist, 3, in, c.n_in, i + 3);
    const double s1 = (double)vload(hist, 3, in, c.n_in, i + 2);
    const double s0 = (double)vload(hist, 3, in, c.n_in, i + 1);
    const double sm1 = (double)vload(hist, 3, in, c.n_in, i);
    const double x = c.phase[n];
    const double b = __dsub_rn(__dmul_rn(0.5, __dadd_rn(s1, sm1)), s0);
    const double t = __dsub_rn(__dsub_rn(__dadd_rn(__dsub_rn(s2, s1), sm1), s0), __dmul_rn(4.0, b));
    const double a = __dmul_rn(1.0 / 6.0, t);
    const double cc = __dsub_rn(__dsub_rn(__dsub_rn(s1, s0), a), b);
    const double y = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(a, x), b), x), cc), x), s0);
    (static_cast<T*>(c.out) + row * c.out_stride)[n] = (T)y;
}

template <typename T>
__global__ void carry_kernel(const T* hist, int64_t hist_stride, int hist_len, const T* in, int64_t in_stride, int n_in,
                             T* hist_out, int64_t hist_out_stride, int drop, int new_len) {
    const int64_t row = blockIdx.x;
    carry_row(hist + row * hist_stride, hist_len, in + row * in_stride, n_in, hist_out + row * hist_out_stride, drop,
              new_len);
}

template <typename S, typename D>
__global__ void cast_kernel(const S* src, int64_t src_stride, D* dst, int64_t dst_stride, int n, int n_rows) {
    // launched as a programmatic dependent: its blocks may be resident before the launch in front has drained (and it lets the
    // launch behind it in early) — the cast launches around a streaming-size call are pure launch latency otherwise
    pdl_wait();
    if (gridDim.x * gridDim.y <= 296) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // rows are strided over gridDim.y (limited to 65535) so any row count is covered
    for (int64_t row = blockIdx.y; row < n_rows; row += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            dst[row * dst_stride + i] = (D)src[row * src_stride + i];
}


// ---- N1: interleaved / integer-PCM boundary -------------------------------------------------------
template <typename TI, typename T>
__global__ void deinterleave_kernel(const TI* __restrict__ in, int channels, int64_t n_frames, T* __restrict__ planar,
                                    int64_t stride, double inv_max, int round_f32) {
    const int64_t total = n_frames * channels;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / channels;
        const int ch = (int)(idx - i * channels);
        const double v = (double)in[idx];
        double y = inv_max != 0.0 ? __dmul_rn(v, inv_max) : v;  // main.go:444-470 deinterleaveInto: F(float64(v) * invMaxVal)
        if (round_f32) y = (double)(float)y;                    // float32 engine computing in float64: F = float32 first
        planar[ch * stride + i] = (T)y;
    }
}
template <typename T, typename TO>
__global__ void interleave_kernel(const T* __restrict__ planar, int64_t stride, int channels, int64_t n_frames,
                                  TO* __restrict__ out, double max_val, int round_f32) {
    const int64_t total = n_frames * channels;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / channels;
        const int ch = (int)(idx - i * channels);
        double v = (double)planar[ch * stride + i];
        if (round_f32) v = (double)(float)v;  // the float32 engine hands float32 samples to interleaveInto
        if (max_val != 0.0) {  // main.go:474-520 interleaveInto: clamp, scale, truncate
            v = v > 1.0 ? 1.0 : (v < -1.0 ? -1.0 : v);
            out[idx] = (TO)__double2ll_rz(__dmul_rn(v, max_val));
        } else {
            out[idx] = (TO)v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) fma_probe_kernel(T* out, int iters, T b, T cadd) {
    T a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, cadd);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed 2 x fp32 FMA (PTX fma.rn.f32x2, SASS FFMA2 — new on sm_100): operands are even/odd register pairs
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
    unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<const unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__global__ void __launch_bounds__(256) ffma2_probe_kernel(float2* out, int iters, float2 b, float2 cadd) {
    float2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = make_float2((float)(threadIdx.x + k), (float)(threadIdx.x - k));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = ffma2(a[k], b, cadd);
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s.x += a[k].x;
        s.y += a[k].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

const char* launch_cubic(const CubicCall& c, int dtype, cudaStream_t s) {
    if (c.n_streams <= 0) return "none";
    const int n_tiles = (c.n_out + 255) / 256;
    const int64_t blocks = (int64_t)(n_tiles + 1) * c.n_streams;
    count_launch();
    if (dtype == DT_F32) cubic_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(c, n_tiles);
    else cubic_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(c, n_tiles);
    return dtype == DT_F32 ? "cubic_f32" : "cubic_f64";
}

void launch_carry(const void* hist, int64_t hist_stride, int32_t hist_len, const void* in, int64_t in_stride,
                  int32_t n_in, void* hist_out, int64_t hist_out_stride, int32_t drop, int32_t new_len,
                  int32_t n_streams, int dtype, cudaStream_t s) {
    if (n_streams <= 0 || new_len <= 0) return;
    count_launch();
    if (dtype == DT_F32)
        carry_kernel<float><<<n_streams, 128, 0, s>>>((const float*)hist, hist_stride, hist_len, (const float*)in,
                                                      in_stride, n_in, (float*)hist_out, hist_out_stride, drop, new_len);
    else
        carry_kernel<double><<<n_streams, 128, 0, s>>>((const double*)hist, hist_stride, hist_len, (const double*)in,
                                                       in_stride, n_in, (double*)hist_out, hist_out_stride, drop,
                                                       new_len);
}

void launch_cast(const void* src, int64_t src_stride, int src_dtype, void* dst, int64_t dst_stride, int dst_dtype,
                 int32_t n, int32_t n_rows, cudaStream_t s) {
    if (n <= 0 || n_rows <= 0) return;
    dim3 grid((unsigned)((n + 1023) / 1024 < 4096 ? (n + 1023) / 1024 : 4096), (unsigned)(n_rows < 65535 ? n_rows : 65535));
    count_launch();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (src_dtype == DT_F32 && dst_dtype == DT_F64)
        cudaLaunchKernelEx(&cfg, cast_kernel<float, double>, (const float*)src, src_stride, (double*)dst, dst_stride, (int)n, (int)n_rows);
    else if (src_dtype == DT_F64 && dst_dtype == DT_F32)
        cudaLaunchKernelEx(&cfg, cast_kernel<double, float>, (const double*)src, src_stride, (float*)dst, dst_stride, (int)n, (int)n_rows);
    else if (src_dtype == DT_F32)
        cudaLaunchKernelEx(&cfg, cast_kernel<float, float>, (const float*)src, src_stride, (float*)dst, dst_stride, (int)n, (int)n_rows);
    else
        cudaLaunchKernelEx(&cfg, cast_kernel<double, double>, (const double*)src, src_stride, (double*)dst, dst_stride, (int)n, (int)n_rows);
}

namespace {
inline unsigned grid_for(int64_t total) {
    int64_t g = (total + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}
template <typename TI>
void deint_dispatch(const void* in, int channels, int64_t n, void* planar, int64_t stride, int dtype, double inv,
                    int round_f32, cudaStream_t s) {
    const unsigned g = grid_for(n * channels);
    if (dtype == DT_F32) deinterleave_kernel<TI, float><<<g, 256, 0, s>>>((const TI*)in, channels, n, (float*)planar, stride, inv, round_f32);
    else deinterleave_kernel<TI, double><<<g, 256, 0, s>>>((const TI*)in, channels, n, (double*)planar, stride, inv, round_f32);
}
template <typename TO>
void int_dispatch(const void* planar, int64_t stride, int dtype, int channels, int64_t n, void* out, double maxv,
                  int round_f32, cudaStream_t s) {
    const unsigned g = grid_for(n * channels);
    if (dtype == DT_F32) interleave_kernel<float, TO><<<g, 256, 0, s>>>((const float*)planar, stride, channels, n, (TO*)out, maxv, round_f32);
    else interleave_kernel<double, TO><<<g, 256, 0, s>>>((const double*)planar, stride, channels, n, (TO*)out, maxv, round_f32);
}
}  // namespace

void launch_deinterleave(const void* in, int fmt, int channels, int64_t n_frames, void* planar, int64_t stride,
                         int dtype, double inv_max, int round_f32, cudaStream_t s) {
    if (n_frames <= 0 || channels <= 0) return;
    count_launch();
    switch (fmt) {
        case 0: deint_dispatch<double>(in, channels, n_frames, planar, stride, dtype, 0.0, round_f32, s); break;
        case 1: deint_dispatch<float>(in, channels, n_frames, planar, stride, dtype, 0.0, round_f32, s); break;
        case 2: deint_dispatch<int16_t>(in, channels, n_frames, planar, stride, dtype, inv_max, round_f32, s); break;
        case 3: deint_dispatch<int32_t>(in, channels, n_frames, planar, stride, dtype, inv_max, round_f32, s); break;
        default: deint_dispatch<long long>(in, channels, n_frames, planar, stride, dtype, inv_max, round_f32, s); break;
    }
}

void launch_interleave(const void* planar, int64_t stride, int dtype, int channels, int64_t n_frames, void* out, int fmt,
                       double max_val, int round_f32, cudaStream_t s) {
    if (n_frames <= 0 || channels <= 0) return;
    count_launch();
    switch (fmt) {
        case 0: int_dispatch<double>(planar, stride, dtype, channels, n_frames, out, 0.0, round_f32, s); break;
        case 1: int_dispatch<float>(planar, stride, dtype, channels, n_frames, out, 0.0, round_f32, s); break;
        case 2: int_dispatch<int16_t>(planar, stride, dtype, channels, n_frames, out, max_val, round_f32, s); break;
        case 3: int_dispatch<int32_t>(planar, stride, dtype, channels, n_frames, out, max_val, round_f32, s); break;
        default: int_dispatch<long long>(planar, stride, dtype, channels, n_frames, out, max_val, round_f32, s); break;
    }
}

long long launch_count(bool reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

// FP64 tensor-core probe: 8 independent accumulator tiles per warp, mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) back to back.
// One instruction = 8x8x4 = 256 FMAs per warp = 8 per thread; 8 tiles x 8 reps = 64 instructions per iteration.
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double a0, double b0) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3 + i;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

float run_fma_probe(int dtype, int iters, double* flops, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    void* buf = nullptr;
    cudaMalloc(&buf, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](int n) {
        if (dtype == 3) dmma_probe_kernel<<<blocks, threads, 0, s>>>((double*)buf, n, 0.999999, 1e-7);
        else if (dtype == 2)  // packed f32x2: 4 reps x 8 chains x 2 lanes = 64 FMA per iteration, same as the others
            ffma2_probe_kernel<<<blocks, threads, 0, s>>>((float2*)buf, n, make_float2(0.999999f, 0.999998f),
                                                          make_float2(1e-7f, 2e-7f));
        else if (dtype == DT_F32) fma_probe_kernel<float><<<blocks, threads, 0, s>>>((float*)buf, n, 0.999999f, 1e-7f);
        else fma_probe_kernel<double><<<blocks, threads, 0, s>>>((double*)buf, n, 0.999999, 1e-7);
    };
    run(iters / 8 + 1);  // warm-up
    cudaEventRecord(e0, s);
    run(iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    // per thread and iteration: 64 scalar FMAs, or 64 DMMA instructions x 8 FMAs per thread
    *flops = 2.0 * 64.0 * (dtype == 3 ? 8.0 : 1.0) * (double)iters * (double)blocks * (double)threads;
    return ms;
}


}  // namespace gar
