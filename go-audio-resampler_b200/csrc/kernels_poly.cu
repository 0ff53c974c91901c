// kernels_poly.cu — arbitrary-ratio polyphase stage with cubic coefficient interpolation (polyphase_stage.go:186-312):
// K3 (one thread per output), K3i (lanes = lock-step rows, coefficients interpolated once per batch), K3m (the same
// contraction on the FP64 tensor cores), K3p (K3m as a TMA producer / MMA consumer pipeline with the coefficient matrix in
// registers), and launch_poly.
#include "mma_cores.cuh"
#include <type_traits>

namespace gar {
namespace {

// =============================================================================================
// K3i — polyphase stage for ANY ratio over a batch of lock-step rows, register-tiled (the batched form of
// polyphase_stage.go:186-312 with cubic coefficient interpolation).
//
// With an irrational ratio no two outputs of a stream share their coefficients, but the rows of a lock-step batch do:
// output n of every row uses the same phase and the same fraction x. So the LANES of a warp are 32 rows, and a warp
// task is RN adjacent outputs of those rows:
//   1. the warp evaluates the interpolated coefficients a + x(b + x(c + x d)) of its RN outputs ONCE (the three Horner
//      FMAs per tap amortise over the rows) into a per-task tile [tap][RN] in shared memory, each output's filter e_i
//      taps late (static window slots, as in the rational kernel: e_i = o_i - i*S + Dg, zero taps are exact no-ops);
//   2. every lane slides a register window over its row (rows at an odd pitch: conflict-free LDS.64): per tap one sample
//      LDS + RN/2 single-wavefront broadcast LDS.128 feed RN FMAs, sums strictly in tap order (bit-identical to
//      poly_kernel in float64).
// A block = one tile of 8*RN outputs (8 warp tasks) for up to 4 x 32 rows: the coefficient tiles are evaluated once
// and reused for every 32-row block, whose samples are staged in turn with asynchronous element copies; two blocks
// per SM overlap each other's load and compute phases. Trailing blocks write the carried tails.
// =============================================================================================
struct RowsGeom {
    int32_t TO, span, pitch, tp, D, n_tiles, nrb;  // outputs per tile, staged samples per row (max), row pitch, taps
                                                   // walked, tiles per row, 32-row blocks per thread block
};

template <typename T, int S, int RN>
__global__ void __launch_bounds__(256, 2) poly_rows_kernel(const PolyCall c, const RowsGeom g) {
    using V = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    static_assert(RN % VEC == 0, "a coefficient vector load covers whole outputs");
    constexpr int RB = 32;                 // rows per pass = lanes of a warp task
    constexpr int WN = (RN - 1) * S + 1;   // register window
    constexpr int NTASK = 8;               // warp tasks per block

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);                      // [RB][pitch] staged samples
    T* ct = xs + RB * g.pitch;                                   // [NTASK][tp*RN] coefficient tiles
    int* pat = reinterpret_cast<int*>(ct + NTASK * g.tp * RN);   // [NTASK][RN][4] phase row offset, lag, x bits

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_rg = (c.n_streams + RB * g.nrb - 1) / (RB * g.nrb);  // row groups
    const int n_work = g.n_tiles * n_rg;
    if ((int)blockIdx.x >= n_work) {  // carried tail of one row
        const int64_t row = (int)blockIdx.x - n_work;
        carry_row(static_cast<const T*>(c.hist) + row * c.hist_stride, c.hist_len,
                  static_cast<const T*>(c.in) + row * c.in_stride, c.n_in,
                  static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const int tile = blockIdx.x % g.n_tiles;
    const int rows_base = (blockIdx.x / g.n_tiles) * RB * g.nrb;
    const int64_t L = c.L;
    const int n0 = tile * g.TO;
    const int n1 = min(c.n_out, n0 + g.TO);  // outputs [n0, n1)
    // first staged sample: D before the window of output n0
    const int64_t d_base = (((c.at0 + (int64_t)n0 * c.step) >> 16) / L) - g.D;
    const int64_t d_last = (((c.at0 + (int64_t)(n1 - 1) * c.step) >> 16) / L);
    const int span_t = min((int)(d_last - d_base) + g.tp + 2 * WN + 2, g.span);
    const int64_t total = (int64_t)c.hist_len + c.n_in;

    // ---- stage the samples of 32 rows: warp w copies rows w, w+8, ... (coalesced along the row) ----
    // staged index i is sample d_base + i: zeros before the stream, [i0, i1) from the carried tail, [i1, i2) from `in`
    // (asynchronous element copies), zeros behind the end
    const int i0 = (int)min((int64_t)span_t, max((int64_t)0, -d_base));
    const int i1 = (int)min((int64_t)span_t, max((int64_t)i0, (int64_t)c.hist_len - d_base));
    const int i2 = (int)min((int64_t)span_t, max((int64_t)i1, total - d_base));
    auto stage_rows = [&](const int row0) {
        for (int r = warp; r < RB; r += 8) {
            const int64_t row = row0 + r;
            T* __restrict__ dst = xs + r * g.pitch;
            if (row >= c.n_streams) {
                for (int i = lane; i < span_t; i += 32) dst[i] = T(0);
                continue;
            }
            const T* __restrict__ hsrc = static_cast<const T*>(c.hist) + row * c.hist_stride + d_base;
            const T* __restrict__ isrc = static_cast<const T*>(c.in) + row * c.in_stride + (d_base - c.hist_len);
            for (int i = lane; i < i0; i += 32) dst[i] = T(0);
            for (int i = i0 + lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
            for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
            for (int i = i2 + lane; i < span_t; i += 32) dst[i] = T(0);
        }
    };
    stage_rows(rows_base);

    // ---- pattern + coefficient tile of this warp's task (overlaps the copies above) ----
    const int nf = n0 + warp * RN;  // first output of the task
    T* __restrict__ ctile = ct + warp * g.tp * RN;
    int* __restrict__ ptask = pat + warp * RN * 4;
    int div0 = 0, Dg = 0;
    {
        // lane i < RN: geometry of output nf + i (polyphase_stage.go:260-264)
        const int i = lane < RN ? lane : RN - 1;
        const int64_t at = c.at0 + (int64_t)(nf + i) * c.step;
        const int64_t full = at >> 16;
        const int64_t dv = full / L;
        const int ph = (int)(full - dv * L);
        const int dv0 = __shfl_sync(0xffffffffu, (int)(dv - d_base), 0);  // window offsets are relative to the tile
        const int o = (int)(dv - d_base) - dv0;
        int m = i * S - o;  // lag of the static slot behind the true offset
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, sft));
        Dg = m;
        div0 = dv0;
        if (lane < RN) {
            ptask[i * 4 + 0] = ph * c.taps;
            ptask[i * 4 + 1] = o - i * S + Dg;  // e_i in [0, D]
            ptask[i * 4 + 2] = (int)(at & 0xFFFF);
        }
    }
    __syncwarp();
    {
        const T* __restrict__ ga = static_cast<const T*>(c.bank_a);
        const T* __restrict__ gb = static_cast<const T*>(c.bank_b);
        const T* __restrict__ gc = static_cast<const T*>(c.bank_c);
        const T* __restrict__ gd = static_cast<const T*>(c.bank_d);
#pragma unroll 6
        for (int idx = lane; idx < g.tp * RN; idx += 32) {  // six independent (4-load) evaluations in flight per lane
            const int kk = idx / RN, i = idx - kk * RN;
            const int k = kk - ptask[i * 4 + 1];
            T v = T(0);
            if (k >= 0 && k < c.taps && nf + i < n1) {
                const int co = ptask[i * 4 + 0] + k;
                v = ga[co];
                if (c.interp) {
                    const T x = (T)ptask[i * 4 + 2] * (T)(1.0 / 65536.0);
                    v = fma(x, fma(x, fma(x, gd[co], gc[co]), gb[co]), v);
                }
            }
            ctile[idx] = v;
        }
    }

    // ---- tap loops: lane = row, RN adjacent outputs, static window slots; one 32-row block after the other ----
    for (int j = 0; j < g.nrb; ++j) {
        const int row0 = rows_base + j * RB;
        if (row0 >= c.n_streams) break;
        if (j > 0) {
            __syncthreads();  // everyone is done with the previous rows' samples
            stage_rows(row0);
        }
        cp_async_wait_all();
        __syncthreads();
        if (nf < n1) {
            const T* __restrict__ sp = xs + lane * g.pitch + (div0 - Dg);  // window slot 0 (div0 >= D >= Dg)
            T W[WN], acc[RN];
#pragma unroll
            for (int x = 0; x < WN; ++x) W[x] = sp[x];
#pragma unroll
            for (int i = 0; i < RN; ++i) acc[i] = T(0);
            auto tap = [&](const int u, const int kk) {
                T cf[RN];
#pragma unroll
                for (int q = 0; q < RN / VEC; ++q)
                    vec_unpack(*reinterpret_cast<const V*>(ctile + kk * RN + q * VEC), cf + q * VEC);
#pragma unroll
                for (int i = 0; i < RN; ++i) acc[i] = fma(W[(u + i * S) % WN], cf[i], acc[i]);
                W[u] = sp[kk + WN];
            };
            int it0 = 0;
            for (; it0 + WN <= g.tp; it0 += WN) {
#pragma unroll
                for (int u = 0; u < WN; ++u) tap(u, it0 + u);
            }
#pragma unroll
            for (int u = 0; u < WN; ++u)
                if (it0 + u < g.tp) tap(u, it0 + u);
            const int64_t row = row0 + lane;
            if (row < c.n_streams) {
                T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
#pragma unroll
                for (int i = 0; i < RN; ++i)
                    if (nf + i < n1) out[nf + i] = acc[i];
            }
        }
    }
}

template <typename T, int S, int RN>
static bool launch_poly_rows_t(const PolyCall& c, cudaStream_t s) {
    constexpr int WN = (RN - 1) * S + 1;
    RowsGeom g{};
    g.TO = 8 * RN;
    // intermediate samples per output r = step / (L * 65536); worst lag of a static slot behind the true offset
    const double r = (double)c.step / ((double)c.L * 65536.0);
    g.D = (RN - 1) * S - (int)std::floor((RN - 1) * r) + 1;
    if (g.D < 0 || g.D > 200) return false;
    g.tp = c.taps + g.D;
    g.span = (int)std::ceil((g.TO - 1) * r) + g.D + g.tp + 2 * WN + 4;
    g.pitch = g.span | 1;
    g.n_tiles = (c.n_out + g.TO - 1) / g.TO;
    // 32-row blocks per thread block: reuse every coefficient tile as often as possible while the grid still fills the GPU
    const int n_rb = (c.n_streams + 31) / 32;
    g.nrb = 1;
    while (g.nrb < 4 && g.nrb * 2 <= n_rb && (int64_t)g.n_tiles * ((n_rb + g.nrb * 2 - 1) / (g.nrb * 2)) >= 4 * 148) g.nrb *= 2;
    const size_t smem = ((size_t)32 * g.pitch + (size_t)8 * g.tp * RN) * sizeof(T) + (size_t)8 * RN * 4 * sizeof(int);
    if (smem > 113 * 1024) return false;
    auto k = poly_rows_kernel<T, S, RN>;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)g.n_tiles * ((c.n_streams + 32 * g.nrb - 1) / (32 * g.nrb)) + c.n_streams;
    k<<<(unsigned)blocks, 256, smem, s>>>(c, g);
    count_launch();
    return true;
}

// =============================================================================================
// K3m — K3i on the FP64 tensor cores. A warp task is 8 adjacent outputs x 32 rows:
//     D[i][s] = sum_w A[i][w] * X[w][s],   A[i][w] = coef_i[w - o_i]  (0 outside the filter),
// A = the task's interpolated coefficient matrix (8 x K, K = o_7 + taps rounded to 4; evaluated once per task as in K3i,
// stored [w][i] so that a k-step's fragment is 32 consecutive doubles), X = the rows' sample windows from the first
// output's offset on. Per k-step one A fragment (LDS.64) and four B fragments (8 rows each) feed four DMMA.8x8x4 —
// 1024 FMAs for 5 shared-memory loads, against 40 loads in K3i. No static window slots, no per-ratio template variants.
// 17-24 % of A is structural zeros (o_7 of K), still twice K3i's throughput. DMMA accumulates in window order = tap order.
// =============================================================================================
struct RowsMmaGeom {
    int32_t span, pitch, kp, n_tiles, nrb, nbuf;  // staged samples per row, row pitch, K (multiple of 4), tiles per row,
                                                  // 32-row blocks per thread block, sample buffers (2: prefetch)
};

template <int NTASK>
__global__ void __launch_bounds__(NTASK * 32, NTASK == 8 ? 2 : 3) poly_rows_mma_kernel(const PolyCall c, const RowsMmaGeom g) {
    constexpr int RB = 32, RN = 8, TO = RN * NTASK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);       // [2] mbarriers of the bulk-copied row buffers
    double* xs0 = reinterpret_cast<double*>(smem_raw + 16);      // [nbuf][RB][pitch] staged samples
    double* ct = xs0 + g.nbuf * RB * g.pitch;                    // [NTASK][kp][RN] coefficient matrices
    int* pat = reinterpret_cast<int*>(ct + NTASK * g.kp * RN);   // [NTASK][RN][4] phase row offset, window offset, x bits

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_rg = (c.n_streams + RB * g.nrb - 1) / (RB * g.nrb);
    const int n_work = g.n_tiles * n_rg;
    if ((int)blockIdx.x >= n_work) {  // carried tail of one row
        const int64_t row = (int)blockIdx.x - n_work;
        carry_row(static_cast<const double*>(c.hist) + row * c.hist_stride, c.hist_len,
                  static_cast<const double*>(c.in) + row * c.in_stride, c.n_in,
                  static_cast<double*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const int tile = blockIdx.x % g.n_tiles;
    const int rows_base = (blockIdx.x / g.n_tiles) * RB * g.nrb;
    const int64_t L = c.L;
    const int n0 = tile * TO;
    const int n1 = min(c.n_out, n0 + TO);
    const int64_t d_base = ((c.at0 + (int64_t)n0 * c.step) >> 16) / L;  // first staged sample = window of output n0
    const int64_t d_last = ((c.at0 + (int64_t)(n1 - 1) * c.step) >> 16) / L;
    const int span_t = min((int)(d_last - d_base) + g.kp + 4, g.span);
    const int64_t total = (int64_t)c.hist_len + c.n_in;
    const int i0 = 0;
    const int i1 = (int)min((int64_t)span_t, max((int64_t)i0, (int64_t)c.hist_len - d_base));
    const int i2 = (int)min((int64_t)span_t, max((int64_t)i1, total - d_base));
    // A 32-row block whose staged span lies inside `in` is moved by 32 TMA bulk copies (one thread), started `a` samples
    // early so that the sources are 16-byte aligned (all rows share the alignment when the row stride is even);
    // anything else (carried tail, end of the rows, ragged last block) by element copies. Returns the pad.
    const int64_t gi = d_base - c.hist_len;
    uint32_t ph0 = 0u, ph1 = 0u;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    auto stage_rows = [&](const int row0, const int buf, bool& bulk) -> int {
        double* xs = xs0 + buf * RB * g.pitch;
        bulk = false;
        if ((c.in_stride & 1) == 0 && row0 + RB <= c.n_streams && gi >= 0) {
            const double* __restrict__ src0 = static_cast<const double*>(c.in) + (int64_t)row0 * c.in_stride + gi;
            const int a = (int)((reinterpret_cast<uintptr_t>(src0) & 15u) >> 3);
            const int wlen = (span_t + a + 1) & ~1;
            if (gi - a >= 0 && gi - a + wlen <= c.n_in && wlen <= g.pitch) {
                bulk = true;
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(bar + buf, (uint32_t)(RB * wlen * sizeof(double)));
                    for (int r = 0; r < RB; ++r)
                        bulk_g2s(xs + r * g.pitch, src0 + (int64_t)r * c.in_stride - a, (uint32_t)(wlen * sizeof(double)), bar + buf);
                }
                return a;
            }
        }
        for (int r = warp; r < RB; r += NTASK) {
            const int64_t row = row0 + r;
            double* __restrict__ dst = xs + r * g.pitch;
            if (row >= c.n_streams) {
                for (int i = lane; i < span_t; i += 32) dst[i] = 0.0;
                continue;
            }
            const double* __restrict__ hsrc = static_cast<const double*>(c.hist) + row * c.hist_stride + d_base;
            const double* __restrict__ isrc = static_cast<const double*>(c.in) + row * c.in_stride + (d_base - c.hist_len);
            for (int i = i0 + lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
            for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
            for (int i = i2 + lane; i < span_t; i += 32) dst[i] = 0.0;
        }
        return 0;
    };
    __syncthreads();  // the mbarrier is initialised
    bool bulk = false, bulk_next = false;
    int apad = stage_rows(rows_base, 0, bulk), apad_next = 0;

    // ---- geometry + coefficient matrix of this warp's task (overlaps the copies above) ----
    const int nf = n0 + warp * RN;
    double* __restrict__ ctile = ct + warp * g.kp * RN;
    int* __restrict__ ptask = pat + warp * RN * 4;
    int base = 0;
    {
        const int i = lane < RN ? lane : RN - 1;  // lane i: output nf + i (polyphase_stage.go:260-264)
        const int64_t at = c.at0 + (int64_t)(nf + i) * c.step;
        const int64_t full = at >> 16;
        const int64_t dv = full / L;
        const int ph = (int)(full - dv * L);
        const int dv0 = __shfl_sync(0xffffffffu, (int)(dv - d_base), 0);
        base = dv0;
        if (lane < RN) {
            ptask[i * 4 + 0] = ph * c.taps;
            ptask[i * 4 + 1] = (int)(dv - d_base) - dv0;  // o_i
            ptask[i * 4 + 2] = (int)(at & 0xFFFF);
        }
    }
    __syncwarp();
    {
        const double* __restrict__ ga = static_cast<const double*>(c.bank_a);
        const double* __restrict__ gb = static_cast<const double*>(c.bank_b);
        const double* __restrict__ gc = static_cast<const double*>(c.bank_c);
        const double* __restrict__ gd = static_cast<const double*>(c.bank_d);
#pragma unroll 6
        for (int idx = lane; idx < g.kp * RN; idx += 32) {
            const int w = idx >> 3, i = idx & 7;
            const int k = w - ptask[i * 4 + 1];
            double v = 0.0;
            if (k >= 0 && k < c.taps && nf + i < n1) {
                const int co = ptask[i * 4 + 0] + k;
                v = ga[co];
                if (c.interp) {
                    const double x = (double)ptask[i * 4 + 2] * (1.0 / 65536.0);
                    v = fma(x, fma(x, fma(x, gd[co], gc[co]), gb[co]), v);
                }
            }
            ctile[idx] = v;
        }
    }

    const int nks = g.kp >> 2;
    for (int j = 0; j < g.nrb; ++j) {
        const int row0 = rows_base + j * RB;
        if (row0 >= c.n_streams) break;
        const int buf = g.nbuf == 2 ? (j & 1) : 0;
        const double* __restrict__ xs = xs0 + buf * RB * g.pitch;
        if (j > 0) {
            if (g.nbuf == 2) {  // staged by the previous iteration's prefetch
                bulk = bulk_next;
                apad = apad_next;
            } else {
                __syncthreads();  // everyone is done with the previous rows' samples
                apad = stage_rows(row0, 0, bulk);
            }
        }
        if (bulk) {
            const uint32_t ph = buf ? ph1 : ph0;
            while (!mbar_try_wait(bar + buf, ph)) {
            }
            if (buf) ph1 ^= 1u;
            else ph0 ^= 1u;
            if (j == 0) __syncthreads();  // also orders the first use after the set-up
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
        if (g.nbuf == 2 && j + 1 < g.nrb && row0 + RB < c.n_streams) {
            // prefetch the next 32 rows into the other buffer under this block of MMAs; that buffer was last read two
            // iterations ago, and every warp has passed this iteration's barrier / mbarrier wait since
            if (j > 0) __syncthreads();
            apad_next = stage_rows(row0 + RB, buf ^ 1, bulk_next);
        }
        if (nf < n1) {
            double acc[4][2];
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
            // A fragment: lane l holds A[i = l/4][w = 4*kk + l%4] = ctile[w][i]; B: X[w = 4*kk + l%4][row 8*t + l/4]
            const double* __restrict__ ap = ctile + (lane & 3) * RN + (lane >> 2);
            const double* __restrict__ bp = xs + (lane >> 2) * g.pitch + base + apad + (lane & 3);
#pragma unroll 2
            for (int kk = 0; kk < nks; ++kk) {
                const double a = ap[kk * 4 * RN];
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma884(acc[t][0], acc[t][1], a, bp[t * 8 * g.pitch + 4 * kk]);
            }
            const int i = lane >> 2;
            if (nf + i < n1) {
                if (c.out_f32) {  // float32(v) on the way out (constant.go:195-197)
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                        if (s0 < c.n_streams) (static_cast<float*>(c.out) + s0 * c.out_stride)[nf + i] = (float)acc[t][0];
                        if (s0 + 1 < c.n_streams) (static_cast<float*>(c.out) + (s0 + 1) * c.out_stride)[nf + i] = (float)acc[t][1];
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                        if (s0 < c.n_streams) (static_cast<double*>(c.out) + s0 * c.out_stride)[nf + i] = acc[t][0];
                        if (s0 + 1 < c.n_streams) (static_cast<double*>(c.out) + (s0 + 1) * c.out_stride)[nf + i] = acc[t][1];
                    }
                }
            }
        }
    }
}

// =============================================================================================
// K3p — K3m as a producer / consumer pipeline. Each of the 8 MMA warps keeps its task's whole coefficient matrix in
// REGISTERS (lane l holds A[i = l/4][4*kk + l%4] for every k-step: NK doubles, gathered straight from the L2-resident banks,
// no shared-memory copy), so shared memory holds nothing but two 32-row sample stages. A ninth warp is the producer: it waits
// for a stage to be released (mbarrier `empty`, one arrival per MMA warp), issues the 32 TMA bulk copies of the next 32 rows
// (mbarrier `full`, transaction bytes) and runs ahead of the MMA warps by one stage — no block-wide barrier in the loop,
// the copy latency of row block j+1 hides under the MMAs of row block j, and two such blocks still fit an SM.
// Per k-step 4 B fragments (LDS.64) feed 4 DMMA.8x8x4 (K3m: 5 loads). Same A values, same accumulation order as K3m:
// bit-identical results.
// =============================================================================================
template <int NK, int RB, int NST>
__global__ void __launch_bounds__(288, 2) poly_rows_pipe_kernel(const PolyCall c, const RowsMmaGeom g) {
    constexpr int RN = 8, NTASK = 8, TO = RN * NTASK, NT8 = RB / 8;
    static_assert(RB == 32 || RB == 16, "a stage is 32 or 16 rows");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);      // [NST] stage filled (producer arrival + TMA bytes)
    uint64_t* empty = full + NST;                                // [NST] stage released (8 MMA warps)
    double* xs0 = reinterpret_cast<double*>(smem_raw + 64);      // [NST][RB][pitch] staged samples

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_rg = (c.n_streams + RB * g.nrb - 1) / (RB * g.nrb);
    const int n_work = g.n_tiles * n_rg;
    if ((int)blockIdx.x >= n_work) {  // carried tail of one row
        const int64_t row = (int)blockIdx.x - n_work;
        carry_row(static_cast<const double*>(c.hist) + row * c.hist_stride, c.hist_len,
                  static_cast<const double*>(c.in) + row * c.in_stride, c.n_in,
                  static_cast<double*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const int tile = blockIdx.x % g.n_tiles;
    const int rows_base = (blockIdx.x / g.n_tiles) * RB * g.nrb;
    const int64_t L = c.L;
    const int n0 = tile * TO;
    const int n1 = min(c.n_out, n0 + TO);
    const int64_t d_base = ((c.at0 + (int64_t)n0 * c.step) >> 16) / L;  // first staged sample = window of output n0
    const int64_t d_last = ((c.at0 + (int64_t)(n1 - 1) * c.step) >> 16) / L;
    const int span_t = min((int)(d_last - d_base) + g.kp + 4, g.span);
    const int64_t total = (int64_t)c.hist_len + c.n_in;
    const int64_t gi = d_base - c.hist_len;
    const int nj = min(g.nrb, (c.n_streams - rows_base + RB - 1) / RB);  // 32-row stages of this block

    // stage kind, the same pure function on both sides of the pipeline: a full 32-row block whose span lies inside `in` is
    // moved by TMA, started `a` samples early (16-byte aligned sources; rows share the alignment when the stride is even)
    auto stage_is_bulk = [&](const int row0, int& a, int& wlen) -> bool {
        a = 0;
        wlen = 0;
        if ((c.in_stride & 1) != 0 || row0 + RB > c.n_streams || gi < 0) return false;
        const double* src0 = static_cast<const double*>(c.in) + (int64_t)row0 * c.in_stride + gi;
        a = (int)((reinterpret_cast<uintptr_t>(src0) & 15u) >> 3);
        wlen = (span_t + a + 1) & ~1;
        if (gi - a >= 0 && gi - a + wlen <= c.n_in && wlen <= g.pitch) return true;
        a = 0;
        return false;
    };

    if (tid == 0) {
        for (int b = 0; b < NST; ++b) {
            mbar_init(full + b, 1);
            mbar_init(empty + b, NTASK);
        }
    }
    __syncthreads();

    if (warp == NTASK) {
        // ---------------- producer warp ----------------
        for (int j = 0; j < nj; ++j) {
            const int buf = j % NST;
            const int row0 = rows_base + j * RB;
            double* xs = xs0 + buf * RB * g.pitch;
            if (j >= NST) {  // the MMA warps have released the stage's previous contents
                const uint32_t par = (uint32_t)((j / NST - 1) & 1);
                while (!mbar_try_wait(empty + buf, par)) {
                }
                __syncwarp();
            }
            int a, wlen;
            if (stage_is_bulk(row0, a, wlen)) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(full + buf, (uint32_t)(RB * wlen * sizeof(double)));
                }
                __syncwarp();
                if (lane < RB)
                    bulk_g2s(xs + lane * g.pitch, static_cast<const double*>(c.in) + (int64_t)(row0 + lane) * c.in_stride + gi - a,
                             (uint32_t)(wlen * sizeof(double)), full + buf);
            } else {  // edge stage (carried tail, end of the rows, ragged last row block): element copies by this warp
                const int i1 = (int)min((int64_t)span_t, max((int64_t)0, (int64_t)c.hist_len - d_base));
                const int i2 = (int)min((int64_t)span_t, max((int64_t)i1, total - d_base));
                for (int r = 0; r < RB; ++r) {
                    const int64_t row = row0 + r;
                    double* __restrict__ dst = xs + r * g.pitch;
                    if (row >= c.n_streams) {
                        for (int i = lane; i < span_t; i += 32) dst[i] = 0.0;
                        continue;
                    }
                    const double* __restrict__ hsrc = static_cast<const double*>(c.hist) + row * c.hist_stride + d_base;
                    const double* __restrict__ isrc = static_cast<const double*>(c.in) + row * c.in_stride + (d_base - c.hist_len);
                    for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
                    for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
                    for (int i = i2 + lane; i < span_t; i += 32) dst[i] = 0.0;
                }
                cp_async_wait_all();
                __threadfence_block();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + buf)) : "memory");
            }
        }
        return;
    }

    // ---------------- MMA warps: task = 8 outputs nf .. nf+7 ----------------
    const int nf = n0 + warp * RN;
    const int i = lane >> 2;  // this lane's output row of the MMA tile (polyphase_stage.go:260-264)
    const int64_t at = c.at0 + (int64_t)(nf + i) * c.step;
    const int64_t fullp = at >> 16;
    const int64_t dv = fullp / L;
    const int ph = (int)(fullp - dv * L);
    const int base = __shfl_sync(0xffffffffu, (int)(dv - d_base), 0);  // window offset of the task's first output
    const int o_i = (int)(dv - d_base) - base;
    const bool live = nf + i < n1;
    const int nks = g.kp >> 2;
    double A[NK];  // (gather and MMA loop: mma_cores.cuh, shared with the chain kernel K5)
    poly_gather_coeffs<NK>(A, c, ph, o_i, (double)(int)(at & 0xFFFF) * (1.0 / 65536.0), live, nks, lane);
    for (int j = 0; j < nj; ++j) {
        const int buf = j % NST;
        const int row0 = rows_base + j * RB;
        const double* __restrict__ xs = xs0 + buf * RB * g.pitch;
        int apad, wlen;
        stage_is_bulk(row0, apad, wlen);
        const uint32_t par = (uint32_t)((j / NST) & 1);
        while (!mbar_try_wait(full + buf, par)) {
        }
        __syncwarp();
        if (nf < n1) {
            double acc[NT8][2];
#pragma unroll
            for (int t = 0; t < NT8; ++t) acc[t][0] = acc[t][1] = 0.0;
            // B fragment: X[w = 4*kk + l%4][row 8*t + l/4]
            const double* __restrict__ bp = xs + (lane >> 2) * g.pitch + base + apad + (lane & 3);
            // (a straight-line block per k-step count, without the uniform predicates, measured 4 % SLOWER here: 871 vs 835 us)
            poly_mma_stage<NK, NT8>(acc, A, bp, g.pitch, nks);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + buf)) : "memory");
            if (live) {
                if (c.out_f32) {  // float32(v) on the way out (constant.go:195-197): no cast launch behind this one
#pragma unroll
                    for (int t = 0; t < NT8; ++t) {
                        const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                        if (s0 < c.n_streams) (static_cast<float*>(c.out) + s0 * c.out_stride)[nf + i] = (float)acc[t][0];
                        if (s0 + 1 < c.n_streams) (static_cast<float*>(c.out) + (s0 + 1) * c.out_stride)[nf + i] = (float)acc[t][1];
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < NT8; ++t) {
                        const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                        if (s0 < c.n_streams) (static_cast<double*>(c.out) + s0 * c.out_stride)[nf + i] = acc[t][0];
                        if (s0 + 1 < c.n_streams) (static_cast<double*>(c.out) + (s0 + 1) * c.out_stride)[nf + i] = acc[t][1];
                    }
                }
            }
        } else {
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + buf)) : "memory");
        }
    }
}

template <int NK, int RB, int NST>
static bool launch_poly_rows_pipe_t(const PolyCall& c, cudaStream_t s, const bool dry = false) {
    constexpr int TO = 64;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    RowsMmaGeom g{};
    const int omax = (int)std::ceil(7 * r) + 1;
    g.kp = ((omax + c.taps + 3) / 4) * 4;
    if (g.kp > 4 * NK || g.kp > 2 * c.taps + 8) return false;
    if (c.in_stride & 1) return false;  // no TMA (16-byte aligned rows): K3m stages with all of its warps instead
    g.span = (int)std::ceil((TO - 1) * r) + 1 + g.kp + 8;
    g.pitch = ((g.span + 2 + 15) / 16) * 16 + 4;  // rows 32 bytes apart modulo 128 (B fragment reads: two wavefronts)
    g.n_tiles = (c.n_out + TO - 1) / TO;
    const int n_rb = (c.n_streams + RB - 1) / RB;
    // stages (of RB rows) per block: the register-resident coefficients are gathered once per block
    g.nrb = 1;
    while (g.nrb < 256 / RB && g.nrb * 2 <= n_rb && (int64_t)g.n_tiles * ((n_rb + g.nrb * 2 - 1) / (g.nrb * 2)) >= 4 * 148) g.nrb *= 2;
    g.nbuf = NST;
    const size_t smem = 64 + (size_t)NST * RB * g.pitch * sizeof(double);
    if (smem > 113 * 1024) return false;  // two blocks per SM
    if (dry) return true;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(poly_rows_pipe_kernel<NK, RB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)g.n_tiles * ((c.n_streams + RB * g.nrb - 1) / (RB * g.nrb)) + c.n_streams;
    poly_rows_pipe_kernel<NK, RB, NST><<<(unsigned)blocks, 288, smem, s>>>(c, g);
    count_launch();
    return true;
}

// K3p variants: coefficient registers for K <= 80 or <= 112, 2 stages of 32 rows or (long spans) 3 stages of 16 rows
static bool launch_poly_rows_pipe(const PolyCall& c, cudaStream_t s, const bool dry = false) {
    static const int rb16_rows = [] { const char* e = gar::tune_env("GAR_K3P_RB16_ROWS"); return e ? std::atoi(e) : 0; }();
    // fewer rows: 16-row stages keep the pipeline at least two stages deep
    if (c.n_streams < rb16_rows && (launch_poly_rows_pipe_t<20, 16, 3>(c, s, dry) || launch_poly_rows_pipe_t<28, 16, 3>(c, s, dry))) return true;
    return launch_poly_rows_pipe_t<20, 32, 2>(c, s, dry) || launch_poly_rows_pipe_t<20, 16, 3>(c, s, dry) ||
           launch_poly_rows_pipe_t<28, 32, 2>(c, s, dry) || launch_poly_rows_pipe_t<28, 16, 3>(c, s, dry);
}

template <int NTASK>
static bool launch_poly_rows_mma_t(const PolyCall& c, cudaStream_t s, const bool dry = false, const bool one_per_sm = false) {
    constexpr int TO = 8 * NTASK;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    RowsMmaGeom g{};
    const int omax = (int)std::ceil(7 * r) + 1;
    g.kp = ((omax + c.taps + 3) / 4) * 4;
    // too many structural zeros: K3i — which stops at 8 samples per output; beyond that the alternative is the one-thread-per-output
    // kernel, and a coefficient matrix that is two thirds zeros still beats it several times over
    if (g.kp > (r > 8.0 ? 3 : 2) * c.taps + 8) return false;
    g.span = (int)std::ceil((TO - 1) * r) + 1 + g.kp + 8;
    g.pitch = ((g.span + 2 + 15) / 16) * 16 + 4;  // rows 32 bytes apart modulo 128 (B fragment reads: two wavefronts)
    g.n_tiles = (c.n_out + TO - 1) / TO;
    const int n_rb = (c.n_streams + 31) / 32;
    g.nrb = 1;
    // measured on the batched 44.1k->48k chain (256 rows): 1 / 2 / 4 / 8 row blocks per coefficient evaluation -> 1.93 / 1.45 /
    // 1.19 / 1.10 ms for the polyphase stage
    static const int max_nrb = [] { const char* e = gar::tune_env("GAR_K3M_NRB"); return e ? std::atoi(e) : 8; }();
    while (g.nrb < max_nrb && g.nrb * 2 <= n_rb && (int64_t)g.n_tiles * ((n_rb + g.nrb * 2 - 1) / (g.nrb * 2)) >= 4 * 148) g.nrb *= 2;
    static const int force_nbuf = [] { const char* e = gar::tune_env("GAR_K3M_NBUF"); return e ? std::atoi(e) : 0; }();
    const size_t fixed = 16 + (size_t)NTASK * g.kp * 8 * sizeof(double) + (size_t)NTASK * 8 * 4 * sizeof(int);
    const size_t xbytes = (size_t)32 * g.pitch * sizeof(double);
    g.nbuf = force_nbuf ? force_nbuf : 1;
    const size_t smem = fixed + g.nbuf * xbytes;
    // two blocks per SM normally; one (up to 227 KB) for steep ratios, whose tiles span several hundred samples per row
    if (smem > ((g.nbuf == 2 || one_per_sm) ? 227 : 113) * 1024) return false;
    if (dry) return true;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(poly_rows_mma_kernel<NTASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)g.n_tiles * ((c.n_streams + 32 * g.nrb - 1) / (32 * g.nrb)) + c.n_streams;
    poly_rows_mma_kernel<NTASK><<<(unsigned)blocks, NTASK * 32, smem, s>>>(c, g);
    count_launch();
    return true;
}

// 0: not taken, 1: K3m, 2: K3p. dry: eligibility only (no launch).
static int launch_poly_rows_mma(const PolyCall& c, cudaStream_t s, const bool dry = false) {
    if (!tensor_fir_enabled() || c.n_streams < 8 || (int64_t)c.n_out * c.n_streams < 16384 || c.L > 4096 || c.taps > 1024) return 0;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.0) || r > 16.0) return 0;
    static const int ntask = [] { const char* e = gar::tune_env("GAR_K3M_NTASK"); return e ? std::atoi(e) : 8; }();
    static const bool pipe = [] { const char* e = gar::tune_env("GAR_K3M_PIPE"); return !e || e[0] != '0'; }();
    static const int pipe_rows = [] { const char* e = gar::tune_env("GAR_K3M_PIPE_ROWS"); return e ? std::atoi(e) : 32; }();
    // measured (44.1k->48k, 21 M samples): K3p against K3m 32 rows 20.9 / 19.9, 48 rows 21.8 / 20.4 TFLOP/s, equal below
    if (pipe && c.n_streams >= pipe_rows && launch_poly_rows_pipe(c, s, dry)) return 2;
    if (ntask == 4) return (launch_poly_rows_mma_t<4>(c, s, dry) || launch_poly_rows_mma_t<8>(c, s, dry)) ? 1 : 0;
    if (launch_poly_rows_mma_t<8>(c, s, dry) || launch_poly_rows_mma_t<4>(c, s, dry)) return 1;
    // steep decimation (more than ~5 samples per output, e.g. the 384k -> 44.1k polyphase stage of a 192k -> 44.1k engine): the
    // tile does not fit half an SM; one block per SM still beats the one-thread-per-output kernel several times over
    return (launch_poly_rows_mma_t<4>(c, s, dry, true) || launch_poly_rows_mma_t<8>(c, s, dry, true)) ? 1 : 0;
}

// K3i dispatch: batches of at least 8 lock-step rows with enough outputs; S = ceil(samples per output)
template <typename T>
static bool launch_poly_rows(const PolyCall& c, cudaStream_t s) {
    if (c.n_streams < 8 || (int64_t)c.n_out * c.n_streams < 16384 || c.L > 4096 || c.taps > 1024) return false;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.0) || r > 8.0) return false;
    const int S = (int)std::ceil(r - 1e-12);
    switch (S) {
        case 1: return launch_poly_rows_t<T, 1, 8>(c, s);
        case 2: return launch_poly_rows_t<T, 2, 8>(c, s);
        case 3: return launch_poly_rows_t<T, 3, 6>(c, s);
        case 4: return launch_poly_rows_t<T, 4, 6>(c, s);
        case 5:
        case 6: return launch_poly_rows_t<T, 6, 4>(c, s);
        default: return launch_poly_rows_t<T, 8, 4>(c, s);
    }
}

// =============================================================================================
// Polyphase stage. One thread per output; the block's input span sits in shared memory.
// =============================================================================================
template <typename T, bool INTERP, int TO>
__global__ void __launch_bounds__(TO) poly_kernel(const PolyCall c, const int n_tiles, const int xcap,
                                                  const int cpitch /*> 0: gather every output's coefficient row by TMA*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* xs = reinterpret_cast<T*>(smem_raw + 16);
    T* crow = xs + ((xcap + 3) & ~3);  // [TO][cpitch]
    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    pdl_trigger_if_small();
    if (tile == n_tiles) {
        pdl_wait();
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int n0 = tile * TO;
    const int n1 = min(n0 + TO, c.n_out) - 1;
    const int64_t L = c.L;
    // this thread's output and its coefficient rows first: their lines are prefetched into L1 while the samples are staged
    const int n = n0 + threadIdx.x;
    const bool live = n < c.n_out;
    const int64_t at = c.at0 + (int64_t)(live ? n : n0) * c.step;
    const int64_t full = at >> 16;
    const int div = (int)(full / L);
    const int phase = (int)(full - (int64_t)div * L);
    const T x = (T)(int)(at & 0xFFFF) * (T)(1.0 / 65536.0);
    const int64_t co = (int64_t)phase * c.taps;
    const T* __restrict__ ca = static_cast<const T*>(c.bank_a) + co;
    const T* __restrict__ cb = static_cast<const T*>(c.bank_b) + co;
    const T* __restrict__ cc = static_cast<const T*>(c.bank_c) + co;
    const T* __restrict__ cd = static_cast<const T*>(c.bank_d) + co;
    // streaming-size launches: one TMA bulk copy per output fetches its coefficient row (all in flight at once) while the
    // samples are staged; larger launches prefetch the rows into L1
    // (interpolated coefficients, float64: the row of the INTERLEAVED bank, [tap]{a,b,c,d}, 32 bytes per tap, always aligned)
    const bool gather = cpitch > 0;
    const bool gather_il = INTERP && gather;
    const int cpad = gather_il ? 0 : (int)((reinterpret_cast<uintptr_t>(ca) & 15u) / sizeof(T));
    if (gather) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        const uint32_t nb = !live ? 0u : gather_il ? (uint32_t)(c.taps * 4 * sizeof(T)) : (uint32_t)(((c.taps + cpad) * sizeof(T) + 15) & ~(size_t)15);
        if (live)
            bulk_g2s(crow + (size_t)threadIdx.x * cpitch, gather_il ? static_cast<const T*>(c.bank_il) + co * 4 : ca - cpad, nb, bar);
        uint32_t bytes = nb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        // every warp adds its bytes without arriving; thread 0 alone arrives (init count 1) after its own warp's share
        if ((threadIdx.x & 31) == 0 && threadIdx.x != 0 && bytes)
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        __syncthreads();
        if (threadIdx.x == 0) mbar_expect_tx(bar, bytes);
    } else if (live) {
        prefetch_row_l1(ca, c.taps);
        if (INTERP) {
            prefetch_row_l1(cb, c.taps);
            prefetch_row_l1(cc, c.taps);
            prefetch_row_l1(cd, c.taps);
        }
    }
    const int div0 = (int)(((c.at0 + (int64_t)n0 * c.step) >> 16) / L);
    const int div1 = (int)(((c.at0 + (int64_t)n1 * c.step) >> 16) / L);
    const int span = div1 - div0 + c.taps;
    const bool staged = span <= xcap;
    pdl_wait();  // up to here: geometry and the constant banks only
    if (staged) {
        block_copy4(span, [&](int i) { return vload(hist, c.hist_len, in, c.n_in, div0 + i); }, [&](int i, T v) { xs[i] = v; });
        __syncthreads();
    }
    if (gather) {
        while (!mbar_try_wait(bar, 0)) {
        }
    }
    if (!live) return;
    const T* __restrict__ cg = gather ? crow + (size_t)threadIdx.x * cpitch + cpad : ca;
    // products of two float32 are exact in float64, so the float32 path only rounds once, at the store (two
    // interleaved chains); float64 sums strictly in tap order, like every other float64 kernel here, so the
    // fused kernels reproduce the stand-alone launches bit for bit
    double acc0 = 0, acc1 = 0;
    const int base = div - div0;
    for (int k = 0; k < c.taps; ++k) {
        T coef;
        if (gather_il) coef = fma(x, fma(x, fma(x, cg[4 * k + 3], cg[4 * k + 2]), cg[4 * k + 1]), cg[4 * k]);  // same operation order
        else {
            coef = cg[k];
            if (INTERP) coef = fma(x, fma(x, fma(x, cd[k], cc[k]), cb[k]), coef);
        }
        const T h = staged ? xs[base + k] : vload(hist, c.hist_len, in, c.n_in, div + k);
        if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h, (double)coef, acc1);
        else acc0 = fma((double)h, (double)coef, acc0);
    }
    (static_cast<T*>(c.out) + row * c.out_stride)[n] = (T)(acc0 + acc1);
}

}  // namespace

bool poly_rows_pipe_out32_takes(const PolyCall& c) {
    return c.n_out > 0 && tiled_polyphase_enabled() && launch_poly_rows_mma(c, nullptr, true) != 0;  // K3m and K3p narrow on the store
}

const char* launch_poly(const PolyCall& c, int dtype, cudaStream_t s, RatCache* cache) {
    if (c.n_streams <= 0) return "none";
    if (c.n_out <= 0) {
        launch_carry(c.hist, c.hist_stride, c.hist_len, c.in, c.in_stride, c.n_in, c.hist_out, c.hist_out_stride, c.drop,
                     c.new_hist_len, c.n_streams, dtype, s);
        return "carry";
    }
    // Batches of >= 8 lock-step rows: K3m, the polyphase stage on the FP64 tensor cores (any ratio)
    if (dtype == DT_F64 && tiled_polyphase_enabled()) {
        const int k = launch_poly_rows_mma(c, s);
        if (k == 2) return c.interp ? "poly_rows_mma_f64_interp_pipe" : "poly_rows_mma_f64_pipe";
        if (k == 1) return c.interp ? "poly_rows_mma_f64_interp" : "poly_rows_mma_f64";
    }
    if (c.out_f32) return nullptr;  // (the engine asked poly_rows_pipe_out32_takes first)
    // Batches of >= 8 rows with an even period length run K3i rather than K3r: K3r then stages its padded periods with
    // element copies from one warp (measured on the batched 48k->44.1k chain: 0.63 ms against 0.78 ms)
    if (dtype == DT_F64 && tiled_polyphase_enabled() && !c.interp && c.n_streams >= 8 && ((c.step >> 16) & 1) == 0 &&
        launch_poly_rows<double>(c, s))
        return "poly_rows_f64";
    if (dtype == DT_F64 && !c.interp && cache) {  // K3r: register-tiled rational-ratio kernel (large calls)
        FusedCall f{};
        f.in = c.in; f.in_stride = c.in_stride; f.n_in = c.n_in;
        f.hist_p = c.hist; f.hist_p_stride = c.hist_stride; f.hp = c.hist_len;
        f.hist_p_out = c.hist_out; f.hist_p_out_stride = c.hist_out_stride;
        f.drop_p = c.drop; f.new_hp = c.new_hist_len;
        f.bank_a = c.bank_a; f.t2 = c.taps; f.L = c.L; f.at0 = c.at0; f.step = c.step;
        f.n_out = c.n_out; f.interp = 0; f.out = c.out; f.out_stride = c.out_stride; f.n_streams = c.n_streams;
        if (launch_rat_poly_only_f64(f, s, cache)) return "poly_rat_f64";
    }
    if (dtype == DT_F64 && tiled_polyphase_enabled() && launch_poly_rows<double>(c, s)) return c.interp ? "poly_rows_f64_interp" : "poly_rows_f64";
    constexpr int TO = 128;
    const int n_tiles = (c.n_out + TO - 1) / TO;
    const int64_t blocks = (int64_t)(n_tiles + 1) * c.n_streams;
    const size_t esz = dtype == DT_F32 ? 4 : 8;
    // span of one tile: ((TO-1)*step >> 16)/L + 2 + taps, capped at 96 KB of shared memory
    int64_t span = (((int64_t)(TO - 1) * c.step) >> 16) / c.L + 2 + c.taps;
    const int64_t cap_words = (96 * 1024) / (int64_t)esz;
    int xcap = (int)(span < cap_words ? span : 0);  // 0 => read straight from global
    // streaming-size launches gather the outputs' coefficient rows by TMA (pitch: 16 bytes mod 128, room for the alignment pad)
    int cpitch = 0;
    if (!c.interp && blocks <= 2 * 148) {
        const int vec = 16 / (int)esz;
        int pitch = ((c.taps + vec + vec - 1) / vec) * vec;
        while ((pitch * (int)esz) % 128 != 16) pitch += vec;
        if ((size_t)(xcap + 4 + TO * pitch) * esz + 16 <= 160 * 1024) cpitch = pitch;
    } else if (c.interp && dtype == DT_F64 && c.bank_il && blocks <= 8) {
        // interpolated coefficients, Flush-size launches: each output's row of the interleaved bank (32 bytes per tap) by TMA —
        // the thread-per-output loop otherwise waits for four scattered loads per tap
        int pitch = 4 * c.taps;
        while ((pitch * 8) % 128 != 16) pitch += 2;
        if ((size_t)(xcap + 4 + TO * pitch) * 8 + 16 <= 200 * 1024) cpitch = pitch;
    }
    size_t smem = 16 + (size_t)((xcap + 3) & ~3) * esz + (size_t)TO * cpitch * esz;
#define LAUNCH(T, I)                                                                                          \
    {                                                                                                         \
        auto k = poly_kernel<T, I, TO>;                                                                       \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); \
        launch_pdl(k, (unsigned)blocks, (unsigned)TO, smem, s, c, n_tiles, xcap, cpitch);                     \
        count_launch();                                                                                       \
    }
    if (dtype == DT_F32) {
        if (c.interp) { LAUNCH(float, true); return "poly_f32_interp"; }
        LAUNCH(float, false);
        return "poly_f32";
    }
    if (c.interp) { LAUNCH(double, true); return "poly_f64_interp"; }
    LAUNCH(double, false);
    return "poly_f64";
#undef LAUNCH
}



}  // namespace gar
