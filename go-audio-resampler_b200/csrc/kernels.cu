// kernels.cu — hand-written sm_100a kernels for the resampling hot path.
//
//  fir_tiled_kernel   K1/K2: integer xL up-sampler and /M decimator as ONE strided
//                     multi-filter FIR.  Replaces simdops.ConvolveValidMulti + Interleave2
//                     (dft_stage.go:259-263,322-327) and DotProductUnsafe (dft_stage.go:531).
//                     The input window of a tile is staged in shared memory with a TMA
//                     bulk copy (cp.async.bulk + mbarrier) when the tile is regular, the
//                     filter bank sits beside it, and every thread slides a register
//                     window over R adjacent outputs so each 16-byte shared-memory read
//                     feeds VEC*R*NF FMAs (shared-memory bandwidth, not FMA issue, is the
//                     practical bound otherwise — SURVEY.md H3).
//  fir_generic_kernel fallback for unusual factors (any stride / phase count).
//  poly_kernel        K3: arbitrary-ratio polyphase stage with cubic coefficient
//                     interpolation; replaces simdops.CubicInterpDot (polyphase_stage.go:288).
//  carry / cast       streaming tail hand-over and f32<->f64 I/O casts.
//  fma_probe_kernel   dependent-FMA micro-benchmark for the roofline denominator.
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace gar {
namespace {

std::atomic<long long> g_launches{0};
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ void vec_unpack(const float4& v, float* d) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
__device__ __forceinline__ void vec_unpack(const double2& v, double* d) { d[0] = v.x; d[1] = v.y; }
__device__ __forceinline__ float4 vec_pack(const float* d) { return make_float4(d[0], d[1], d[2], d[3]); }
__device__ __forceinline__ double2 vec_pack(const double* d) { return make_double2(d[0], d[1]); }

// v[g] of the virtual input  hist ++ in  (zero outside)
template <typename T>
__device__ __forceinline__ T vload(const T* __restrict__ hist, int hist_len, const T* __restrict__ in, int n_in, int g) {
    if (g < 0) return T(0);
    if (g < hist_len) return hist[g];
    g -= hist_len;
    return g < n_in ? in[g] : T(0);
}

template <typename T>
__device__ __forceinline__ void carry_row(const T* hist, int hist_len, const T* in, int n_in, T* hist_out, int drop,
                                          int new_len) {
    for (int i = threadIdx.x; i < new_len; i += blockDim.x) hist_out[i] = vload(hist, hist_len, in, n_in, drop + i);
}

// ---- mbarrier / TMA bulk-copy helpers (PTX; SASS: SYNCS.*, UBLKCP) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// per-thread asynchronous global -> shared copy of one element (SASS LDGSTS); completion via cp.async.wait_all
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Register-tiled sliding-window FIR core shared by the stand-alone and the fused kernels: R adjacent positions
// (window stride M) x NF filters per thread, `xt` = this thread's window in shared memory (16-byte aligned),
// `cs` = [NF][cp] taps shifted by the pad `a`. One LDS.128 of samples + one broadcast LDS.128 per filter feed
// VEC*R*NF FMAs.
template <typename T, int M, int NF, int R>
__device__ __forceinline__ void fir_tile_accumulate(const T* __restrict__ xt, const T* __restrict__ cs, const int cp,
                                                    const int taps, const int a, T (&res)[R][NF]) {
    using V = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    constexpr int NCH = (M * (R - 1) + VEC - 1) / VEC + 1;  // 16-byte chunks spanned by one step's window
    constexpr int WREG = NCH * VEC;
    const int n_iter = (taps + a + VEC - 1) / VEC;
    T xr[WREG];
    T acc[R][NF];
    double tot[R][NF];  // f32: short f32 partial sums are folded into f64 totals (|err| ~ 1e-7)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < NF; ++p) {
            acc[r][p] = T(0);
            tot[r][p] = 0.0;
        }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) vec_unpack(*reinterpret_cast<const V*>(xt + ch * VEC), xr + ch * VEC);

    auto step = [&](const int u, const int it) {
        T cv[NF][VEC];
#pragma unroll
        for (int p = 0; p < NF; ++p) vec_unpack(*reinterpret_cast<const V*>(cs + p * cp + it * VEC), cv[p]);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const T xv = xr[(u * VEC + i + M * r) % WREG];
#pragma unroll
                for (int p = 0; p < NF; ++p) acc[r][p] = fma(xv, cv[p][i], acc[r][p]);
            }
        // the oldest chunk is dead now: refill its slot with chunk it+NCH (needed from the next step on)
        vec_unpack(*reinterpret_cast<const V*>(xt + (it + NCH) * VEC), xr + (u % NCH) * VEC);
    };
    auto fold = [&]() {
        if (sizeof(T) == 4) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int p = 0; p < NF; ++p) {
                    tot[r][p] += (double)acc[r][p];
                    acc[r][p] = T(0);
                }
        }
    };

    // f32 only: the three iterations that hold the centre of the main lobe are folded one by one, so the
    // many small terms that follow are never added to an O(1) float32 partial sum (keeps |err| ~ 1.5e-7).
    const int itf0 = ((taps - 1) / 2 + a) / VEC, itf1 = itf0 + 2;
    constexpr int FOLD_BODIES = 8;  // periodic fold every 8*NCH*VEC taps (F2F.F64.F32 is ~10x an FFMA slot)
    int it0 = 0, since_fold = 0;
    for (; it0 + NCH <= n_iter; it0 += NCH) {
        if (sizeof(T) == 4 && it0 <= itf1 && it0 + NCH > itf0) {
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
                step(u, it0 + u);
                if (it0 + u >= itf0 && it0 + u <= itf1) fold();
            }
        } else {
#pragma unroll
            for (int u = 0; u < NCH; ++u) step(u, it0 + u);
        }
        if (++since_fold == FOLD_BODIES) {
            fold();
            since_fold = 0;
        }
    }
#pragma unroll
    for (int u = 0; u < NCH; ++u)
        if (it0 + u < n_iter) {
            step(u, it0 + u);
            if (sizeof(T) == 4 && it0 + u >= itf0 && it0 + u <= itf1) fold();
        }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < NF; ++p) res[r][p] = (T)(tot[r][p] + (double)acc[r][p]);

}

// =============================================================================================
// Tiled strided multi-filter FIR.
//   M  = window stride between adjacent outputs (1 for the up-sampler, the decimation factor else)
//   NF = filters applied to the same window (up-sampling factor; 1 for the decimator)
//   R  = adjacent output positions per thread; M*R*sizeof(T)/16 is odd for the shipped variants so
//        the eight threads of a quarter-warp hit eight different 16-byte bank groups (conflict-free
//        LDS.128)
//   NT = threads per block; a block produces NT*R positions of one stream row.
// grid.x = rows * (n_tiles + 1): the extra block of every row writes the carried tail.
// =============================================================================================
template <typename T, int M, int NF, int R, int NT>
__global__ void __launch_bounds__(NT) fir_tiled_kernel(const FirCall c, const int n_tiles, const int tiles_per_block,
                                                       const int n_groups, const int cp /*padded taps*/,
                                                       const int xlen) {
    using V = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    constexpr int TJ = NT * R;
    static_assert((TJ * M * sizeof(T)) % 16 == 0, "tile stride must keep the 16-byte alignment of the window");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // two mbarriers, one per window buffer
    T* cs = reinterpret_cast<T*>(smem_raw + 16);            // [NF][cp]
    T* xs0 = cs + NF * cp;                                  // [2][xlen] double-buffered sample window

    // grid.x = rows * (n_groups + 1): a block owns `tiles_per_block` consecutive tiles of one row (filter
    // loaded once, TMA prefetch of tile k+1 under the FMAs of tile k); the extra block writes the carried tail.
    const int group = blockIdx.x % (n_groups + 1);
    const int64_t row = blockIdx.x / (n_groups + 1);
    const int tid = threadIdx.x;

    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;

    if (group == n_groups) {  // carry block
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int t_first = group * tiles_per_block;
    const int nt = min(tiles_per_block, n_tiles - t_first);

    // leading pad so that every tile's bulk source address is 16-byte aligned (constant along the row)
    const int a = (int)(((reinterpret_cast<uintptr_t>(in) / sizeof(T)) + (uintptr_t)(int64_t)(c.first - c.hist_len)) &
                        (uintptr_t)(VEC - 1));
    auto tile_geom = [&](const int t, int& g0a, int& words, bool& bulk) {
        const int j0 = t * TJ;
        const int tj = min(TJ, c.n_pos - j0);
        g0a = c.first + j0 * M - a;                                   // virtual index of xs[0]
        words = (((tj - 1) * M + c.taps + a + VEC - 1) / VEC) * VEC;  // samples the tile reads (16-byte units)
        const int gi = g0a - c.hist_len;                              // index into `in`
        bulk = gi >= 0 && gi + words <= c.n_in && words <= xlen;
    };
    auto issue_bulk = [&](const int t, const int buf) {  // one thread
        int g0a, words;
        bool bulk;
        tile_geom(t, g0a, words, bulk);
        if (bulk) {
            mbar_expect_tx(bar + buf, (uint32_t)(words * sizeof(T)));
            bulk_g2s(xs0 + buf * xlen, in + (g0a - c.hist_len), (uint32_t)(words * sizeof(T)), bar + buf);
        }
    };

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    for (int i = tid; i < 2 * xlen; i += NT) xs0[i] = T(0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before async-proxy (TMA) writes
    {  // filter bank, shifted by the pad, zero elsewhere
        const T* __restrict__ bank = static_cast<const T*>(c.bank);
#pragma unroll
        for (int p = 0; p < NF; ++p)
            for (int kk = tid; kk < cp; kk += NT) {
                const int k = kk - a;
                cs[p * cp + kk] = (k >= 0 && k < c.taps) ? bank[p * c.taps + k] : T(0);
            }
    }
    __syncthreads();
    if (tid == 0) issue_bulk(t_first, 0);
    uint32_t phase0 = 0u, phase1 = 0u;
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;

    for (int k = 0; k < nt; ++k) {
        const int t = t_first + k;
        const int buf = k & 1;
        T* xs = xs0 + buf * xlen;
        int g0a, words;
        bool bulk;
        tile_geom(t, g0a, words, bulk);
        if (tid == 0 && k + 1 < nt) issue_bulk(t + 1, buf ^ 1);  // prefetch (buffer free since the last barrier)
        if (bulk) {
            const uint32_t ph = buf ? phase1 : phase0;
            while (!mbar_try_wait(bar + buf, ph)) {
            }
            if (buf) phase1 ^= 1u;
            else phase0 ^= 1u;
        } else {  // edge tile (touches the carried tail or the end of the row): guarded loads
            for (int i = tid; i < xlen; i += NT)
                xs[i] = i < words ? vload(hist, c.hist_len, in, c.n_in, g0a + i) : T(0);
            __syncthreads();
        }

        // ---- register-tiled sliding-window FIR ----
        T res[R][NF];
        fir_tile_accumulate<T, M, NF, R>(xs + M * R * tid, cs, cp, c.taps, a, res);

        // ---- interleaved, vectorised store: out[(j*NF + p)] ----
        const int jb = t * TJ + R * tid;
        T* op = out + (int64_t)jb * NF;
        if (jb + R <= c.n_pos && (R * NF) % VEC == 0 && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) {
            const T* flat = &res[0][0];
#pragma unroll
            for (int q = 0; q < R * NF / VEC; ++q) reinterpret_cast<V*>(op)[q] = vec_pack(flat + q * VEC);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (jb + r < c.n_pos) {
#pragma unroll
                    for (int p = 0; p < NF; ++p) op[r * NF + p] = res[r][p];
                }
        }
        __syncthreads();  // every thread is done with xs[buf] before the next prefetch overwrites it
    }
}

// =============================================================================================
// K1m/K2m — float64 integer-factor FIR on the FP64 TENSOR cores (PTX mma.sync.m8n8k4.f64, SASS DMMA.8x8x4), for batches
// of >= 8 lock-step rows.
//
// JT = 8/NF consecutive positions x NF phases are the 8 rows of an MMA tile, 8 streams its 8 columns:
//     D[(jj,p)][s] = sum_w A[(jj,p)][w] * X[w][s],   A[(jj,p)][w] = bank[p][w - jj*M]  (0 outside the filter),
// X[w][s] = the sample window of stream s. A is a fixed 8 x (taps + (JT-1)*M) block-Toeplitz matrix (2 % zero padding
// for the x2 stage, 1 % for the 1223-tap /2 decimator): a real dense contraction, 256 FMAs per instruction instead of 32,
// no register-file or shared-memory pressure (measured: DMMA sustains 36.9 TFLOP/s on this B200, vector DFMA 34.1).
// A warp owns MT = 4 consecutive MMA tiles (4*JT positions): they see the same sample window shifted by SH = JT*M/4
// k-steps, so ONE B fragment (LDS.64) and ONE A fragment (LDS.64, kept in a rotating register window) feed 4 MMAs.
// A block = 8 streams x NW*MT*JT positions; A fragments are laid out per k-step in shared memory in fragment order.
// Taps are grouped in fours in window order, so results differ from the strictly sequential vector kernels in the last
// bits (1e-16 relative); identical call sequences are bit-identical.
// =============================================================================================
struct MmaGeom {
    int32_t nk, xlen, pitch, nbuf, n_tiles, tiles_per_block, n_groups, n_sg;  // k-steps, staged samples per stream, row pitch, window buffers
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int M, int NF, int NW, int MT>
__global__ void __launch_bounds__(NW * 32) fir_mma_f64_kernel(const FirCall c, const MmaGeom g) {
    constexpr int JT = 8 / NF;                // positions per MMA tile
    static_assert(8 % NF == 0 && (JT * M) % 4 == 0, "tile shift must be a whole number of k-steps");
    constexpr int SH = JT * M / 4;            // k-steps between consecutive MMA tiles
    constexpr int WA = (MT - 1) * SH + 1;     // rotating A-fragment window
    constexpr int TJ = NW * MT * JT;          // positions per block tile
    constexpr int NT = NW * 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // A[(jj,p)][w] = bank[p][w - jj*M] is a shifted copy of the filter in every row, so the A fragment of k-step kk is a
    // gather from the zero-padded bank with a per-lane offset: no fragment table, the bank itself is all that is staged
    constexpr int BOFF = (JT - 1) * M;                           // leading zeros: the largest negative offset
    const int blen = (4 * g.nk + BOFF + 5) & ~1;                 // padded filter length, even: windows stay 16-byte aligned
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);       // [2] mbarriers of the bulk-copied window buffers
    double* Bs = reinterpret_cast<double*>(smem_raw + 16);       // [NF][blen] zero-padded bank
    double* Xs0 = Bs + (size_t)NF * blen;                        // [nbuf][8][pitch] sample windows of the block's streams

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_work = g.n_sg * g.n_groups;
    if ((int)blockIdx.x >= n_work) {  // carried tail of one row
        const int64_t row = (int)blockIdx.x - n_work;
        carry_row(static_cast<const double*>(c.hist) + row * c.hist_stride, c.hist_len,
                  static_cast<const double*>(c.in) + row * c.in_stride, c.n_in,
                  static_cast<double*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const int grp = blockIdx.x % g.n_groups;
    const int sbase = (blockIdx.x / g.n_groups) * 8;
    const int t_first = grp * g.tiles_per_block;
    const int nt = min(g.tiles_per_block, g.n_tiles - t_first);

    {
        const double* __restrict__ bank = static_cast<const double*>(c.bank);
        for (int idx = tid; idx < NF * blen; idx += NT) {
            const int p = idx / blen, k = idx - p * blen - BOFF;
            Bs[idx] = (k >= 0 && k < c.taps) ? bank[p * c.taps + k] : 0.0;
        }
    }
    const int64_t total = (int64_t)c.hist_len + c.n_in;
    const int nq = g.nk + (MT - 1) * SH;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    uint32_t ph0 = 0u, ph1 = 0u;
    // TMA bulk copies need 16-byte aligned sources: all 8 rows share the alignment when the row stride is even
    const bool rows_bulk = (c.in_stride & 1) == 0 && sbase + 8 <= c.n_streams;
    const double* __restrict__ in0 = static_cast<const double*>(c.in) + (int64_t)sbase * c.in_stride;
    const int xbuf = 8 * g.pitch;

    // geometry of local tile kt: first sample v0, staged length, bulk-copy parameters
    auto tile_geom = [&](const int kt, int64_t& v0, int& len, int& a, int& wlen) -> bool {
        const int jb0 = (t_first + kt) * TJ;
        v0 = (int64_t)c.first + (int64_t)jb0 * M;
        const int npos_t = min(TJ, c.n_pos - jb0);
        len = min(g.xlen, ((npos_t + JT - 1) / JT * JT - 1) * M + 4 * g.nk + 4);
        a = 0;
        wlen = 0;
        const int64_t gi = v0 - c.hist_len;
        if (!rows_bulk || gi < 0) return false;
        a = (int)((reinterpret_cast<uintptr_t>(in0 + gi) & 15u) >> 3);  // start `a` samples early: aligned source
        wlen = (len + a + 1) & ~1;
        if (gi - a >= 0 && gi - a + wlen <= c.n_in && wlen <= g.pitch) return true;
        a = 0;  // element copies start exactly at v0
        return false;
    };
    auto issue = [&](const int kt, const int buf) {  // one thread
        int64_t v0;
        int len, a, wlen;
        if (!tile_geom(kt, v0, len, a, wlen)) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar + buf, (uint32_t)(8 * wlen * sizeof(double)));
        const int64_t gi = v0 - c.hist_len - a;
        for (int r = 0; r < 8; ++r)
            bulk_g2s(Xs0 + buf * xbuf + r * g.pitch, in0 + (int64_t)r * c.in_stride + gi, (uint32_t)(wlen * sizeof(double)),
                     bar + buf);
    };

    for (int kt = 0; kt < nt; ++kt) {
        const int buf = g.nbuf == 2 ? (kt & 1) : 0;
        double* __restrict__ Xs = Xs0 + buf * xbuf;
        int64_t v0;
        int len, a, wlen;
        const bool bulk = tile_geom(kt, v0, len, a, wlen);
        // everyone is done with the windows this iteration overwrites (first tile: the bank and the mbarriers are set)
        __syncthreads();
        if (tid == 0) {
            if (g.nbuf == 2) {  // prefetch the next tile under this tile's MMAs
                if (kt == 0) issue(0, 0);
                if (kt + 1 < nt) issue(kt + 1, buf ^ 1);
            } else {
                issue(kt, 0);
            }
        }
        if (bulk) {
            const uint32_t ph = buf ? ph1 : ph0;
            while (!mbar_try_wait(bar + buf, ph)) {
            }
            if (buf) ph1 ^= 1u;
            else ph0 ^= 1u;
        } else {  // edge tile (touches the carried tail or the end of the rows): element copies
            const int i1 = (int)min((int64_t)len, max((int64_t)0, (int64_t)c.hist_len - v0));
            const int i2 = (int)min((int64_t)len, max((int64_t)i1, total - v0));
            for (int r = warp; r < 8; r += NW) {
                const int64_t row = sbase + r;
                double* __restrict__ dst = Xs + r * g.pitch;
                if (row >= c.n_streams) {
                    for (int i = lane; i < len; i += 32) dst[i] = 0.0;
                    continue;
                }
                const double* __restrict__ hsrc = static_cast<const double*>(c.hist) + row * c.hist_stride + v0;
                const double* __restrict__ isrc = static_cast<const double*>(c.in) + row * c.in_stride + (v0 - c.hist_len);
                for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
                for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
                for (int i = i2 + lane; i < len; i += 32) dst[i] = 0.0;
            }
            cp_async_wait_all();
            __syncthreads();
        }

        // ---- MT MMA tiles per warp; step q loads window chunk Q = warp*MT*SH + q and A fragment q ----
        const int jb0 = (t_first + kt) * TJ;
        const int npos_t = min(TJ, c.n_pos - jb0);
        if (warp * MT * JT < npos_t) {
            double acc[MT][2];
#pragma unroll
            for (int b = 0; b < MT; ++b) acc[b][0] = acc[b][1] = 0.0;
            double Areg[WA];
            // B fragment: lane l reads X[w = 4*Q + l%4][stream l/4]
            const double* __restrict__ xw = Xs + (lane >> 2) * g.pitch + (lane & 3) + a + 4 * (warp * MT * SH);
            // A fragment: lane l holds A[row = l/4][w = 4*kk + l%4] = bank[p][4*kk + l%4 - jj*M], row = jj*NF + p
            const double* __restrict__ aw = Bs + ((lane >> 2) % NF) * blen + BOFF + (lane & 3) - ((lane >> 2) / NF) * M;
            for (int q0 = 0; q0 < nq; q0 += WA) {
#pragma unroll
                for (int u = 0; u < WA; ++u) {
                    const int q = q0 + u;
                    if (q < nq) {
                        Areg[u] = q < g.nk ? aw[4 * q] : 0.0;
                        const double bf = xw[4 * q];
#pragma unroll
                        for (int b = 0; b < MT; ++b) {
                            const int kk = q - b * SH;
                            if (kk >= 0 && kk < g.nk) dmma884(acc[b][0], acc[b][1], Areg[((u - b * SH) % WA + WA) % WA], bf);
                        }
                    }
                }
            }
            // ---- D[row = lane/4][cols 2*(lane%4), +1]: output jb*NF + row of streams sbase + col ----
            const int r8 = lane >> 2;
            const int s0 = sbase + 2 * (lane & 3);
#pragma unroll
            for (int b = 0; b < MT; ++b) {
                const int jb = jb0 + (warp * MT + b) * JT;
                if (jb + r8 / NF < c.n_pos) {
                    const int64_t o = (int64_t)jb * NF + r8;
                    if (s0 < c.n_streams) (static_cast<double*>(c.out) + (int64_t)s0 * c.out_stride)[o] = acc[b][0];
                    if (s0 + 1 < c.n_streams) (static_cast<double*>(c.out) + (int64_t)(s0 + 1) * c.out_stride)[o] = acc[b][1];
                }
            }
        }
    }
}

template <int M, int NF>
static bool launch_fir_mma_t(const FirCall& c, cudaStream_t s) {
    constexpr int JT = 8 / NF, SH = JT * M / 4;
    int dev = 0;
    cudaGetDevice(&dev);
    MmaGeom g{};
    const int kp = c.taps + (JT - 1) * M;
    g.nk = (kp + 3) / 4;
    g.n_sg = (c.n_streams + 7) / 8;
    const size_t bank_bytes = 16 + (size_t)NF * ((4 * g.nk + (JT - 1) * M + 5) & ~1) * sizeof(double);
    auto run = [&](auto kernel, const int NW, const int MT, const int slot) -> bool {
        const int TJ = NW * MT * JT;
        g.xlen = (TJ - 1) * M + 4 * g.nk + 4 * (MT - 1) * SH + 10;
        g.pitch = ((g.xlen + 15) / 16) * 16 + 4;  // rows 32 bytes apart modulo 128: the 8 x 32-byte B fragment reads tile two wavefronts
        const size_t xbytes = (size_t)8 * g.pitch * sizeof(double);
        // two window buffers (the next tile is prefetched under the MMAs) when at least two such blocks fit an SM
        static const int force_nbuf = [] { const char* e = std::getenv("GAR_MMA_NBUF"); return e ? std::atoi(e) : 0; }();
        g.nbuf = bank_bytes + 2 * xbytes <= 110 * 1024 ? 2 : 1;
        if (force_nbuf == 1) g.nbuf = 1;
        const size_t smem = bank_bytes + g.nbuf * xbytes;
        if (smem > 227 * 1024) return false;
        g.n_tiles = (c.n_pos + TJ - 1) / TJ;
        // persistent over a few tiles (bank staged once, prefetch) while the grid still fills the GPU
        const int64_t blocks_per_sm = std::max<int64_t>(1, (int64_t)(227 * 1024) / (int64_t)(smem + 1024));
        const int64_t slots = 148 * std::min<int64_t>(blocks_per_sm, 2048 / (NW * 32));
        int64_t tpb = (int64_t)g.n_tiles * g.n_sg / (slots * 4);
        tpb = std::max<int64_t>(1, std::min<int64_t>(tpb, 8));
        g.tiles_per_block = (int32_t)std::min<int64_t>(tpb, g.n_tiles);
        g.n_groups = (g.n_tiles + g.tiles_per_block - 1) / g.tiles_per_block;
        static size_t configured[64][4] = {{0}};
        size_t& conf = configured[dev & 63][slot];
        if (smem > conf) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            conf = smem;
        }
        const int64_t blocks = (int64_t)g.n_sg * g.n_groups + c.n_streams;
        kernel<<<(unsigned)blocks, NW * 32, smem, s>>>(c, g);
        count_launch();
        return true;
    };
    // Measured on B200 (C3: 8 ch x 1223 taps /2; x2 stage of the batched 44.1k->48k chain): 16 warps x 4 tiles for long
    // filters (one block per SM: the window is dominated by the taps), 8 warps x 4 tiles in several blocks per SM for
    // short ones; 6 or 8 tiles per warp and smaller blocks were slower. GAR_MMA_CFG = 1 / 3 forces one of the two.
    static const int forced = [] { const char* e = std::getenv("GAR_MMA_CFG"); return e ? std::atoi(e) : -1; }();
    if (forced == 1) return run(fir_mma_f64_kernel<M, NF, 16, 4>, 16, 4, 1);
    if (forced == 3) return run(fir_mma_f64_kernel<M, NF, 8, 4>, 8, 4, 3);
    if (c.taps > 600) return run(fir_mma_f64_kernel<M, NF, 16, 4>, 16, 4, 1) || run(fir_mma_f64_kernel<M, NF, 8, 4>, 8, 4, 3);
    return run(fir_mma_f64_kernel<M, NF, 8, 4>, 8, 4, 3);
}

static bool g_fir_mma = [] {
    const char* e = std::getenv("GAR_NO_MMA");
    return !(e && e[0] && e[0] != '0');
}();
// float64 FIR on the FP64 tensor cores: batches of >= 8 lock-step rows, x2 up-sampler and /2 /3 /4 decimators
static const char* launch_fir_mma(const FirCall& c, cudaStream_t s) {
    if (!g_fir_mma || c.n_streams < 8 || (int64_t)c.n_pos * c.n_streams < 32768 || c.taps < 16) return nullptr;
    if (c.stride == 1 && c.nf == 2) return launch_fir_mma_t<1, 2>(c, s) ? "fir_f64_mma_up2" : nullptr;
    if (c.nf == 1 && c.stride == 2) return launch_fir_mma_t<2, 1>(c, s) ? "fir_f64_mma_s2" : nullptr;
    if (c.nf == 1 && c.stride == 3) return launch_fir_mma_t<3, 1>(c, s) ? "fir_f64_mma_s3" : nullptr;
    if (c.nf == 1 && c.stride == 4) return launch_fir_mma_t<4, 1>(c, s) ? "fir_f64_mma_s4" : nullptr;
    return nullptr;
}

// =============================================================================================
// float32 decimator with packed FMAs (PTX fma.rn.f32x2 -> SASS FFMA2, new on sm_100).
// Same tiling as fir_tiled_kernel, but adjacent taps are paired: (x[k],x[k+1]) * (c[k],c[k+1]) is ONE
// instruction on even/odd register pairs. Pairs always span both register banks, so the bank conflicts that
// cap the scalar-FFMA version at ~78 % issue utilisation cannot occur, and the FMA work needs half the issue
// slots. A position whose window offset M*r is odd reads its samples from the even address below and uses a
// copy of the filter shifted by one tap (c'[k] = c[k-1]); lane .x sums even taps, lane .y odd taps.
// =============================================================================================
template <int M, int NF, int R, int NT>
__global__ void __launch_bounds__(NT) fir_f32x2_kernel(const FirCall c, const int n_tiles, const int tiles_per_block,
                                                       const int n_groups, const int cp, const int xlen) {
    typedef unsigned long long u64;
    constexpr int MAXE = M * (R - 1) - ((M * (R - 1)) & 1);
    constexpr int NCH = (MAXE + 4 + 3) / 4;
    constexpr int NS = ((M & 1) && R > 1) ? 2 : 1;  // filter copies: shift 0 (even offsets), shift 1 (odd offsets)
    constexpr int TJ = NT * R;
    static_assert((TJ * M) % 4 == 0, "tile stride must keep the 16-byte alignment of the window");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // two mbarriers, one per window buffer
    float* cs = reinterpret_cast<float*>(smem_raw + 16);    // [NF][NS][cp]
    float* xs0 = cs + NF * NS * cp;                         // [2][xlen] double-buffered sample window

    // grid.x = rows * (n_groups + 1): a block owns `tiles_per_block` consecutive tiles of one row; the extra
    // block of every row writes the carried tail.
    const int group = blockIdx.x % (n_groups + 1);
    const int64_t row = blockIdx.x / (n_groups + 1);
    const int tid = threadIdx.x;
    const float* __restrict__ hist = static_cast<const float*>(c.hist) + row * c.hist_stride;
    const float* __restrict__ in = static_cast<const float*>(c.in) + row * c.in_stride;
    if (group == n_groups) {
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<float*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int t_first = group * tiles_per_block;
    const int nt = min(tiles_per_block, n_tiles - t_first);

    // Leading pad `a` (same for every tile of the row because the tile stride is a multiple of 16 bytes): the
    // window of a tile starts `a` samples early so that its global address is 16-byte aligned for the TMA
    // bulk copy; the filter is shifted by `a` zero taps to compensate.
    const int a = (int)(((reinterpret_cast<uintptr_t>(in) >> 2) + (uintptr_t)(int64_t)(c.first - c.hist_len)) & 3u);
    auto tile_geom = [&](const int t, int& g0a, int& words, bool& bulk) {
        const int j0 = t * TJ;
        const int tj = min(TJ, c.n_pos - j0);
        g0a = c.first + j0 * M - a;                       // virtual index of xs[0]
        words = (((tj - 1) * M + c.taps + a + 3) / 4) * 4;  // samples the tile reads, rounded to 16 bytes
        const int gi = g0a - c.hist_len;                  // index into `in`
        bulk = gi >= 0 && gi + words <= c.n_in && words <= xlen;
    };
    auto issue_bulk = [&](const int t, const int buf) {  // one thread
        int g0a, words;
        bool bulk;
        tile_geom(t, g0a, words, bulk);
        if (bulk) {
            mbar_expect_tx(bar + buf, (uint32_t)(words * sizeof(float)));
            bulk_g2s(xs0 + buf * xlen, in + (g0a - c.hist_len), (uint32_t)(words * sizeof(float)), bar + buf);
        }
    };

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    // both window buffers start as zeros: whatever a later, shorter tile leaves behind is finite
    for (int i = tid; i < 2 * xlen; i += NT) xs0[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before async-proxy (TMA) writes
    {
        const float* __restrict__ bank = static_cast<const float*>(c.bank);
#pragma unroll
        for (int p = 0; p < NF; ++p)
#pragma unroll
            for (int sh = 0; sh < NS; ++sh)
                for (int kk = tid; kk < cp; kk += NT) {
                    const int k = kk - a - sh;
                    cs[(p * NS + sh) * cp + kk] = (k >= 0 && k < c.taps) ? bank[p * c.taps + k] : 0.f;
                }
    }
    __syncthreads();
    if (tid == 0) issue_bulk(t_first, 0);
    uint32_t phase0 = 0u, phase1 = 0u;

    const int n_iter = (c.taps + a + (NS - 1) + 3) / 4;
    const int itf0 = ((c.taps - 1) / 2 + a) / 4, itf1 = itf0 + 2;  // centre-of-main-lobe folds (see fir_tiled_kernel)
    float* __restrict__ out = static_cast<float*>(c.out) + row * c.out_stride;

    for (int k = 0; k < nt; ++k) {
        const int t = t_first + k;
        const int buf = k & 1;
        float* xs = xs0 + buf * xlen;
        int g0a, words;
        bool bulk;
        tile_geom(t, g0a, words, bulk);
        // prefetch the next tile into the other buffer (free since the __syncthreads that ended tile k-1)
        if (tid == 0 && k + 1 < nt) issue_bulk(t + 1, buf ^ 1);
        if (bulk) {
            const uint32_t ph = buf ? phase1 : phase0;
            while (!mbar_try_wait(bar + buf, ph)) {
            }
            if (buf) phase1 ^= 1u;
            else phase0 ^= 1u;
        } else {  // edge tile (touches the carried tail or the end of the row): guarded loads
            for (int i = tid; i < xlen; i += NT)
                xs[i] = i < words ? vload(hist, c.hist_len, in, c.n_in, g0a + i) : 0.f;
            __syncthreads();
        }

        // ---- register-tiled sliding window on packed FMAs ----
        const float* xt = xs + M * R * tid;
        u64 xw[NCH * 2];
        u64 acc[R][NF];
        double tot[R][NF];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < NF; ++p) {
                acc[r][p] = 0ull;
                tot[r][p] = 0.0;
            }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xt + ch * 4);
            xw[ch * 2] = v.x;
            xw[ch * 2 + 1] = v.y;
        }
        auto step = [&](const int u, const int it) {
            // (tools/probe_fma_patterns2.cu: FFMA2 loses ~15 % when one coefficient pair feeds 6 FMAs in a row;
            //  duplicating the pair with a second, unmergeable ld.shared cost more than it gained — measured)
            u64 cv[NF][NS][2];
#pragma unroll
            for (int p = 0; p < NF; ++p)
#pragma unroll
                for (int sh = 0; sh < NS; ++sh) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(cs + (p * NS + sh) * cp + it * 4);
                    cv[p][sh][0] = v.x;
                    cv[p][sh][1] = v.y;
                }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int sh = (M * r) & 1;
                    const int eh = (M * r - sh) / 2;
                    const u64 xv = xw[(u * 2 + q + eh) % (NCH * 2)];
#pragma unroll
                    for (int p = 0; p < NF; ++p) {
                        u64 d;
                        asm("fma.rn.f32x2 %0, %1, %2, %3;"
                            : "=l"(d)
                            : "l"(xv), "l"(cv[p][NS == 2 ? sh : 0][q]), "l"(acc[r][p]));
                        acc[r][p] = d;
                    }
                }
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xt + (it + NCH) * 4);
            xw[(u % NCH) * 2] = v.x;
            xw[(u % NCH) * 2 + 1] = v.y;
        };
        auto fold = [&]() {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int p = 0; p < NF; ++p) {
                    const float lo = __uint_as_float((unsigned)(acc[r][p] & 0xffffffffull));
                    const float hi = __uint_as_float((unsigned)(acc[r][p] >> 32));
                    tot[r][p] += (double)lo + (double)hi;
                    acc[r][p] = 0ull;
                }
        };
        // three straight loops (plain bodies, the one or two bodies holding the centre, plain bodies): no
        // per-body bookkeeping; float32 partial sums are folded into float64 only around the centre of the
        // main lobe and at the loop boundaries (the tails stay small, see fir_tile_accumulate)
        const int nb = n_iter / NCH;
        const int bs0 = min(nb, itf0 / NCH), bs1 = min(nb, itf1 / NCH + 1);
        int b = 0;
        for (; b < bs0; ++b) {
#pragma unroll
            for (int u = 0; u < NCH; ++u) step(u, b * NCH + u);
        }
        fold();
        for (; b < bs1; ++b) {
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
                step(u, b * NCH + u);
                if (b * NCH + u >= itf0 && b * NCH + u <= itf1) fold();
            }
        }
        for (; b < nb; ++b) {
#pragma unroll
            for (int u = 0; u < NCH; ++u) step(u, b * NCH + u);
        }
        const int it0 = nb * NCH;
#pragma unroll
        for (int u = 0; u < NCH; ++u)
            if (it0 + u < n_iter) {
                step(u, it0 + u);
                if (it0 + u >= itf0 && it0 + u <= itf1) fold();
            }
        fold();

        // ---- vectorised store ----
        const int jb = t * TJ + R * tid;
        float* op = out + (int64_t)jb * NF;
        if (jb + R <= c.n_pos && (R * NF) % 4 == 0 && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) {
#pragma unroll
            for (int q = 0; q < R * NF / 4; ++q) {
                float4 v;
                v.x = (float)tot[(q * 4 + 0) / NF][(q * 4 + 0) % NF];
                v.y = (float)tot[(q * 4 + 1) / NF][(q * 4 + 1) % NF];
                v.z = (float)tot[(q * 4 + 2) / NF][(q * 4 + 2) % NF];
                v.w = (float)tot[(q * 4 + 3) / NF][(q * 4 + 3) % NF];
                reinterpret_cast<float4*>(op)[q] = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (jb + r < c.n_pos) {
#pragma unroll
                    for (int p = 0; p < NF; ++p) op[r * NF + p] = (float)tot[r][p];
                }
        }
        __syncthreads();  // every thread is done with xs[buf] before the next prefetch overwrites it
    }
}



// Carried tails of a fused x2 -> polyphase call (one block per row): the x2 stage's new tail is a plain copy,
// the polyphase stage's new tail needs the last few intermediate samples, recomputed here with the same
// strictly sequential chain as the tile core (bit-identical in float64).
template <typename T>
__device__ __forceinline__ void fused_carry_tails_rt(const FusedCall& c, const int64_t row, T* scratch, const int scratch_cap) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
    carry_row(hist_u, c.hu, in, c.n_in, static_cast<T*>(c.hist_u_out) + row * c.hist_u_out_stride, c.drop_u, c.new_hu);
    T* hp_out = static_cast<T*>(c.hist_p_out) + row * c.hist_p_out_stride;
    // intermediate samples [j_lo, j_hi) are needed: stage their input span and the x2 bank in shared memory first, so the
    // strictly sequential chains below run on shared-memory latency instead of one global round trip per tap (this block
    // is on the critical path of every streaming-size call)
    const int j_lo = max(0, c.drop_p - c.hp), j_hi = c.drop_p + c.new_hp - c.hp;
    const int p_lo = j_lo >> 1;
    const int span = j_hi > j_lo ? (((j_hi - 1) >> 1) - p_lo) + c.t1 : 0;
    const bool staged = span > 0 && span + 2 * c.t1 <= scratch_cap;
    T* xb = scratch;
    T* bk = scratch + span;
    if (staged) {
        for (int i = tid; i < span; i += NT) xb[i] = vload(hist_u, c.hu, in, c.n_in, p_lo + i);
        for (int i = tid; i < 2 * c.t1; i += NT) bk[i] = bank_u[i];
        __syncthreads();
    }
    for (int i = tid; i < c.new_hp; i += NT) {
        const int idx = c.drop_p + i;  // index into vp = hist_p ++ mid
        T v;
        if (idx < c.hp) {
            v = hist_p[idx];
        } else {
            const int j = idx - c.hp;
            if (staged) {
                const T* __restrict__ xx = xb + ((j >> 1) - p_lo);
                const T* __restrict__ cc = bk + (j & 1) * c.t1;
                if (sizeof(T) == 8) {  // same order as the tile core: bit-identical
                    T acc = 0;
                    for (int t = 0; t < c.t1; ++t) acc = fma(xx[t], cc[t], acc);
                    v = acc;
                } else {
                    double acc = 0;
                    for (int t = 0; t < c.t1; ++t) acc = fma((double)xx[t], (double)cc[t], acc);
                    v = (T)acc;
                }
            } else {
                const T* __restrict__ bkg = bank_u + (j & 1) * c.t1;
                if (sizeof(T) == 8) {
                    T acc = 0;
                    for (int t = 0; t < c.t1; ++t) acc = fma(vload(hist_u, c.hu, in, c.n_in, (j >> 1) + t), bkg[t], acc);
                    v = acc;
                } else {
                    double acc = 0;
                    for (int t = 0; t < c.t1; ++t)
                        acc = fma((double)vload(hist_u, c.hu, in, c.n_in, (j >> 1) + t), (double)bkg[t], acc);
                    v = (T)acc;
                }
            }
        }
        hp_out[i] = v;
    }
}

// =============================================================================================
// K4 — fused x2 up-sampler + polyphase stage. One block = one tile of MT = 2*NT*R intermediate samples:
//   1. stage the input window (TMA bulk copy when regular) and the x2 bank in shared memory,
//   2. run the register-tiled FIR core and write the tile's intermediate samples to SHARED memory,
//   3. produce every polyphase output whose T2-sample window lies in the tile (tiles overlap by T2-1
//      intermediate samples, recomputed rather than exchanged: 4 % for T2 = 64).
// Block 0 of a row also sees the polyphase stage's carried tail in front of its tile. The last block of
// every row writes both carried tails (the polyphase tail needs the last few intermediate samples, which
// it recomputes directly).
// =============================================================================================
template <typename T, bool INTERP, int R, int NT>
__global__ void __launch_bounds__(NT) fused_up2_poly_kernel(const FusedCall c, const int n_tiles, const int ms,
                                                            const int cp, const int xlen, const int hpf,
                                                            const int bank_pitch /*0: read banks through L1*/) {
    using V = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    constexpr int NF = 2;
    constexpr int TP = NT * R;      // positions per tile
    constexpr int MT = TP * NF;     // intermediate samples per tile

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* cs = reinterpret_cast<T*>(smem_raw + 16);  // [2][cp]
    T* xs = cs + NF * cp;                         // [xlen]
    T* vp = xs + xlen;                            // [hpf + MT] polyphase input: (tail |) intermediate tile
    T* pbank = vp + hpf + MT;                     // [L][bank_pitch] a-bank copy (optional)

    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const int tid = threadIdx.x;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
    const int n_mid = c.np * NF;

    if (tile == n_tiles) {  // ---- carried tails ----
        fused_carry_tails_rt<T>(c, row, xs, xlen + hpf + MT);
        return;
    }

    // ---- 1. stage the x2 stage's input window ----
    const int p0 = (tile * ms) >> 1;  // first position of the tile (ms is even)
    const int tp = min(TP, c.np - p0);
    const int need = tp > 0 ? tp - 1 + c.t1 : 0;
    int a = 0;
    bool bulk = false;
    {
        const int gi = p0 - c.hu;
        if (gi >= 0 && need > 0) {
            const uintptr_t addr = reinterpret_cast<uintptr_t>(in + gi);
            const int mis = (int)((addr & 15u) / sizeof(T));
            const int words = ((need + mis + VEC - 1) / VEC) * VEC;
            if (gi - mis >= 0 && gi - mis + words <= c.n_in && words <= xlen) {
                bulk = true;
                a = mis;
            }
        }
    }
    if (bulk) {
        const int gi = p0 - c.hu - a;
        const int words = ((need + a + VEC - 1) / VEC) * VEC;
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(words * sizeof(T)));
            bulk_g2s(xs, in + gi, (uint32_t)(words * sizeof(T)), bar);
        }
        for (int i = words + tid; i < xlen; i += NT) xs[i] = T(0);
    } else {
        for (int i = tid; i < xlen; i += NT) xs[i] = i < need ? vload(hist_u, c.hu, in, c.n_in, p0 + i) : T(0);
    }
    for (int i = tid; i < NF * cp; i += NT) {
        const int p = i / cp, k = i % cp - a;
        cs[i] = (k >= 0 && k < c.t1) ? bank_u[p * c.t1 + k] : T(0);
    }
    // polyphase side: carried tail right-aligned in front of tile 0 (the tile itself starts 16-byte aligned),
    // optional a-bank copy (odd pitch: conflict-free rows)
    const int front = tile == 0 ? hpf : 0;
    if (tile == 0)
        for (int i = tid; i < c.hp; i += NT) vp[hpf - c.hp + i] = hist_p[i];
    if (bank_pitch > 0) {
        const T* __restrict__ ba = static_cast<const T*>(c.bank_a);
#pragma unroll 8
        for (int i = tid; i < c.L * c.t2; i += NT) pbank[(i / c.t2) * bank_pitch + (i % c.t2)] = ba[i];  // 8 loads in flight
    }
    __syncthreads();
    if (bulk) {
        while (!mbar_try_wait(bar, 0)) {
        }
    }

    // ---- 2. x2 FIR core -> intermediate tile in shared memory ----
    {
        T res[R][NF];
        fir_tile_accumulate<T, 1, NF, R>(xs + R * tid, cs, cp, c.t1, a, res);
        T* mp = vp + front + (size_t)R * NF * tid;
#pragma unroll
        for (int q = 0; q < R * NF / VEC; ++q) reinterpret_cast<V*>(mp)[q] = vec_pack(&res[0][0] + q * VEC);
    }
    __syncthreads();

    // ---- 3. polyphase outputs whose window starts in [lo, hi) of vp = hist_p ++ mid ----
    const int64_t Lq = (int64_t)c.L << 16;
    const int64_t lo = tile == 0 ? 0 : (int64_t)c.hp + (int64_t)tile * ms;
    const int64_t hi = (int64_t)c.hp + (int64_t)(tile + 1) * ms;
    auto first_n = [&](const int64_t d) -> int64_t {  // smallest n with div_n >= d
        const int64_t need_at = d * Lq - c.at0;
        return need_at <= 0 ? 0 : (need_at + c.step - 1) / c.step;
    };
    const int64_t n_lo = min((int64_t)c.n_out, first_n(lo));
    const int64_t n_hi = tile == n_tiles - 1 ? (int64_t)c.n_out : min((int64_t)c.n_out, first_n(hi));
    // virtual vp index d lives at shared-memory element d - vbase
    const int64_t vbase = tile == 0 ? (int64_t)c.hp - hpf : lo;
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
    const T* __restrict__ ga = static_cast<const T*>(c.bank_a);
    const T* __restrict__ gb = static_cast<const T*>(c.bank_b);
    const T* __restrict__ gc = static_cast<const T*>(c.bank_c);
    const T* __restrict__ gd = static_cast<const T*>(c.bank_d);
    (void)n_mid;
    for (int64_t n = n_lo + tid; n < n_hi; n += NT) {
        const int64_t at = c.at0 + n * c.step;
        const int64_t full = at >> 16;
        const int64_t div = full / c.L;
        const int phase = (int)(full - div * c.L);
        const T* h = vp + (div - vbase);
        double acc0 = 0, acc1 = 0;
        if (INTERP) {
            const T x = (T)(int)(at & 0xFFFF) * (T)(1.0 / 65536.0);
            const int64_t co = (int64_t)phase * c.t2;
            for (int k = 0; k < c.t2; ++k) {
                const T coef = fma(x, fma(x, fma(x, gd[co + k], gc[co + k]), gb[co + k]), ga[co + k]);
                if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h[k], (double)coef, acc1);
                else acc0 = fma((double)h[k], (double)coef, acc0);
            }
        } else {
            const T* __restrict__ ca = bank_pitch > 0 ? pbank + phase * bank_pitch : ga + (int64_t)phase * c.t2;
            for (int k = 0; k < c.t2; ++k) {
                if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h[k], (double)ca[k], acc1);
                else acc0 = fma((double)h[k], (double)ca[k], acc0);
            }
        }
        out[n] = (T)(acc0 + acc1);
    }
}

// =============================================================================================
// K4r — fused x2 up-sampler + polyphase stage for RATIONAL ratios (step and phase accumulator have no
// fractional bits: 44.1k<->48k, 8k->12k, ... every ratio whose L/M is exact), register-tiled.
//
// The polyphase stage is periodic: L outputs consume exactly Mi = step>>16 intermediate samples, and output
// n + L uses the same phase filter as output n on a window Mi samples later. A tile is P such periods of one
// row (P = 16 or 32). Work on a tile has two kinds of items:
//   x2 chunk   32 thread tasks of the register-tiled x2 core (fir_tile_accumulate, FMA-bound): R positions each,
//              input window staged by a TMA bulk copy, results written to the tile's intermediate buffer in
//              SHARED memory (period j at pitch Mi + PAD; PAD = 1 when Mi is even keeps the period lanes on
//              different banks);
//   poly task  a warp owns RN adjacent outputs of the period pattern and its LANES are the periods, so all
//              lanes use the same phase filters -> every coefficient read is a shared-memory BROADCAST, and a
//              lane slides a register window over its period: per tap, 1 sample LDS + RN broadcast coefficient
//              LDS feed RN FMAs (the stand-alone kernel needs 2 loads per FMA). Output i of a thread sits at a
//              STATIC window slot i*S (S = ceil(Mi/L)); the true offset lags the slot by e_i in [0, D] samples,
//              absorbed by reading its filter row (stored with D zero taps on both sides) e_i taps early.
//              Sums run strictly in tap order (zero taps are exact no-ops): float64 results are bit-identical
//              to the stand-alone kernels.
// The intermediate buffer is double-buffered and a block is persistent over consecutive tiles of a row: in one
// iteration its warps pull items from ONE queue holding the poly tasks of tile t-1 and the x2 chunks of tile t,
// so the shared-memory-bound and the FMA-bound work overlap, quantisation of either kind is absorbed by the
// other, and there is one barrier per tile. The extra block of every row writes both carried tails.
// =============================================================================================
struct RatGeom {
    int32_t Mi, P, D, tp, gpitch, vlen, G;  // see launch_fused_rat_t
    int32_t cp, xlen, xbufs, nv;            // x2 filter (padded), input window capacity, window / intermediate buffers
    int32_t n_tiles, tiles_per_block, n_groups;
    const void* cg_src;  // this launch's coefficient tile in global memory ([G][gpitch] T, then goff[G] ints)
    uint32_t cg_bytes;   // its size, a multiple of 16
};

// Builds the coefficient tile of one start phase F0 (RatCache): for output group gi = outputs [gi*RN, gi*RN + RN) of the
// period, tile[gi][tap][i] = a-bank[phase_i][tap - e_i] (zero outside the filter), then goff[gi] = window offset.
// Output i sits at window slot i*S; its true offset lags the slot by e_i = o_i - i*S + Dg taps.
template <typename T, int S, int RN>
__global__ void __launch_bounds__(256) rat_build_tile_kernel(const T* __restrict__ bank_a, const int t2, const int L,
                                                             const int Mi, const int F0, const int G, const int tp,
                                                             const int gpitch, T* __restrict__ tile) {
    extern __shared__ int gtab_s[];  // [G*RN] phase << 8 | lag
    int* goff = reinterpret_cast<int*>(tile + (size_t)G * gpitch);
    const int qM = Mi / L, rM = Mi - qM * L;
    for (int gi = threadIdx.x; gi < G; gi += blockDim.x) {
        const unsigned rf = (unsigned)(F0 + gi * RN * Mi);
        const int div0 = (int)(rf / (unsigned)L);
        int ph = (int)rf - div0 * L, dv = 0, Dg = 0;
        int o[RN], php[RN];
#pragma unroll
        for (int i = 0; i < RN; ++i) {
            o[i] = dv;
            php[i] = ph;
            Dg = max(Dg, i * S - dv);
            ph += rM;
            dv += qM;
            if (ph >= L) {
                ph -= L;
                ++dv;
            }
        }
        goff[gi] = div0 - Dg;  // >= -D
#pragma unroll
        for (int i = 0; i < RN; ++i) gtab_s[gi * RN + i] = (php[i] << 8) | (o[i] - i * S + Dg);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * gpitch; idx += blockDim.x) {
        const int gi = idx / gpitch, rem = idx - gi * gpitch;
        T v = T(0);
        if (rem < tp * RN) {
            const int kk = rem / RN, i = rem - kk * RN;
            const int pe = gtab_s[gi * RN + i];
            const int k = kk - (pe & 255);
            if (k >= 0 && k < t2) v = bank_a[(pe >> 8) * t2 + k];
        }
        tile[idx] = v;
    }
}

constexpr int RAT_MAXT = 16;                 // tiles per block (pick_tiles_per_block caps at 16)
constexpr int RAT_CTL_INTS = 8 + 11 * (RAT_MAXT + 4);
constexpr int RAT_CTL_BYTES = ((RAT_CTL_INTS * 4 + 15) / 16) * 16;

template <typename T, int S, int RN, int PAD, bool FUSED>
__global__ void __launch_bounds__(512, 1) fused_up2_rat_kernel(const FusedCall c, const RatGeom g) {
    using V = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    static_assert(RN % VEC == 0, "a coefficient vector load covers whole outputs");
    constexpr int NF = 2;
    constexpr int R = sizeof(T) == 8 ? 6 : 12;  // x2 core: positions per thread task
    constexpr int WN = (RN - 1) * S + 1;        // register window of the polyphase phase
    const int NT = blockDim.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // [4] one mbarrier per input window buffer, [4]: the tile
    int* ctl = reinterpret_cast<int*>(smem_raw + 48);
    int* qhead = ctl;                                        // work-queue head
    int* in_done = ctl + 8;                                  // [t] input item of local tile t finished
    int* up_done = in_done + RAT_MAXT + 4;                   // [t] x2 chunks of tile t finished
    int* po_done = up_done + RAT_MAXT + 4;                   // [t] poly tasks of tile t finished
    int* gstart = po_done + RAT_MAXT + 4;                    // [k] first queue item of group k
    int* nchunk = gstart + RAT_MAXT + 4;                     // [t] x2 chunks of tile t
    int* tflags = nchunk + RAT_MAXT + 4;                     // [t] bit0: bulk input, bit1: mbarrier parity
    int* tg_src = tflags + RAT_MAXT + 4;                     // [t] tile geometry: i_lo - hu (index into `in`)
    int* tg_npos = tg_src + RAT_MAXT + 4;                    // [t] x2 positions
    int* tg_wend = tg_npos + RAT_MAXT + 4;                   // [t] samples of vp the tile covers
    int* tg_words = tg_wend + RAT_MAXT + 4;                  // [t] bulk-copy length in samples
    int* tg_woff = tg_words + RAT_MAXT + 4;                  // [t] w of the first sample of the first position
    T* cs = reinterpret_cast<T*>(smem_raw + 48 + RAT_CTL_BYTES);  // [2][cp]       x2 bank
    T* xs0 = cs + NF * g.cp;                                      // [xbufs][xlen] x2 input windows
    T* vs0 = xs0 + (FUSED ? g.xbufs * g.xlen : 0);                // [nv][vlen]    intermediate samples of a tile
    T* cg = vs0 + g.nv * g.vlen;  // [G][gpitch] polyphase coefficients, one tile [tap][RN outputs] per output group
    int* goff = reinterpret_cast<int*>(cg + (size_t)g.G * g.gpitch);  // [G] window offset of the group (part of the tile)

    const int group = blockIdx.x % (g.n_groups + 1);
    const int64_t row = blockIdx.x / (g.n_groups + 1);
    const int tid = threadIdx.x, lane = tid & 31;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
    if (group == g.n_groups) {  // carried tails
        if (FUSED) fused_carry_tails_rt<T>(c, row, xs0, g.xbufs * g.xlen + g.nv * g.vlen);
        else carry_row(hist_p, c.hp, in, c.n_in, static_cast<T*>(c.hist_p_out) + row * c.hist_p_out_stride, c.drop_p, c.new_hp);
        return;
    }

    const int Mi = g.Mi, P = g.P, D = g.D, L = c.L;
    const int FM = D + 1;  // front margin of the intermediate buffer (window starts up to D + PAD before w = 0)
    // polyphase input vp = carried tail ++ (FUSED: the x2 stage's 2*np outputs, made here; else the stage's input `in`)
    const int64_t total_vp = (int64_t)c.hp + (FUSED ? 2 * (int64_t)c.np : (int64_t)c.n_in);
    const int t_first = group * g.tiles_per_block;
    const int nt = min(g.tiles_per_block, g.n_tiles - t_first);
    const int gpw = 32 / P, jl = lane & (P - 1), gsub = lane / P;
    const int n_wt = (g.G + gpw - 1) / gpw;

    // geometry of tile t: x2 positions [i_lo, i_lo + n_pos) feed vp[m0, m0 + wend)
    auto tile_geom = [&](const int t, int64_t& m0, int& wend, int64_t& i_lo, int& n_pos, bool& bulk, int& words) {
        m0 = (int64_t)t * P * Mi;
        wend = (int)min((int64_t)P * Mi + c.t2 - 1, total_vp - m0);
        i_lo = 0;
        n_pos = 0;
        bulk = false;
        words = 0;
        if (!FUSED) return;
        const int64_t j_lo = max((int64_t)0, m0 - c.hp), j_hi = m0 + wend - c.hp;
        i_lo = j_lo >> 1;
        const int64_t i_hi = (j_hi + 1) >> 1;
        n_pos = i_hi > i_lo ? (int)(i_hi - i_lo) : 0;
        int64_t gi = i_lo - c.hu;
        if (gi >= 0 && n_pos > 0) {
            const int mis = (int)((reinterpret_cast<uintptr_t>(in + gi) & 15u) / sizeof(T));
            if (gi - mis >= 0) {  // start `mis` positions early: 16-byte aligned source, results below w = 0 are dropped
                gi -= mis;
                const int w = ((n_pos + mis - 1 + c.t1 + VEC - 1) / VEC) * VEC;
                if (gi + w <= c.n_in && w <= g.xlen) {
                    bulk = true;
                    words = w;
                    i_lo -= mis;
                    n_pos += mis;
                }
            }
        }
    };

    // ---- block set-up: control words, queue layout, zeroed buffers, banks ----
    const int NX = FUSED ? g.xbufs : g.nv;  // input prefetch depth (poly-only: the input lands in the tile buffers)
    if (tid == 0) {
        for (int b = 0; b < 5; ++b) mbar_init(bar + b, 1);
        *qhead = 0;
        // this launch's coefficient tile (built once per start phase, RatCache): one bulk copy, overlapped with the set-up
        mbar_expect_tx(bar + 4, g.cg_bytes);
        bulk_g2s(cg, g.cg_src, g.cg_bytes, bar + 4);
    }
    if (tid < nt) {  // geometry of the block's tiles, one thread each
        const int k = tid;
        int64_t m0, i_lo;
        int wend, n_pos, words;
        bool bulk;
        tile_geom(t_first + k, m0, wend, i_lo, n_pos, bulk, words);
        nchunk[k] = ((n_pos + R - 1) / R + 31) / 32;
        tflags[k] = bulk ? 1 : 0;
        tg_src[k] = (int)(i_lo - c.hu);
        tg_npos[k] = n_pos;
        tg_wend[k] = wend;
        tg_words[k] = words;
        tg_woff[k] = (int)((int64_t)c.hp + 2 * i_lo - m0);
        if (!FUSED) {
            // poly-only: samples w in [w0, wend) come from `in`. With an unpadded period pitch they are one contiguous
            // run: a TMA bulk copy moves the 16-byte aligned middle (the tile's front margin absorbs the parity between
            // source and destination), single elements at either end are copied by hand.
            int w0 = (int)max((int64_t)0, (int64_t)c.hp - m0);
            int fm = FM, nb = 0;
            int64_t e0 = m0 + w0 - c.hp;
            if (PAD == 0 && sizeof(T) == 8 && w0 < wend) {
                if (reinterpret_cast<uintptr_t>(in + e0) & 15u) {
                    ++w0;
                    ++e0;
                }
                fm = FM + ((FM + w0) & 1);
                nb = (wend - w0) & ~1;
                if (nb < 0) nb = 0;
            }
            tflags[k] = nb > 0 ? 1 : 0;
            tg_src[k] = (int)e0;  // first bulk element of `in`
            tg_words[k] = nb;     // bulk elements
            tg_woff[k] = w0;      // w of the first bulk element
            tg_npos[k] = fm;      // front margin of this tile's buffer
        }
        in_done[k] = 0;
        up_done[k] = 0;
        po_done[k] = 0;
    }
    for (int i = tid; i < (FUSED ? g.xbufs * g.xlen : 0) + g.nv * g.vlen; i += NT) xs0[i] = T(0);  // xs, vs adjacent
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (FUSED) {
        const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
        for (int i = tid; i < NF * g.cp; i += NT) {
            const int p = i / g.cp, k = i - p * g.cp;
            cs[i] = k < c.t1 ? bank_u[p * c.t1 + k] : T(0);
        }
    }
    __syncthreads();
    if (tid == 0) {
        // mbarrier parities and the queue: group 0 = {I(0..NX-1), U(0)*}; group k = {I(k+NX-1), U(k)*, P(k-1)*};
        // group nt = {P(nt-1)*}
        unsigned parbits = 0u;  // current mbarrier parity of every window buffer
        int at = 0;
        for (int k = 0; k <= nt; ++k) {
            gstart[k] = at;
            if (k < nt) {
                const int xb = k % NX;
                if (tflags[k] & 1) {
                    tflags[k] |= (int)(((parbits >> xb) & 1u) << 1);
                    parbits ^= 1u << xb;
                }
            }
            if (k == 0) at += min(NX, nt) + nchunk[0];
            else if (k < nt) at += (k + NX - 1 < nt ? 1 : 0) + nchunk[k] + n_wt;
            else at += n_wt;
        }
        gstart[nt + 1] = at;
    }
    while (!mbar_try_wait(bar + 4, 0)) __nanosleep(20);  // the coefficient tile has landed
    __syncthreads();  // the only block-wide barrier: from here on warps synchronise through the counters

    auto wait_ge = [&](int* cnt, const int target) {  // whole warp, all lanes poll (a broadcast read), then reconverge
        while (*reinterpret_cast<volatile int*>(cnt) < target) __nanosleep(200);
        __threadfence_block();
        __syncwarp();
    };
    auto signal = [&](int* cnt) {  // whole warp: everything this warp wrote is visible before the count moves
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(cnt, 1);
        __syncwarp();
    };

    int gk = 0;  // group of the last item this warp took (items come in increasing order)
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(qhead, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= gstart[nt + 1]) break;
        while (item >= gstart[gk + 1]) ++gk;
        int sub = item - gstart[gk];
        // decode: kind 0 = input item I(k), 1 = x2 chunk U(k, sub), 2 = poly task P(k, sub)
        int kind, k;
        if (gk == 0) {
            const int ni = min(NX, nt);
            if (sub < ni) { kind = 0; k = sub; }
            else { kind = 1; k = 0; sub -= ni; }
        } else if (gk < nt) {
            const int ni = gk + NX - 1 < nt ? 1 : 0;
            if (sub < ni) { kind = 0; k = gk + NX - 1; }
            else if (sub - ni < nchunk[gk]) { kind = 1; k = gk; sub -= ni; }
            else { kind = 2; k = gk - 1; sub -= ni + nchunk[gk]; }
        } else {
            kind = 2;
            k = nt - 1;
        }
        const int t = t_first + k;
        const int xb = k % NX;

        if (kind == 2) {
            // ---- poly task: RN adjacent outputs of the period pattern, lanes = periods ----
            wait_ge(in_done + k, 1);
            wait_ge(up_done + k, nchunk[k]);
            if (!FUSED && (tflags[k] & 1)) {  // poly-only: the bulk part of the tile has landed
                const uint32_t par = (uint32_t)(tflags[k] >> 1) & 1u;
                while (!mbar_try_wait(bar + xb, par)) __nanosleep(20);
                __syncwarp();
            }
            const int FMk = FUSED ? FM : tg_npos[k];
            const T* __restrict__ vr = vs0 + (k % g.nv) * g.vlen;
            const int gi = sub * gpw + gsub;
            if (gi < g.G) {
                const int64_t n_lo = (int64_t)t * P * L;  // first output of the tile; its full_n is F0 past the tile origin
                const int64_t n_hi = min((int64_t)c.n_out, n_lo + (int64_t)P * L);
                const int iota0 = gi * RN;
                // window slot x of lane jl is sample w = jl*Mi + off + x, stored at FM + w + PAD*floor(w/Mi)
                const int off = goff[gi];
                const int c0 = off < 0 ? -1 : off / Mi;
                const int xc0 = (c0 + 1) * Mi - off, xc1 = xc0 + Mi;  // slots at which the walk enters the next period
                const T* __restrict__ sp = vr + FMk + jl * (Mi + PAD) + off + (PAD ? c0 : 0);
                const T* __restrict__ cp0 = cg + (size_t)gi * g.gpitch;  // [tap][RN], 16-byte aligned
                T W[WN], acc[RN];
#pragma unroll
                for (int x = 0; x < WN; ++x) W[x] = sp[x + (PAD ? (x >= xc0) + (x >= xc1) : 0)];
#pragma unroll
                for (int i = 0; i < RN; ++i) acc[i] = T(0);
                auto tap = [&](const int u, const int kk) {  // u = kk % WN (compile-time in the unrolled bodies)
                    T cf[RN];
#pragma unroll
                    for (int q = 0; q < RN / VEC; ++q)
                        vec_unpack(*reinterpret_cast<const V*>(cp0 + kk * RN + q * VEC), cf + q * VEC);
#pragma unroll
                    for (int i = 0; i < RN; ++i) acc[i] = fma(W[(u + i * S) % WN], cf[i], acc[i]);
                    const int x = kk + WN;
                    W[u] = sp[x + (PAD ? (x >= xc0) + (x >= xc1) : 0)];
                };
                int it0 = 0;
                for (; it0 + WN <= g.tp; it0 += WN) {  // branch-free bodies: the loads of the next taps overlap the FMAs
#pragma unroll
                    for (int u = 0; u < WN; ++u) tap(u, it0 + u);
                }
#pragma unroll
                for (int u = 0; u < WN; ++u)
                    if (it0 + u < g.tp) tap(u, it0 + u);
                const int64_t nb = n_lo + (int64_t)jl * L + iota0;
#pragma unroll
                for (int i = 0; i < RN; ++i)
                    if (iota0 + i < L && nb + i < n_hi) out[nb + i] = acc[i];
            }
            signal(po_done + k);
            continue;
        }

        const int n_pos = tg_npos[k], wend = tg_wend[k];
        const bool bulk = (tflags[k] & 1) != 0;
        T* __restrict__ xs = xs0 + xb * g.xlen;
        T* __restrict__ vw = vs0 + (k % g.nv) * g.vlen;

        if (kind == 0) {
            // ---- input item: stage the x2 input window of tile k (TMA bulk copy when regular) and the part of
            //      the tile that is the polyphase stage's carried tail ----
            const int kprev = k - NX;  // last user of this window buffer
            if (kprev >= 0) wait_ge(up_done + kprev, nchunk[kprev]);
            // The intermediate buffer is written here only by the poly-only tile load and by the carried-tail copy (first
            // tile of a row). A fused input item must NOT wait for it otherwise: P(k - nv) sits later in the queue, and the
            // prefetch of tile k would be held back until those tasks have run.
            const bool writes_tile = !FUSED || (int64_t)t * P * Mi < (int64_t)c.hp;
            if (writes_tile && k >= g.nv) wait_ge(po_done + (k - g.nv), n_wt);
            if (!FUSED) {
                // poly-only: the tile's samples come straight from the carried tail / the stage input (period j at pitch
                // Mi + PAD): one TMA bulk copy when the periods are contiguous, else asynchronous element copies
                const int64_t m0 = (int64_t)t * P * Mi;
                const int fm = tg_npos[k], nb = tg_words[k], wb = tg_woff[k];
                if (nb > 0 && lane == 0) {
                    const uint32_t bytes = (uint32_t)(nb * sizeof(T));
                    mbar_expect_tx(bar + xb, bytes);
                    bulk_g2s(vw + fm + wb, in + tg_src[k], bytes, bar + xb);
                }
                // elements outside the bulk range: [0, wb) and [wb + nb, wend)
                const int n_head = nb > 0 ? wb : wend, n_rest = nb > 0 ? wend - (wb + nb) : 0;
                for (int q = lane; q < n_head + n_rest; q += 32) {
                    const int w = q < n_head ? q : wb + nb + (q - n_head);
                    const int64_t d = m0 + w;
                    T* dst = vw + fm + w + (PAD ? w / Mi : 0);
                    if (d < c.hp) *dst = hist_p[d];
                    else cp_async_elem(dst, in + (d - c.hp));
                }
                cp_async_wait_all();
                signal(in_done + k);
                continue;
            }
            if (bulk) {
                if (lane == 0) {
                    const uint32_t bytes = (uint32_t)(tg_words[k] * sizeof(T));
                    mbar_expect_tx(bar + xb, bytes);
                    bulk_g2s(xs, in + tg_src[k], bytes, bar + xb);
                }
            } else {  // edge tile (touches the carried tail or the end of the row): guarded loads
                const int need = n_pos > 0 ? n_pos - 1 + c.t1 : 0;
                const int i_lo = tg_src[k] + c.hu;
                const int tot = c.hu + c.n_in;
#pragma unroll 8
                for (int i = lane; i < g.xlen; i += 32) {  // independent predicated loads: eight in flight per lane
                    const int gidx = i_lo + i;
                    const bool ok = i < need && gidx >= 0 && gidx < tot;
                    const T* __restrict__ src = gidx < c.hu ? hist_u + gidx : in + (gidx - c.hu);
                    xs[i] = ok ? *src : T(0);
                }
            }
            const int64_t m0 = (int64_t)t * P * Mi;
            for (int64_t d = m0 + lane; d < min((int64_t)c.hp, m0 + wend); d += 32) {
                const int w = (int)(d - m0);
                vw[FM + w + (PAD ? w / Mi : 0)] = hist_p[d];
            }
            signal(in_done + k);
            continue;
        }

        // ---- x2 chunk: 32 thread tasks of the register-tiled FIR core -> intermediate buffer ----
        wait_ge(in_done + k, 1);
        if (k >= g.nv) wait_ge(po_done + (k - g.nv), n_wt);
        if (bulk) {
            const uint32_t par = (uint32_t)(tflags[k] >> 1) & 1u;
            while (!mbar_try_wait(bar + xb, par)) __nanosleep(20);
            __syncwarp();  // lanes leave the poll loop at different times: reconverge before the FIR core
        }
        const int n_tasks = (n_pos + R - 1) / R;
        const int task = sub * 32 + lane;
        if (task < n_tasks) {
            const int woff = tg_woff[k];  // w of the first sample of the tile's first position
            T res[R][NF];
            fir_tile_accumulate<T, 1, NF, R>(xs + R * task, cs, g.cp, c.t1, 0, res);
            int w = woff + 2 * R * task;
            int j = (w + Mi) / Mi - 1, r = w - j * Mi;  // floor division (w >= -Mi)
            T* __restrict__ vp = vw + FM + w + (PAD ? j : 0);
#pragma unroll
            for (int q = 0; q < R; ++q)
#pragma unroll
                for (int h = 0; h < NF; ++h) {
                    if (w >= 0 && w < wend && R * task + q < n_pos) *vp = res[q][h];
                    ++w;
                    ++vp;
                    if (++r == Mi) {
                        r = 0;
                        if (PAD) ++vp;
                    }
                }
        }
        signal(up_done + k);
    }
}

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

template <int NTASK>
static bool launch_poly_rows_mma_t(const PolyCall& c, cudaStream_t s) {
    constexpr int TO = 8 * NTASK;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    RowsMmaGeom g{};
    const int omax = (int)std::ceil(7 * r) + 1;
    g.kp = ((omax + c.taps + 3) / 4) * 4;
    if (g.kp > 2 * c.taps + 8) return false;  // too many structural zeros: K3i
    g.span = (int)std::ceil((TO - 1) * r) + 1 + g.kp + 8;
    g.pitch = ((g.span + 2 + 15) / 16) * 16 + 4;  // rows 32 bytes apart modulo 128 (B fragment reads: two wavefronts)
    g.n_tiles = (c.n_out + TO - 1) / TO;
    const int n_rb = (c.n_streams + 31) / 32;
    g.nrb = 1;
    // measured on the batched 44.1k->48k chain (256 rows): 1 / 2 / 4 / 8 row blocks per coefficient evaluation -> 1.93 / 1.45 /
    // 1.19 / 1.10 ms for the polyphase stage
    static const int max_nrb = [] { const char* e = std::getenv("GAR_K3M_NRB"); return e ? std::atoi(e) : 8; }();
    while (g.nrb < max_nrb && g.nrb * 2 <= n_rb && (int64_t)g.n_tiles * ((n_rb + g.nrb * 2 - 1) / (g.nrb * 2)) >= 4 * 148) g.nrb *= 2;
    static const int force_nbuf = [] { const char* e = std::getenv("GAR_K3M_NBUF"); return e ? std::atoi(e) : 0; }();
    const size_t fixed = 16 + (size_t)NTASK * g.kp * 8 * sizeof(double) + (size_t)NTASK * 8 * 4 * sizeof(int);
    const size_t xbytes = (size_t)32 * g.pitch * sizeof(double);
    g.nbuf = force_nbuf ? force_nbuf : 1;
    const size_t smem = fixed + g.nbuf * xbytes;
    if (smem > (g.nbuf == 2 ? 227 : 113) * 1024) return false;
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

static bool launch_poly_rows_mma(const PolyCall& c, cudaStream_t s) {
    if (!g_fir_mma || c.n_streams < 8 || (int64_t)c.n_out * c.n_streams < 16384 || c.L > 4096 || c.taps > 1024) return false;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.0) || r > 8.0) return false;
    static const int ntask = [] { const char* e = std::getenv("GAR_K3M_NTASK"); return e ? std::atoi(e) : 8; }();
    if (ntask == 4) return launch_poly_rows_mma_t<4>(c, s) || launch_poly_rows_mma_t<8>(c, s);
    return launch_poly_rows_mma_t<8>(c, s) || launch_poly_rows_mma_t<4>(c, s);
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

// Fallback: one thread per output element, operands straight from global/L1.
template <typename T>
__global__ void __launch_bounds__(256) fir_generic_kernel(const FirCall c, const int n_tiles) {
    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    if (tile == n_tiles) {
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int64_t o = (int64_t)tile * 256 + threadIdx.x;
    if (o >= (int64_t)c.n_pos * c.nf) return;
    const int j = (int)(o / c.nf), p = (int)(o % c.nf);
    const int g = c.first + j * c.stride;
    const T* __restrict__ bank = static_cast<const T*>(c.bank) + (int64_t)p * c.taps;
    double tot = 0;
    T acc = 0;
    const int kc = (c.taps - 1) / 2;
    for (int k = 0; k < c.taps; ++k) {
        acc = fma(vload(hist, c.hist_len, in, c.n_in, g + k), bank[k], acc);
        if (sizeof(T) == 4 && ((k & 255) == 255 || (k >= kc && k < kc + 12 && ((k - kc) & 3) == 3))) {
            tot += (double)acc;
            acc = 0;
        }
    }
    (static_cast<T*>(c.out) + row * c.out_stride)[o] = (T)(tot + (double)acc);
}

// =============================================================================================
// Polyphase stage. One thread per output; the block's input span sits in shared memory.
// =============================================================================================
template <typename T, bool INTERP, int TO>
__global__ void __launch_bounds__(TO) poly_kernel(const PolyCall c, const int n_tiles, const int xcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);
    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    if (tile == n_tiles) {
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int n0 = tile * TO;
    const int n1 = min(n0 + TO, c.n_out) - 1;
    const int64_t L = c.L;
    const int div0 = (int)(((c.at0 + (int64_t)n0 * c.step) >> 16) / L);
    const int div1 = (int)(((c.at0 + (int64_t)n1 * c.step) >> 16) / L);
    const int span = div1 - div0 + c.taps;
    const bool staged = span <= xcap;
    if (staged) {
        for (int i = threadIdx.x; i < span; i += TO) xs[i] = vload(hist, c.hist_len, in, c.n_in, div0 + i);
        __syncthreads();
    }
    const int n = n0 + threadIdx.x;
    if (n >= c.n_out) return;
    const int64_t at = c.at0 + (int64_t)n * c.step;
    const int64_t full = at >> 16;
    const int div = (int)(full / L);
    const int phase = (int)(full - (int64_t)div * L);
    const T x = (T)(int)(at & 0xFFFF) * (T)(1.0 / 65536.0);
    const int64_t co = (int64_t)phase * c.taps;
    const T* __restrict__ ca = static_cast<const T*>(c.bank_a) + co;
    const T* __restrict__ cb = static_cast<const T*>(c.bank_b) + co;
    const T* __restrict__ cc = static_cast<const T*>(c.bank_c) + co;
    const T* __restrict__ cd = static_cast<const T*>(c.bank_d) + co;
    // products of two float32 are exact in float64, so the float32 path only rounds once, at the store (two
    // interleaved chains); float64 sums strictly in tap order, like every other float64 kernel here, so the
    // fused kernels reproduce the stand-alone launches bit for bit
    double acc0 = 0, acc1 = 0;
    const int base = div - div0;
    for (int k = 0; k < c.taps; ++k) {
        T coef = ca[k];
        if (INTERP) coef = fma(x, fma(x, fma(x, cd[k], cc[k]), cb[k]), coef);
        const T h = staged ? xs[base + k] : vload(hist, c.hist_len, in, c.n_in, div + k);
        if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h, (double)coef, acc1);
        else acc0 = fma((double)h, (double)coef, acc0);
    }
    (static_cast<T*>(c.out) + row * c.out_stride)[n] = (T)(acc0 + acc1);
}

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
__global__ void cast_kernel(const S* src, int64_t src_stride, D* dst, int64_t dst_stride, int n) {
    const int64_t row = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[row * dst_stride + i] = (D)src[row * src_stride + i];
}


// ---- N1: interleaved / integer-PCM boundary -------------------------------------------------------
template <typename TI, typename T>
__global__ void deinterleave_kernel(const TI* __restrict__ in, int channels, int64_t n_frames, T* __restrict__ planar,
                                    int64_t stride, double inv_max) {
    const int64_t total = n_frames * channels;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / channels;
        const int ch = (int)(idx - i * channels);
        const double v = (double)in[idx];
        planar[ch * stride + i] = inv_max != 0.0 ? (T)__dmul_rn(v, inv_max) : (T)v;  // main.go:444-470 deinterleaveInto
    }
}
template <typename T, typename TO>
__global__ void interleave_kernel(const T* __restrict__ planar, int64_t stride, int channels, int64_t n_frames,
                                  TO* __restrict__ out, double max_val) {
    const int64_t total = n_frames * channels;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / channels;
        const int ch = (int)(idx - i * channels);
        double v = (double)planar[ch * stride + i];
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

// ---- variant table for the tiled FIR -----------------------------------------------------------
struct FirVariant {
    int dtype, stride, nf, r;
    const char* name;
};

// tiles per block: long runs amortise the per-block filter load and hide the TMA prefetch under the FMAs,
// but keep >= ~3 blocks per resident slot in flight for balance
static int pick_tiles_per_block(int n_tiles, int n_streams, int* n_groups) {
    int dev = 0;
    cudaGetDevice(&dev);
    static int sm_count[64] = {0};
    if (!sm_count[dev & 63]) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    const int sms = sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148;
    const int64_t total_tiles = (int64_t)n_tiles * n_streams;
    int tpb = (int)(total_tiles / ((int64_t)sms * 4 * 3));
    tpb = tpb < 1 ? 1 : (tpb > 16 ? 16 : tpb);
    if (tpb > n_tiles) tpb = n_tiles;
    *n_groups = (n_tiles + tpb - 1) / tpb;
    return tpb;
}

template <typename T, int M, int NF, int R>
void launch_fir_tiled(const FirCall& c, cudaStream_t s) {
    constexpr int NT = 128;
    constexpr int VEC = VecOf<T>::N;
    constexpr int NCH = (M * (R - 1) + VEC - 1) / VEC + 1;
    constexpr int TJ = NT * R;
    const int cp = ((c.taps + VEC - 1 + VEC - 1) / VEC) * VEC;  // room for any alignment pad
    const int xlen = M * R * (NT - 1) + (cp / VEC + NCH + 1) * VEC;
    const size_t smem = 16 + (size_t)(NF * cp + 2 * xlen) * sizeof(T);
    const int n_tiles = (c.n_pos + TJ - 1) / TJ;
    int n_groups = 1;
    const int tpb = pick_tiles_per_block(n_tiles, c.n_streams, &n_groups);
    auto k = fir_tiled_kernel<T, M, NF, R, NT>;
    static size_t configured[64] = {0};  // per instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(n_groups + 1) * c.n_streams;
    k<<<(unsigned)blocks, NT, smem, s>>>(c, n_tiles, tpb, n_groups, cp, xlen);
    count_launch();
}

template <int M, int NF, int R>
void launch_fir_f32x2(const FirCall& c, cudaStream_t s) {
    constexpr int NT = 128;
    constexpr int MAXE = M * (R - 1) - ((M * (R - 1)) & 1);
    constexpr int NCH = (MAXE + 4 + 3) / 4;
    constexpr int NS = ((M & 1) && R > 1) ? 2 : 1;
    constexpr int TJ = NT * R;
    const int cp = ((c.taps + 3 + (NS - 1) + 3) / 4) * 4;
    const int xlen = M * R * (NT - 1) + (cp / 4 + NCH + 1) * 4;
    const size_t smem = 16 + (size_t)(NF * NS * cp + 2 * xlen) * sizeof(float);
    const int n_tiles = (c.n_pos + TJ - 1) / TJ;
    int n_groups = 1;
    const int tpb = pick_tiles_per_block(n_tiles, c.n_streams, &n_groups);
    int dev = 0;
    cudaGetDevice(&dev);
    auto k = fir_f32x2_kernel<M, NF, R, NT>;
    static size_t configured[64] = {0};
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(n_groups + 1) * c.n_streams;
    k<<<(unsigned)blocks, NT, smem, s>>>(c, n_tiles, tpb, n_groups, cp, xlen);
    count_launch();
}

}  // namespace

// float32 decimators on packed FMAs
#define GAR_FIR_X2_VARIANTS(X)          \
    X(3, 1, 12, "fir_f32x2_s3_r12")     \
    X(2, 1, 10, "fir_f32x2_s2_r10")     \
    X(4, 1, 7, "fir_f32x2_s4_r7")

#define GAR_FIR_VARIANTS(X)                 \
    X(float, DT_F32, 3, 1, 12, "fir_f32_s3_r12")  \
    X(float, DT_F32, 2, 1, 10, "fir_f32_s2_r10")  \
    X(float, DT_F32, 4, 1, 7, "fir_f32_s4_r7")    \
    X(float, DT_F32, 1, 2, 12, "fir_f32_up2_r12") \
    X(float, DT_F32, 1, 3, 4, "fir_f32_up3_r4")   \
    X(float, DT_F32, 1, 4, 4, "fir_f32_up4_r4")   \
    X(double, DT_F64, 2, 1, 7, "fir_f64_s2_r7")   \
    X(double, DT_F64, 3, 1, 6, "fir_f64_s3_r6")   \
    X(double, DT_F64, 4, 1, 5, "fir_f64_s4_r5")   \
    X(double, DT_F64, 1, 2, 6, "fir_f64_up2_r6")  \
    X(double, DT_F64, 1, 3, 2, "fir_f64_up3_r2")  \
    X(double, DT_F64, 1, 4, 2, "fir_f64_up4_r2")

void set_tensor_fir(bool on) { g_fir_mma = on; }

const char* fir_variant_name(int dtype, int stride, int nf, int taps, int64_t n_pos, int n_streams) {
    (void)taps; (void)n_pos; (void)n_streams;
#define X(M, NF, R, NAME) \
    if (dtype == DT_F32 && stride == M && nf == NF) return NAME;
    GAR_FIR_X2_VARIANTS(X)
#undef X
#define X(T, DT, M, NF, R, NAME) \
    if (dtype == DT && stride == M && nf == NF) return NAME;
    GAR_FIR_VARIANTS(X)
#undef X
    return dtype == DT_F32 ? "fir_f32_generic" : "fir_f64_generic";
}

const char* launch_fir(const FirCall& c, int dtype, cudaStream_t s) {
    if (c.n_streams <= 0) return "none";
    if (c.n_pos <= 0) {
        launch_carry(c.hist, c.hist_stride, c.hist_len, c.in, c.in_stride, c.n_in, c.hist_out, c.hist_out_stride, c.drop,
                     c.new_hist_len, c.n_streams, dtype, s);
        return "carry";
    }
    if (dtype == DT_F64)
        if (const char* nm = launch_fir_mma(c, s)) return nm;
#define X(M, NF, R, NAME)                                       \
    if (dtype == DT_F32 && c.stride == M && c.nf == NF) {       \
        launch_fir_f32x2<M, NF, R>(c, s);                       \
        return NAME;                                            \
    }
    GAR_FIR_X2_VARIANTS(X)
#undef X
#define X(T, DT, M, NF, R, NAME)                           \
    if (dtype == DT && c.stride == M && c.nf == NF) {      \
        launch_fir_tiled<T, M, NF, R>(c, s);               \
        return NAME;                                       \
    }
    GAR_FIR_VARIANTS(X)
#undef X
    const int64_t n_el = (int64_t)c.n_pos * c.nf;
    const int n_tiles = (int)((n_el + 255) / 256);
    const int64_t blocks = (int64_t)(n_tiles + 1) * c.n_streams;
    count_launch();
    if (dtype == DT_F32) {
        fir_generic_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(c, n_tiles);
        return "fir_f32_generic";
    }
    fir_generic_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(c, n_tiles);
    return "fir_f64_generic";
}

// A/B toggle (GAR_NO_RAT=1 or set_tiled_polyphase(false)): fall back to the one-thread-per-output kernels (no K4r / K3r / K3i)
static bool g_fused_rat = [] {
    const char* e = std::getenv("GAR_NO_RAT");
    return !(e && e[0] && e[0] != '0');
}();
void set_tiled_polyphase(bool on) { g_fused_rat = on; }

template <typename T, bool FUSED>
static bool launch_rat(const FusedCall& c, cudaStream_t s, RatCache* cache);

const char* launch_poly(const PolyCall& c, int dtype, cudaStream_t s, RatCache* cache) {
    if (c.n_streams <= 0) return "none";
    if (c.n_out <= 0) {
        launch_carry(c.hist, c.hist_stride, c.hist_len, c.in, c.in_stride, c.n_in, c.hist_out, c.hist_out_stride, c.drop,
                     c.new_hist_len, c.n_streams, dtype, s);
        return "carry";
    }
    // Batches of >= 8 lock-step rows: K3m, the polyphase stage on the FP64 tensor cores (any ratio)
    if (dtype == DT_F64 && g_fused_rat && launch_poly_rows_mma(c, s)) return c.interp ? "poly_rows_mma_f64_interp" : "poly_rows_mma_f64";
    // Batches of >= 8 rows with an even period length run K3i rather than K3r: K3r then stages its padded periods with
    // element copies from one warp (measured on the batched 48k->44.1k chain: 0.63 ms against 0.78 ms)
    if (dtype == DT_F64 && g_fused_rat && !c.interp && c.n_streams >= 8 && ((c.step >> 16) & 1) == 0 &&
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
        if (launch_rat<double, false>(f, s, cache)) return "poly_rat_f64";
    }
    if (dtype == DT_F64 && g_fused_rat && launch_poly_rows<double>(c, s)) return c.interp ? "poly_rows_f64_interp" : "poly_rows_f64";
    constexpr int TO = 128;
    const int n_tiles = (c.n_out + TO - 1) / TO;
    const int64_t blocks = (int64_t)(n_tiles + 1) * c.n_streams;
    const size_t esz = dtype == DT_F32 ? 4 : 8;
    // span of one tile: ((TO-1)*step >> 16)/L + 2 + taps, capped at 96 KB of shared memory
    int64_t span = (((int64_t)(TO - 1) * c.step) >> 16) / c.L + 2 + c.taps;
    const int64_t cap_words = (96 * 1024) / (int64_t)esz;
    int xcap = (int)(span < cap_words ? span : 0);  // 0 => read straight from global
    size_t smem = (size_t)xcap * esz;
#define LAUNCH(T, I)                                                                                          \
    {                                                                                                         \
        auto k = poly_kernel<T, I, TO>;                                                                       \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
        k<<<(unsigned)blocks, TO, smem, s>>>(c, n_tiles, xcap);                                               \
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


template <typename T, bool INTERP, int R>
static bool launch_fused_r(const FusedCall& c, cudaStream_t s) {
    constexpr int NT = 128;
    constexpr int VEC = VecOf<T>::N;
    constexpr int NCH = (R - 1 + VEC - 1) / VEC + 1;
    constexpr int MT = NT * R * 2;
    const int ms = (MT - (c.t2 - 1)) & ~1;  // tile stride in intermediate samples
    const int hpf = (c.hp + VEC - 1) / VEC * VEC;
    if (ms < MT / 2 || c.hp > 1024 || c.np <= 0) return false;
    const int n_mid = c.np * 2;
    const int n_tiles = (n_mid + ms - 1) / ms;
    const int cp = ((c.t1 + VEC - 1 + VEC - 1) / VEC) * VEC;
    const int xlen = R * (NT - 1) + (cp / VEC + NCH + 1) * VEC;
    size_t words = (size_t)2 * cp + xlen + hpf + MT;
    int bank_pitch = 0;
    // a-bank copy in shared memory: pays only when a block's set-up is amortised over many outputs. A streaming-size call
    // (a few small tiles) reads its coefficients through L1 instead: ncu showed 55 % of such a launch inside the copy.
    if (!INTERP && (int64_t)n_tiles * c.n_streams >= 2 * 148) {
        const int pitch = c.t2 | 1;
        if (((size_t)c.L * pitch + words) * sizeof(T) + 16 <= 100 * 1024) {
            bank_pitch = pitch;
            words += (size_t)c.L * pitch;
        }
    }
    const size_t smem = 16 + words * sizeof(T);
    auto k = fused_up2_poly_kernel<T, INTERP, R, NT>;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(n_tiles + 1) * c.n_streams;
    k<<<(unsigned)blocks, NT, smem, s>>>(c, n_tiles, ms, cp, xlen, hpf, bank_pitch);
    count_launch();
    return true;
}

// Tile size by call size: big tiles (R = 6 / 12 positions per thread) keep the FIR core FMA-bound; a streaming-size call
// has only a handful of them, so its critical path is one tile long — small tiles (R = 2 / 4) spread it over more SMs.
template <typename T, bool INTERP>
static bool launch_fused_t(const FusedCall& c, cudaStream_t s) {
    constexpr int RBIG = sizeof(T) == 8 ? 6 : 12, RSMALL = sizeof(T) == 8 ? 2 : 4;
    const int64_t big_tiles = ((int64_t)c.np * 2 + 128 * RBIG * 2 - 1) / (128 * RBIG * 2) * c.n_streams;
    if (big_tiles < 148 && c.t2 - 1 < 128 * RSMALL) return launch_fused_r<T, INTERP, RSMALL>(c, s);
    return launch_fused_r<T, INTERP, RBIG>(c, s);
}

// K4r / K3r launcher: picks the geometry; returns false when the call is not a rational-ratio case it covers.
// FUSED: x2 stage + polyphase stage (FusedCall as documented). !FUSED: polyphase stage alone; the call carries the
// stage in its polyphase fields and `in`/`n_in` = the stage input.
template <typename T, int S, int RN, int PAD, bool FUSED>
static bool launch_rat_t(const FusedCall& c, cudaStream_t s, RatCache* cache) {
    constexpr int VEC = VecOf<T>::N;
    constexpr int R = sizeof(T) == 8 ? 6 : 12;
    constexpr int NCH = (R - 1 + VEC - 1) / VEC + 1;
    constexpr int WN = (RN - 1) * S + 1;
    RatGeom g{};
    g.Mi = (int32_t)(c.step >> 16);
    const int Mi = g.Mi, L = c.L;
    g.D = (RN - 1) * S - (RN - 1) * Mi / L;  // worst lag of a static window slot behind the true offset
    g.tp = c.t2 + g.D;
    g.G = (L + RN - 1) / RN;
    g.cp = FUSED ? ((c.t1 + VEC - 1 + VEC - 1) / VEC) * VEC : 0;
    if (g.tp + 2 * WN > 2 * Mi) return false;  // the window walk may enter at most two further periods
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t div_last = ((c.at0 >> 16) + (int64_t)(c.n_out - 1) * Mi) / L;
    // one coefficient tile [tp][RN] per output group; with two groups per warp (P = 16) the two broadcast reads of a
    // tap must fall into different banks: tile pitch = 64 bytes mod 128
    int gpitch = g.tp * RN;
    while ((gpitch * (int)sizeof(T)) % 128 != 64) gpitch += VEC;
    g.gpitch = gpitch;
    const size_t tile_bytes = (size_t)g.G * gpitch * sizeof(T) + (((size_t)g.G * sizeof(int) + 15) & ~(size_t)15);
    // configurations in order of preference: two 256-thread blocks per SM, else one 512-thread block
    struct Opt { int P, nt, xbufs, nv; size_t lim; };
    Opt opts[4] = {{16, 256, 2, 2, 113 * 1024}, {32, 512, 2, 2, 227 * 1024}, {16, 512, 2, 2, 227 * 1024},
                   {16, 512, 1, 2, 227 * 1024}};
    if (!FUSED) {
        opts[0] = Opt{16, 256, 0, 3, 113 * 1024};
        opts[1] = Opt{16, 256, 0, 2, 113 * 1024};
        opts[2] = Opt{16, 512, 0, 3, 227 * 1024};
        opts[3] = Opt{16, 512, 0, 2, 227 * 1024};
    }
    static const int* forced = [] {  // tuning override: GAR_RAT_OPT="P,threads,xbufs,nv"
        static int v[4];
        const char* e = std::getenv("GAR_RAT_OPT");
        return (e && std::sscanf(e, "%d,%d,%d,%d", v, v + 1, v + 2, v + 3) == 4) ? v : (const int*)nullptr;
    }();
    if (forced) opts[0] = Opt{forced[0], forced[1], FUSED ? forced[2] : 0, forced[3], 227 * 1024};
    size_t smem = 0;
    int nthreads = 0;
    for (const Opt& o : opts) {
        int xlen = 0;
        if (FUSED) {
            const int n_pos_max = (o.P * Mi + c.t2) / 2 + 2 + VEC;
            const int tasks = (n_pos_max + R - 1) / R;
            xlen = R * (tasks - 1) + (g.cp / VEC + NCH + 1) * VEC;
        }
        const int vlen = (((g.D + 2) + o.P * (Mi + PAD) + g.tp + 2 * WN + 4) + 1) & ~1;
        const size_t need = 48 + RAT_CTL_BYTES +
                            ((size_t)2 * g.cp + (size_t)o.xbufs * xlen + (size_t)o.nv * vlen) * sizeof(T) + tile_bytes;
        if (need <= o.lim) {
            g.P = o.P;
            g.xlen = xlen;
            g.xbufs = o.xbufs;
            g.nv = o.nv;
            g.vlen = vlen;
            smem = need;
            nthreads = o.nt;
            break;
        }
    }
    if (!nthreads) return false;
    g.n_tiles = (int32_t)(div_last / ((int64_t)g.P * Mi)) + 1;
    {  // tiles per block, measured on the batched 44.1k<->48k chains: ~12 (fused) / ~6 (poly-only) blocks per resident
        // slot, at least 3 / 6 tiles (set-up amortisation) unless that would leave resident slots empty; longer blocks
        // lose more to the drain of their last tile than they save in set-up
        static int sms = 0;
        if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t slots = (int64_t)(sms > 0 ? sms : 148) * 2;
        const int64_t total_tiles = (int64_t)g.n_tiles * c.n_streams;
        int64_t tpb = total_tiles / (slots * (FUSED ? 12 : 6));
        tpb = std::max<int64_t>(tpb, FUSED ? 3 : 6);
        tpb = std::min<int64_t>(tpb, std::max<int64_t>(1, total_tiles / slots));
        tpb = std::min<int64_t>(std::min<int64_t>(tpb, RAT_MAXT), g.n_tiles);
        g.tiles_per_block = (int32_t)tpb;
        g.n_groups = (g.n_tiles + g.tiles_per_block - 1) / g.tiles_per_block;
    }
    if (const char* e = std::getenv("GAR_RAT_TPB")) {  // tuning override
        const int v = std::atoi(e);
        if (v >= 1 && v <= RAT_MAXT) {
            g.tiles_per_block = v < g.n_tiles ? v : g.n_tiles;
            g.n_groups = (g.n_tiles + g.tiles_per_block - 1) / g.tiles_per_block;
        }
    }
    {  // coefficient tile of this call's start phase: built on first use, cached per polyphase stage
        const int F0 = (int)(c.at0 >> 16);
        const int key = (int)sizeof(T) * 1000003 + S * 100003 + RN * 10007 + g.tp * 131 + g.gpitch * 7 + L;
        if (!cache->dev || cache->tile_bytes != tile_bytes || cache->key != key || (int)cache->built.size() != L) {
            if (cache->dev) {
                cudaStreamSynchronize(s);
                cudaFree(cache->dev);
                cache->dev = nullptr;
            }
            if (cudaMalloc(&cache->dev, tile_bytes * (size_t)L) != cudaSuccess) {
                cache->dev = nullptr;
                cudaGetLastError();
                return false;
            }
            cache->tile_bytes = tile_bytes;
            cache->key = key;
            cache->built.assign((size_t)L, 0);
        }
        T* tile = reinterpret_cast<T*>(static_cast<char*>(cache->dev) + (size_t)F0 * tile_bytes);
        if (!cache->built[(size_t)F0]) {
            rat_build_tile_kernel<T, S, RN><<<1, 256, (size_t)g.G * RN * sizeof(int), s>>>(
                static_cast<const T*>(c.bank_a), c.t2, L, Mi, F0, g.G, g.tp, g.gpitch, tile);
            count_launch();
            cache->built[(size_t)F0] = 1;
        }
        g.cg_src = tile;
        g.cg_bytes = (uint32_t)tile_bytes;
    }
    auto k = fused_up2_rat_kernel<T, S, RN, PAD, FUSED>;
    static size_t configured[64] = {0};
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(g.n_groups + 1) * c.n_streams;
    k<<<(unsigned)blocks, nthreads, smem, s>>>(c, g);
    count_launch();
    return true;
}

template <typename T, bool FUSED>
static bool launch_rat(const FusedCall& c, cudaStream_t s, RatCache* cache) {
    if (!cache || !g_fused_rat || c.interp || ((c.step | c.at0) & 0xFFFF) != 0 || c.n_out <= 0) return false;
    // small calls (streaming chunks) are latency-bound: the simpler kernels have far less per-block set-up
    if ((int64_t)c.n_out * c.n_streams < 65536) return false;
    if (FUSED && c.np <= 0) return false;
    const int64_t Mi = c.step >> 16;
    if (Mi <= c.L || Mi > 4 * (int64_t)c.L || Mi > 4096 || c.t2 > 512 || c.L > 255) return false;
    const int S = (int)((Mi + c.L - 1) / c.L);
    const bool pad = (Mi & 1) == 0;  // even period length: pad the period pitch to keep the lanes on distinct banks
    switch (S) {
        case 2: return pad ? launch_rat_t<T, 2, 8, 1, FUSED>(c, s, cache) : launch_rat_t<T, 2, 8, 0, FUSED>(c, s, cache);
        case 3: return pad ? launch_rat_t<T, 3, 6, 1, FUSED>(c, s, cache) : launch_rat_t<T, 3, 6, 0, FUSED>(c, s, cache);
        case 4: return pad ? launch_rat_t<T, 4, 6, 1, FUSED>(c, s, cache) : launch_rat_t<T, 4, 6, 0, FUSED>(c, s, cache);
    }
    return false;
}

const char* launch_fused_up2_poly(const FusedCall& c, int dtype, cudaStream_t s, RatCache* cache) {
    if (c.n_streams <= 0) return "none";
    if (dtype == DT_F32) {
        if (c.interp) return launch_fused_t<float, true>(c, s) ? "fused_up2_poly_f32_interp" : nullptr;
        return launch_fused_t<float, false>(c, s) ? "fused_up2_poly_f32" : nullptr;
    }
    // Batches of >= 8 lock-step rows run the x2 stage on the FP64 tensor cores (K1m) and the polyphase stage as its own
    // launch (K3r / K3i): faster than the fused vector-FMA kernel (measured: 0.53 against 0.55 ms on the batched
    // 44.1k->48k chain, 0.63 against 0.70 ms on 48k->44.1k); the fused kernel serves 1-7 rows.
    if (g_fir_mma && g_fused_rat && c.n_streams >= 8 && (int64_t)c.np * c.n_streams >= 32768) return nullptr;
    if (!c.interp && launch_rat<double, true>(c, s, cache)) return "fused_up2_rat_f64";
    // a large lock-step batch that the rational kernel does not cover runs as two launches: the stand-alone x2 kernel and
    // K3i (lanes = rows, interpolated coefficients evaluated once per batch) beat the one-thread-per-output fused kernel
    {
        const double r = (double)c.step / ((double)c.L * 65536.0);
        if (g_fused_rat && c.n_streams >= 8 && (int64_t)c.n_out * c.n_streams >= 16384 && r > 0.0 && r <= 8.0 &&
            c.L <= 4096 && c.t2 <= 1024)
            return nullptr;
    }
    if (c.interp) return launch_fused_t<double, true>(c, s) ? "fused_up2_poly_f64_interp" : nullptr;
    return launch_fused_t<double, false>(c, s) ? "fused_up2_poly_f64" : nullptr;
}

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
    dim3 grid((unsigned)((n + 1023) / 1024 < 4096 ? (n + 1023) / 1024 : 4096), (unsigned)n_rows);
    count_launch();
    if (src_dtype == DT_F32 && dst_dtype == DT_F64)
        cast_kernel<float, double><<<grid, 256, 0, s>>>((const float*)src, src_stride, (double*)dst, dst_stride, n);
    else if (src_dtype == DT_F64 && dst_dtype == DT_F32)
        cast_kernel<double, float><<<grid, 256, 0, s>>>((const double*)src, src_stride, (float*)dst, dst_stride, n);
    else if (src_dtype == DT_F32)
        cast_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, src_stride, (float*)dst, dst_stride, n);
    else
        cast_kernel<double, double><<<grid, 256, 0, s>>>((const double*)src, src_stride, (double*)dst, dst_stride, n);
}

namespace {
inline unsigned grid_for(int64_t total) {
    int64_t g = (total + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}
template <typename TI>
void deint_dispatch(const void* in, int channels, int64_t n, void* planar, int64_t stride, int dtype, double inv,
                    cudaStream_t s) {
    const unsigned g = grid_for(n * channels);
    if (dtype == DT_F32) deinterleave_kernel<TI, float><<<g, 256, 0, s>>>((const TI*)in, channels, n, (float*)planar, stride, inv);
    else deinterleave_kernel<TI, double><<<g, 256, 0, s>>>((const TI*)in, channels, n, (double*)planar, stride, inv);
}
template <typename TO>
void int_dispatch(const void* planar, int64_t stride, int dtype, int channels, int64_t n, void* out, double maxv,
                  cudaStream_t s) {
    const unsigned g = grid_for(n * channels);
    if (dtype == DT_F32) interleave_kernel<float, TO><<<g, 256, 0, s>>>((const float*)planar, stride, channels, n, (TO*)out, maxv);
    else interleave_kernel<double, TO><<<g, 256, 0, s>>>((const double*)planar, stride, channels, n, (TO*)out, maxv);
}
}  // namespace

void launch_deinterleave(const void* in, int fmt, int channels, int64_t n_frames, void* planar, int64_t stride,
                         int dtype, double inv_max, cudaStream_t s) {
    if (n_frames <= 0 || channels <= 0) return;
    count_launch();
    switch (fmt) {
        case 0: deint_dispatch<double>(in, channels, n_frames, planar, stride, dtype, 0.0, s); break;
        case 1: deint_dispatch<float>(in, channels, n_frames, planar, stride, dtype, 0.0, s); break;
        case 2: deint_dispatch<int16_t>(in, channels, n_frames, planar, stride, dtype, inv_max, s); break;
        case 3: deint_dispatch<int32_t>(in, channels, n_frames, planar, stride, dtype, inv_max, s); break;
        default: deint_dispatch<long long>(in, channels, n_frames, planar, stride, dtype, inv_max, s); break;
    }
}

void launch_interleave(const void* planar, int64_t stride, int dtype, int channels, int64_t n_frames, void* out, int fmt,
                       double max_val, cudaStream_t s) {
    if (n_frames <= 0 || channels <= 0) return;
    count_launch();
    switch (fmt) {
        case 0: int_dispatch<double>(planar, stride, dtype, channels, n_frames, out, 0.0, s); break;
        case 1: int_dispatch<float>(planar, stride, dtype, channels, n_frames, out, 0.0, s); break;
        case 2: int_dispatch<int16_t>(planar, stride, dtype, channels, n_frames, out, max_val, s); break;
        case 3: int_dispatch<int32_t>(planar, stride, dtype, channels, n_frames, out, max_val, s); break;
        default: int_dispatch<long long>(planar, stride, dtype, channels, n_frames, out, max_val, s); break;
    }
}

long long launch_count(bool reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
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
        if (dtype == 2)  // packed f32x2: 4 reps x 8 chains x 2 lanes = 64 FMA per iteration, same as the others
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
    *flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    return ms;
}

}  // namespace gar
