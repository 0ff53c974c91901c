// kernels_fir.cu — integer-factor FIR stages (dft_stage.go): K1/K2 register-tiled vector kernels (fir_tiled_kernel,
// fir_f32x2_kernel with packed fma.rn.f32x2), K1m/K2m on the FP64 tensor cores (fir_mma_f64_kernel, DMMA), the generic
// fallback, and launch_fir.
#include "mma_cores.cuh"
#include <type_traits>

namespace gar {
namespace {

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
                                                       const int xlen, const int uc_a /*>= 0: taps from PU, pad = uc_a*/,
                                                       const __grid_constant__ FirTapsD<(sizeof(T) == 8 ? NF : 1), (sizeof(T) == 8 ? FIR_TAPS_CPD : 2)> PU) {
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

    pdl_trigger_if_small();
    if (group == n_groups) {  // carry block
        pdl_wait();
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop,
                  c.new_hist_len);
        return;
    }
    const int t_first = group * tiles_per_block;
    const int nt = min(tiles_per_block, n_tiles - t_first);

    // leading pad so that every tile's bulk source address is 16-byte aligned (constant along the row)
    // (uc_a >= 0: the pad is the same for every row and comes in as a kernel parameter, which keeps the constant-bank tap
    //  addressing of fir_tile_accumulate_uc provably warp-uniform)
    const int a = uc_a >= 0 ? uc_a
                            : (int)(((reinterpret_cast<uintptr_t>(in) / sizeof(T)) + (uintptr_t)(int64_t)(c.first - c.hist_len)) &
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
        block_copy4(NF * cp, [&](int i) {
            const int p = i / cp, k = i % cp - a;
            return (k >= 0 && k < c.taps) ? bank[p * c.taps + k] : T(0);
        }, [&](int i, T v) { cs[i] = v; });
    }
    pdl_wait();  // everything above touched shared memory and the constant bank only
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
            block_copy4(xlen, [&](int i) { return i < words ? vload(hist, c.hist_len, in, c.n_in, g0a + i) : T(0); },
                        [&](int i, T v) { xs[i] = v; });
            __syncthreads();
        }

        // ---- register-tiled sliding-window FIR ----
        T res[R][NF];
        if constexpr (sizeof(T) == 8) {
            if (uc_a >= 0) fir_tile_accumulate_uc<M, NF, R>(xs + M * R * tid, PU, c.taps, a, res);
            else fir_tile_accumulate<T, M, NF, R>(xs + M * R * tid, cs, cp, c.taps, a, res);
        } else {
            fir_tile_accumulate<T, M, NF, R>(xs + M * R * tid, cs, cp, c.taps, a, res);
        }

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
// Fewer than 8 rows: the columns are TIME SEGMENTS of the rows (virtual stream = row x segment of `ps` positions), so a
// single long stream fills the 8 columns with 8 of its own segments — A does not depend on the position.
// A block = 8 (virtual) streams x NW*MT*JT positions; A fragments are laid out per k-step in shared memory in fragment order.
// Taps are grouped in fours in window order, so results differ from the strictly sequential vector kernels in the last
// bits (1e-16 relative); identical call sequences are bit-identical.
// =============================================================================================
struct MmaGeom {
    int32_t nk, xlen, pitch, nbuf, n_tiles, tiles_per_block, n_groups, n_sg;  // k-steps, staged samples per stream, row pitch, window buffers
    int32_t nseg, ps, nvs;  // time segments per row, positions per segment, virtual streams = rows x segments (MMA columns)
    int32_t blen;           // zero-padded filter length of the A-fragment gather
    int32_t nkc, nqc;       // window chunks per tile along the taps (long filters), k-steps per chunk (multiple of the A window)
};

// IN32: `in` holds float32 samples (float64 pipelines behind the float32 API, constant.go:161-199; float32 engines computing in
// float64): the window is bulk-copied as float32 into the upper half of its row buffer and widened in place by all threads
// (read - barrier - write), so the stand-alone cast launch and its 12 bytes per sample of HBM traffic disappear.
template <int M, int NF, int NW, int MT, bool KC, bool IN32 = false>
__global__ void __launch_bounds__(NW * 32, (NW == 8 && M == 1 && !KC) ? 4 : (NW == 8 ? 2 : 1)) fir_mma_f64_kernel(const FirCall c, const MmaGeom g) {
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
    const int blen = g.blen;                                     // padded filter length, even: windows stay 16-byte aligned
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);       // [2] mbarriers of the bulk-copied window buffers
    double* Bs = reinterpret_cast<double*>(smem_raw + 16);       // [NF][blen] zero-padded bank
    double* Xs0 = Bs + (size_t)NF * blen;                        // [nbuf][8][pitch] sample windows of the block's streams

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_work = g.n_sg * g.n_groups;
    if ((int)blockIdx.x >= n_work) {  // carried tail of one row
        const int64_t row = (int)blockIdx.x - n_work;
        if (IN32) {
            const double* __restrict__ hist = static_cast<const double*>(c.hist) + row * c.hist_stride;
            const float* __restrict__ in32 = static_cast<const float*>(c.in) + row * c.in_stride;
            double* __restrict__ ho = static_cast<double*>(c.hist_out) + row * c.hist_out_stride;
            block_copy4(c.new_hist_len,
                        [&](int i) {
                            const int v = c.drop + i;
                            return v < c.hist_len ? hist[v] : (v - c.hist_len < c.n_in ? (double)in32[v - c.hist_len] : 0.0);
                        },
                        [&](int i, double v) { ho[i] = v; });
            return;
        }
        carry_row(static_cast<const double*>(c.hist) + row * c.hist_stride, c.hist_len,
                  static_cast<const double*>(c.in) + row * c.in_stride, c.n_in,
                  static_cast<double*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const int grp = blockIdx.x % g.n_groups;
    const int sbase = (blockIdx.x / g.n_groups) * 8;  // first virtual stream (row x segment) of the block's 8 MMA columns
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
    // TMA bulk copies need 16-byte aligned sources: all 8 columns share the alignment when the row stride and the segment
    // length in samples are even
    constexpr int AL = IN32 ? 4 : 2;  // samples per 16 bytes of the source
    const bool rows_bulk = (c.in_stride & (AL - 1)) == 0 && (((int64_t)g.ps * M) & (AL - 1)) == 0 && sbase + 8 <= g.nvs;
    const int xbuf = 8 * g.pitch;
    // column r of the block: row, first position of its segment — fixed for the block, tabulated once (no divisions per tile)
    __shared__ int s_col_row[8], s_col_pos0[8];
    if (tid < 8) {
        s_col_row[tid] = (sbase + tid) / g.nseg;
        s_col_pos0[tid] = ((sbase + tid) % g.nseg) * g.ps;
    }
    __syncthreads();
    auto col_row = [&](const int r) -> int64_t { return s_col_row[r]; };
    auto col_pos0 = [&](const int r) -> int64_t { return s_col_pos0[r]; };
    int pos0_min = INT32_MAX, pos0_max = 0;
    for (int r = 0; r < 8; ++r) {
        pos0_min = min(pos0_min, s_col_pos0[r]);
        pos0_max = max(pos0_max, s_col_pos0[r]);
    }
    const size_t isz = IN32 ? sizeof(float) : sizeof(double);
    const char* __restrict__ col0_in = static_cast<const char*>(c.in) + (col_row(0) * c.in_stride + col_pos0(0) * M) * isz;

    // geometry of iteration `it` = (local tile kt, tap chunk kc): segment-local first position jb0, first k-step qa of the
    // chunk, staged length (from window sample 4*qa on), bulk-copy parameters
    auto tile_geom = [&](const int it, int& jb0, int& qa, int& len, int& a, int& wlen) -> bool {
        const int kt = KC ? it / g.nkc : it, kc = KC ? it - kt * g.nkc : 0;  // KC = false: one chunk, the whole window
        jb0 = (t_first + kt) * TJ;
        qa = KC ? kc * g.nqc : 0;
        const int npos_t = min(TJ, g.ps - jb0);
        // the MMA tiles that hold a valid position read window samples below this bound
        const int len_full = ((npos_t + JT - 1) / JT * JT - 1) * M + 4 * g.nk + 4;
        len = max(0, min(g.xlen, len_full - 4 * qa));
        a = 0;
        wlen = 0;
        if (!rows_bulk || len == 0) return false;
        // index into `in` of every column's window start: all must lie inside the rows
        const int64_t g0 = (int64_t)c.first + (int64_t)jb0 * M - c.hist_len + 4 * qa;
        const int64_t gmin = g0 + (int64_t)pos0_min * M, gmax = g0 + (int64_t)pos0_max * M;
        if (gmin < 0) return false;
        const char* src0 = col0_in + g0 * (int64_t)isz;
        a = (int)((reinterpret_cast<uintptr_t>(src0) & 15u) / isz);  // start `a` samples early: aligned sources
        wlen = (len + a + AL - 1) & ~(AL - 1);
        if (gmin - a >= 0 && gmax - a + wlen <= c.n_in && wlen <= g.pitch) return true;
        a = 0;  // element copies start exactly at the window
        return false;
    };
    auto issue = [&](const int it, const int buf) {  // one thread
        int jb0, qa, len, a, wlen;
        if (!tile_geom(it, jb0, qa, len, a, wlen)) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar + buf, (uint32_t)(8 * wlen * isz));
        for (int r = 0; r < 8; ++r) {
            const int64_t gi = (int64_t)c.first + (col_pos0(r) + jb0) * M - c.hist_len + 4 * qa - a;
            // float32 windows land in the upper half of the row buffer (float index pitch ..), widened in place after the wait
            void* dst = IN32 ? static_cast<void*>(reinterpret_cast<float*>(Xs0 + buf * xbuf + r * g.pitch) + g.pitch)
                             : static_cast<void*>(Xs0 + buf * xbuf + r * g.pitch);
            bulk_g2s(dst, static_cast<const char*>(c.in) + (col_row(r) * c.in_stride + gi) * (int64_t)isz, (uint32_t)(wlen * isz), bar + buf);
        }
    };

    const int n_it = KC ? nt * g.nkc : nt;
    double acc[MT][2];
    double Areg[WA];
    for (int it = 0; it < n_it; ++it) {
        const int buf = g.nbuf == 2 ? (it & 1) : 0;
        double* __restrict__ Xs = Xs0 + buf * xbuf;
        int jb0, qa, len, a, wlen;
        const bool bulk = tile_geom(it, jb0, qa, len, a, wlen);
        // everyone is done with the windows this iteration overwrites (first tile: the bank and the mbarriers are set)
        __syncthreads();
        if (tid == 0) {
            if (g.nbuf == 2) {  // prefetch the next tile / chunk under this one's MMAs
                if (it == 0) issue(0, 0);
                if (it + 1 < n_it) issue(it + 1, buf ^ 1);
            } else {
                issue(it, 0);
            }
        }
        if (bulk) {
            const uint32_t ph = buf ? ph1 : ph0;
            while (!mbar_try_wait(bar + buf, ph)) {
            }
            if (buf) ph1 ^= 1u;
            else ph0 ^= 1u;
            if (IN32) {  // widen the 8 float32 windows in place: every thread reads its share, barrier, writes
                constexpr int PER = 16;  // elements per thread and round (8 * wlen <= 8 * pitch; two rounds cover any window)
                const int tot = 8 * wlen;
                for (int base = 0; base < tot; base += NT * PER) {
                    float v[PER];
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int e = base + u * NT + tid;
                        if (e < tot) {
                            const int r = e / wlen, i = e - r * wlen;
                            v[u] = reinterpret_cast<const float*>(Xs + r * g.pitch)[g.pitch + i];
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int e = base + u * NT + tid;
                        if (e < tot) {
                            const int r = e / wlen, i = e - r * wlen;
                            Xs[r * g.pitch + i] = (double)v[u];
                        }
                    }
                    __syncthreads();
                }
            }
        } else {  // edge tile (a column touches the carried tail or the end of its row): element copies
            for (int r = warp; r < 8; r += NW) {
                double* __restrict__ dst = Xs + r * g.pitch;
                if (sbase + r >= g.nvs) {
                    for (int i = lane; i < len; i += 32) dst[i] = 0.0;
                    continue;
                }
                const int64_t row = col_row(r);
                const int64_t v0 = (int64_t)c.first + (col_pos0(r) + jb0) * M + 4 * qa;  // element i of the window is v[v0 + i]
                const int i1 = (int)min((int64_t)len, max((int64_t)0, (int64_t)c.hist_len - v0));
                const int i2 = (int)min((int64_t)len, max((int64_t)i1, total - v0));
                const double* __restrict__ hsrc = static_cast<const double*>(c.hist) + row * c.hist_stride + v0;
                for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
                if (IN32) {
                    const float* __restrict__ isrc32 = static_cast<const float*>(c.in) + row * c.in_stride + (v0 - c.hist_len);
#pragma unroll 4
                    for (int i = i1 + lane; i < i2; i += 32) dst[i] = (double)isrc32[i];
                } else {
                    const double* __restrict__ isrc = static_cast<const double*>(c.in) + row * c.in_stride + (v0 - c.hist_len);
#pragma unroll 4
                    for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
                }
                for (int i = i2 + lane; i < len; i += 32) dst[i] = 0.0;
            }
            cp_async_wait_all();
            __syncthreads();
        }

        // ---- MT MMA tiles per warp; step q loads window chunk Q = warp*MT*SH + q and A fragment q ----
        const int npos_t = min(TJ, g.ps - jb0);
        if (warp * MT * JT < npos_t) {
            if (!KC || qa == 0) {
#pragma unroll
                for (int b = 0; b < MT; ++b) acc[b][0] = acc[b][1] = 0.0;
            }
            const int qb = KC ? min(nq, qa + g.nqc) : nq;
            // B fragment: lane l reads X[w = 4*Q + l%4][column l/4]; the staged window starts at sample 4*qa of the tile's
            const double* __restrict__ xw = Xs + (lane >> 2) * g.pitch + (lane & 3) + a + 4 * (warp * MT * SH) - 4 * qa;
            // A fragment: lane l holds A[row = l/4][w = 4*kk + l%4] = bank[p][4*kk + l%4 - jj*M], row = jj*NF + p
            const double* __restrict__ aw = Bs + ((lane >> 2) % NF) * blen + BOFF + (lane & 3) - ((lane >> 2) / NF) * M;
            // WA steps per group: steady-state groups (all MT tiles inside the filter) are straight-line code, the first group and
            // the ramp-down groups use compile-time tile masks (mma_cores.cuh, shared with the chain kernel K5)
            fir_mma_warp_tiles<MT, SH>(acc, Areg, xw, aw, g.nk, nq, qa, qb);
            if (!KC || qb == nq) {
                // ---- D[row = lane/4][cols 2*(lane%4), +1]: output (pos0 + jb)*NF + row of the columns' rows ----
                const int r8 = lane >> 2;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 2 * (lane & 3) + e;
                    if (sbase + col >= g.nvs) continue;
                    const int64_t pos0 = col_pos0(col);
                    const int64_t plim = min((int64_t)c.n_pos, pos0 + g.ps);  // the column's segment ends here
                    if (c.out_f32) {  // float32(v) on the way out (constant.go:195-197): no cast launch behind this one
                        float* __restrict__ orow = static_cast<float*>(c.out) + col_row(col) * c.out_stride;
#pragma unroll
                        for (int b = 0; b < MT; ++b) {
                            const int64_t jb = pos0 + jb0 + (warp * MT + b) * JT;
                            if (jb + r8 / NF < plim) orow[jb * NF + r8] = (float)acc[b][e];
                        }
                        continue;
                    }
                    double* __restrict__ orow = static_cast<double*>(c.out) + col_row(col) * c.out_stride;
#pragma unroll
                    for (int b = 0; b < MT; ++b) {
                        const int64_t jb = pos0 + jb0 + (warp * MT + b) * JT;
                        if (jb + r8 / NF < plim) orow[jb * NF + r8] = acc[b][e];
                    }
                }
            }
        }
    }
}

template <int M, int NF>
static bool launch_fir_mma_t(const FirCall& c, cudaStream_t s, const bool dry = false) {
    constexpr int JT = 8 / NF, SH = JT * M / 4;
    int dev = 0;
    cudaGetDevice(&dev);
    MmaGeom g{};
    const int kp = c.taps + (JT - 1) * M;
    g.nk = (kp + 3) / 4;
    // fewer than 8 rows: split every row into time segments, the MMA columns are (row, segment) pairs
    g.nseg = 1;
    if (c.n_streams < 8) {
        static const int tbl[8] = {0, 8, 4, 8, 2, 8, 4, 8};  // rows x segments = 8, 8, 24, 8, 40, 24, 56
        g.nseg = tbl[c.n_streams];
    }
    g.nvs = c.n_streams * g.nseg;
    g.n_sg = (g.nvs + 7) / 8;
    // Zero-padded filter length. Two phase filters (x2 up-sampler): a half-warp's A-fragment LDS.64 reads a few doubles around
    // the same offset of BOTH filters; 8 doubles apart modulo the 16 eight-byte banks (length = 8 mod 16) the two groups do
    // not collide (ncu: the A loads took 4 wavefronts instead of 2 with the unpadded length)
    g.blen = (4 * g.nk + (JT - 1) * M + 5) & ~1;
    static const bool pad_bank = [] { const char* e = gar::tune_env("GAR_MMA_BANKPAD"); return !e || e[0] != '0'; }();
    if (NF == 2 && pad_bank) g.blen = ((g.blen + 7) & ~15) + 8;
    const size_t bank_bytes = 16 + (size_t)NF * g.blen * sizeof(double);
    // nkc > 1: tap-chunked staging (KC kernels); max_smem: give up when a block needs more (0: whatever fits an SM)
    auto run = [&](auto kernel, const int NW, const int MT, const int slot, const int nkc, const size_t max_smem) -> bool {
        const int TJ = NW * MT * JT;
        // long filters: the window of a tile is mostly taps; it is staged in `nkc` chunks along the taps (accumulators and
        // the rotating A window stay in registers), so that smaller / double-buffered stages fit
        const int WA = (MT - 1) * SH + 1, nq = g.nk + (MT - 1) * SH;
        g.nkc = nkc;
        g.nqc = ((nq + g.nkc - 1) / g.nkc + WA - 1) / WA * WA;
        g.nkc = (nq + g.nqc - 1) / g.nqc;
        g.xlen = g.nkc == 1 ? (TJ - 1) * M + 4 * g.nk + 4 * (MT - 1) * SH + 10 : 4 * (NW - 1) * MT * SH + 4 * g.nqc + 10;
        g.pitch = ((g.xlen + 15) / 16) * 16 + 4;  // rows 32 bytes apart modulo 128: the 8 x 32-byte B fragment reads tile two wavefronts
        const size_t xbytes = (size_t)8 * g.pitch * sizeof(double);
        // two window buffers (the next tile is prefetched under the MMAs) when at least two such blocks fit an SM
        static const int force_nbuf = [] { const char* e = gar::tune_env("GAR_MMA_NBUF"); return e ? std::atoi(e) : 0; }();
        g.nbuf = bank_bytes + 2 * xbytes <= 110 * 1024 ? 2 : 1;
        if (force_nbuf == 1) g.nbuf = 1;
        if (force_nbuf == 2 && bank_bytes + 2 * xbytes <= 227 * 1024) g.nbuf = 2;
        const size_t smem = bank_bytes + g.nbuf * xbytes;
        if (smem > 227 * 1024 || (max_smem && smem > max_smem)) return false;
        g.ps = g.nseg == 1 ? c.n_pos : (((c.n_pos + g.nseg - 1) / g.nseg + TJ - 1) / TJ) * TJ;
        if (g.nseg > 1 && g.ps < 2 * TJ) return false;  // segments shorter than two tiles: not worth it
        g.n_tiles = (g.ps + TJ - 1) / TJ;
        // persistent over a few tiles (bank staged once, prefetch) while the grid still fills the GPU
        const int64_t blocks_per_sm = std::max<int64_t>(1, (int64_t)(227 * 1024) / (int64_t)(smem + 1024));
        const int64_t slots = 148 * std::min<int64_t>(blocks_per_sm, 2048 / (NW * 32));
        int64_t tpb = (int64_t)g.n_tiles * g.n_sg / (slots * 4);
        tpb = std::max<int64_t>(1, std::min<int64_t>(tpb, 8));
        g.tiles_per_block = (int32_t)std::min<int64_t>(tpb, g.n_tiles);
        g.n_groups = (g.n_tiles + g.tiles_per_block - 1) / g.tiles_per_block;
        if (dry) return true;
        static size_t configured[64][8] = {{0}};
        size_t& conf = configured[dev & 63][slot + (c.in_f32 ? 4 : 0)];
        if (smem > conf) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            conf = smem;
        }
        const int64_t blocks = (int64_t)g.n_sg * g.n_groups + c.n_streams;
        kernel<<<(unsigned)blocks, NW * 32, smem, s>>>(c, g);
        count_launch();
        return true;
    };
    // Measured on B200. Short filters (x2 stage of the batched 44.1k->48k chain): 8 warps x 4 tiles, several blocks per SM;
    // 6 or 8 tiles per warp were slower. Long filters (C3: 8 ch x 1223 taps /2): the window of a tile is mostly taps, so the
    // 8-warp block is staged in 2-4 chunks along the taps until TWO blocks fit an SM (C3: 0.398 -> 0.374 ms against one
    // 16-warp block per SM with the whole window; chunking that 16-warp block or double-buffering the chunks was slower).
    // GAR_MMA_CFG = 1 / 3 forces 16 / 8 warps with the whole window, GAR_MMA_NKC = n the chunk count.
    static const int forced = [] { const char* e = gar::tune_env("GAR_MMA_CFG"); return e ? std::atoi(e) : -1; }();
    static const int nkc_env = [] { const char* e = gar::tune_env("GAR_MMA_NKC"); return e ? std::atoi(e) : 0; }();
    // float32 input (FirCall::in_f32): the IN32 instantiation of the same variant
    auto k16 = c.in_f32 ? fir_mma_f64_kernel<M, NF, 16, 4, false, true> : fir_mma_f64_kernel<M, NF, 16, 4, false, false>;
    auto k8 = c.in_f32 ? fir_mma_f64_kernel<M, NF, 8, 4, false, true> : fir_mma_f64_kernel<M, NF, 8, 4, false, false>;
    auto k8c = c.in_f32 ? fir_mma_f64_kernel<M, NF, 8, 4, true, true> : fir_mma_f64_kernel<M, NF, 8, 4, true, false>;
    if (forced == 1) return run(k16, 16, 4, 1, 1, 0);
    if (forced == 3) return run(k8, 8, 4, 3, 1, 0);
    if (c.taps > 600) {
        constexpr size_t two_per_sm = 113 * 1024;
        if (nkc_env > 1) return run(k8c, 8, 4, 2, nkc_env, 0);
        if (run(k8, 8, 4, 3, 1, two_per_sm)) return true;
        for (int nkc = 2; nkc <= 4; ++nkc)
            if (run(k8c, 8, 4, 2, nkc, two_per_sm)) return true;
        return run(k16, 16, 4, 1, 1, 0) || run(k8, 8, 4, 3, 1, 0);
    }
    return run(k8, 8, 4, 3, 1, 0);
}

static bool g_fir_mma = [] {
    const char* e = gar::tune_env("GAR_NO_MMA");
    return !(e && e[0] && e[0] != '0');
}();
// float64 FIR on the FP64 tensor cores: x2 up-sampler and /2 /3 /4 decimators, any number of lock-step rows (fewer than 8:
// time segments of the rows fill the MMA columns), calls of at least 32768 positions
static const char* launch_fir_mma(const FirCall& c, cudaStream_t s, const bool dry = false) {
    if (!g_fir_mma || c.n_streams < 1 || (int64_t)c.n_pos * c.n_streams < 32768 || c.taps < 16) return nullptr;
    // fewer than 8 rows (time-segment columns): only long calls — measured +5 % on 60 s of stereo 96k->48k, but a loss on the
    // sub-millisecond stages of a single 10 s stream (C5a), which are launch-latency-bound
    static const int64_t seg_min = [] { const char* e = gar::tune_env("GAR_MMA_SEG_MIN"); return e ? std::atoll(e) : 2000000ll; }();
    if (c.n_streams < 8 && (int64_t)c.n_pos * c.n_streams < seg_min) return nullptr;
    if (c.stride == 1 && c.nf == 2) return launch_fir_mma_t<1, 2>(c, s, dry) ? "fir_f64_mma_up2" : nullptr;
    if (c.nf == 1 && c.stride == 2) return launch_fir_mma_t<2, 1>(c, s, dry) ? "fir_f64_mma_s2" : nullptr;
    if (c.nf == 1 && c.stride == 3) return launch_fir_mma_t<3, 1>(c, s, dry) ? "fir_f64_mma_s3" : nullptr;
    if (c.nf == 1 && c.stride == 4) return launch_fir_mma_t<4, 1>(c, s, dry) ? "fir_f64_mma_s4" : nullptr;
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
            auto fma2 = [&](const int q, const int r, const u64 cf) {
                const int sh = (M * r) & 1;
                const int eh = (M * r - sh) / 2;
                const u64 xv = xw[(u * 2 + q + eh) % (NCH * 2)];
#pragma unroll
                for (int p = 0; p < NF; ++p) {
                    u64 d;
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(xv), "l"(cf), "l"(acc[r][p]));
                    acc[r][p] = d;
                }
            };
            if constexpr (NS == 2 && R == 12 && NF == 1) {
                // Two register copies of the step's four coefficient pairs, the second loaded again through an opaque
                // ld.shared so that ptxas cannot merge them: positions 0..5 read one copy, 6..11 the other. With ONE copy
                // ptxas chains 5-6 FFMA2s on the same coefficient register (.reuse) and the loop runs at the rate
                // tools/probe_fma_patterns3.cu measures for that operand pattern (50-63 TFLOP/s depending on the order it
                // picks); with two copies the kernel gains 2.4 % (a third copy loses it again: register pressure).
                u64 cw[NS][2];
#pragma unroll
                for (int sh = 0; sh < NS; ++sh) {
                    ulonglong2 v;
                    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(smem_u32(cs + sh * cp + it * 4)));
                    cw[sh][0] = v.x;
                    cw[sh][1] = v.y;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) fma2(q, r, r < R / 2 ? cv[0][r & 1][q] : cw[r & 1][q]);
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) fma2(q, r, cv[0][NS == 2 ? ((M * r) & 1) : 0][q]);
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



// ---- the same kernel with the taps in the CONSTANT bank (round 2) -----------------------------------------------------
// An FFMA2 whose three 64-bit operands all come from the register file runs at ~2/3 rate (tools/probe_fma_patterns3.cu);
// ptxas hides part of that with .reuse chains on the coefficient operand, which is what capped the kernel above at 77 % of
// the FMA probe. Here the filter travels as a KERNEL PARAMETER (two copies, shifted by 0 and 1 tap, zero-padded): every lane
// of a warp uses the same tap pair, so ptxas loads it into a UNIFORM register (LDCU.64 c[0x0][UR+imm]) and issues
// FFMA2 R, R, UR, R — two register-file operands, no coefficient LDS, no coefficient registers
// (tools/probe_fma_const.cu: 68.7 against 56.2 TFLOP/s for this loop's operand pattern). CPF = floats per copy.
template <int CPF>
struct FirTapsParam {
    unsigned long long pair[2][CPF / 2];  // pair[s][i] = (tap[2i - 4 - s], tap[2i + 1 - 4 - s]), zero outside the filter
};

template <int M, int NF, int R, int NT, int CPF>
__global__ void __launch_bounds__(NT) fir_f32x2c_kernel(const FirCall c, const int n_tiles, const int tiles_per_block,
                                                        const int n_groups, const int xlen, const int a_param,
                                                        const __grid_constant__ FirTapsParam<CPF> P) {
    static_assert(NF == 1, "decimators only");
    typedef unsigned long long u64;
    constexpr int MAXE = M * (R - 1) - ((M * (R - 1)) & 1);
    constexpr int NCH = (MAXE + 4 + 3) / 4;
    constexpr int NS = ((M & 1) && R > 1) ? 2 : 1;  // filter copies: shift 0 (even offsets), shift 1 (odd offsets)
    constexpr int TJ = NT * R;
    static_assert((TJ * M) % 4 == 0, "tile stride must keep the 16-byte alignment of the window");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // two mbarriers, one per window buffer
    float* xs0 = reinterpret_cast<float*>(smem_raw + 16);   // [2][xlen] double-buffered sample window

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
    // the pad is the same for every row (row stride a multiple of 16 bytes: checked by the launcher) and comes in as a
    // kernel parameter, so everything that indexes the constant bank is provably warp-uniform for ptxas
    const int a = a_param;
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
    // tap pair (t, t+1) with t = 4*it + 2*q - a - sh lives in copy s = (a + sh) & 1 at pair index (t + 4 + s) / 2
    int csel[NS], cbase[NS];
#pragma unroll
    for (int sh = 0; sh < NS; ++sh) {
        csel[sh] = (a + sh) & 1;
        cbase[sh] = (4 + csel[sh] - a - sh) >> 1;
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
        // tap-pair cursors into the parameter bank: advanced by the steps themselves (which run in order it = 0, 1, 2, ...), so
        // their address arithmetic never mixes with the per-thread shared-memory addressing and stays in uniform registers
        const u64* cq[NS];
#pragma unroll
        for (int sh = 0; sh < NS; ++sh) cq[sh] = &P.pair[csel[sh]][cbase[sh]];
        auto step = [&](const int u, const int it) {
            // (tools/probe_fma_patterns2.cu: FFMA2 loses ~15 % when one coefficient pair feeds 6 FMAs in a row;
            //  duplicating the pair with a second, unmergeable ld.shared cost more than it gained — measured)
            u64 cv[NF][NS][2];
#pragma unroll
            for (int sh = 0; sh < NS; ++sh) {
#pragma unroll
                for (int q = 0; q < 2; ++q) cv[0][sh][q] = cq[sh][q];  // warp-uniform address chain of its own: LDCU.64
                cq[sh] += 2;
            }
            auto fma2 = [&](const int q, const int r, const u64 cf) {
                const int sh = (M * r) & 1;
                const int eh = (M * r - sh) / 2;
                const u64 xv = xw[(u * 2 + q + eh) % (NCH * 2)];
#pragma unroll
                for (int p = 0; p < NF; ++p) {
                    u64 d;
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(xv), "l"(cf), "l"(acc[r][p]));
                    acc[r][p] = d;
                }
            };
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int r = 0; r < R; ++r) fma2(q, r, cv[0][NS == 2 ? ((M * r) & 1) : 0][q]);
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

// Integer factors without a register-tiled variant (x5, x6, x8 ... x24 up-samplers, /5 ... decimators: path-B engines such as
// 8k -> 48k or 8k -> 192k, dft_stage.go:156-338 / :488-554 with any factor): filter bank and sample window in SHARED memory,
// one thread per output, strictly sequential taps (the arithmetic of fir_generic_kernel, bit-identical), blocks persistent over
// `tiles_per_block` tiles so that the bank is staged once. Not FMA-bound (two LDS per FMA), but 5-10x the generic kernel, which
// reads every operand through guarded global loads.
template <typename T>
__global__ void __launch_bounds__(256) fir_smem_kernel(const FirCall c, const int tj /*positions per tile*/, const int n_tiles,
                                                       const int tiles_per_block, const int n_groups, const int bpitch,
                                                       const int xcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* bs = reinterpret_cast<T*>(smem_raw);  // [nf][bpitch] filter bank (odd pitch: the lanes' phases sit in different banks)
    T* xs = bs + (size_t)c.nf * bpitch;      // [xcap] sample window of the tile
    const int grp = blockIdx.x % (n_groups + 1);
    const int64_t row = blockIdx.x / (n_groups + 1);
    const T* __restrict__ hist = static_cast<const T*>(c.hist) + row * c.hist_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    if (grp == n_groups) {
        carry_row(hist, c.hist_len, in, c.n_in, static_cast<T*>(c.hist_out) + row * c.hist_out_stride, c.drop, c.new_hist_len);
        return;
    }
    const T* __restrict__ bank = static_cast<const T*>(c.bank);
    for (int i = threadIdx.x; i < c.nf * c.taps; i += 256) bs[(i / c.taps) * bpitch + (i % c.taps)] = bank[i];
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
    const int kc = (c.taps - 1) / 2;
    for (int t = grp * tiles_per_block; t < min(n_tiles, (grp + 1) * tiles_per_block); ++t) {
        const int j0 = t * tj;
        const int npos = min(tj, c.n_pos - j0);
        const int span = (npos - 1) * c.stride + c.taps;
        __syncthreads();  // the previous tile's window is no longer read (first tile: nothing pending)
        block_copy4(span, [&](int i) { return vload(hist, c.hist_len, in, c.n_in, c.first + j0 * c.stride + i); },
                    [&](int i, T v) { xs[i] = v; });
        __syncthreads();
        for (int o = threadIdx.x; o < npos * c.nf; o += 256) {
            const int j = o / c.nf, p = o - j * c.nf;
            const T* __restrict__ w = xs + j * c.stride;
            const T* __restrict__ cf = bs + p * bpitch;
            double tot = 0;
            T acc = 0;
            for (int k = 0; k < c.taps; ++k) {
                acc = fma(w[k], cf[k], acc);
                if (sizeof(T) == 4 && ((k & 255) == 255 || (k >= kc && k < kc + 12 && ((k - kc) & 3) == 3))) {
                    tot += (double)acc;
                    acc = 0;
                }
            }
            out[(int64_t)(j0 + j) * c.nf + p] = (T)(tot + (double)acc);
        }
    }
}

template <typename T>
static bool launch_fir_smem(const FirCall& c, cudaStream_t s) {
    if ((int64_t)c.n_pos * c.nf < 4096 || c.nf > 64 || c.stride > 64) return false;
    const int bpitch = c.taps | 1;
    // outputs per tile: up to 2048 (eight per thread) for large calls, down to 256 (one per thread) so that a single row of a
    // long-filter decimator (48k -> 8k: 80 000 outputs of ~3000 taps each) still spreads over four blocks per SM
    const int64_t n_el_total = (int64_t)c.n_pos * c.nf * c.n_streams;
    // (up-samplers stage nf filters per block: at least 1024 outputs per tile there, or a streaming-size x24 call spends its
    //  time copying the bank — 173 us per 4096-frame chunk with 256-output tiles)
    const int per_tile = (int)std::max<int64_t>(c.nf > 1 ? 1024 : 256, std::min<int64_t>(2048, n_el_total / (4 * 148)));
    const int tj = std::max(8, per_tile / c.nf);
    const int xcap = (tj - 1) * c.stride + c.taps;
    const size_t smem = ((size_t)c.nf * bpitch + xcap + 4) * sizeof(T);
    if (smem > 160 * 1024) return false;
    const int n_tiles = (c.n_pos + tj - 1) / tj;
    // persistent over a few tiles (the bank is staged once per block) while the grid still fills the GPU
    int tpb = (int)std::max<int64_t>(1, std::min<int64_t>(16, (int64_t)n_tiles * c.n_streams / (4 * 148)));
    const int n_groups = (n_tiles + tpb - 1) / tpb;
    auto k = fir_smem_kernel<T>;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(n_groups + 1) * c.n_streams;
    k<<<(unsigned)blocks, 256, smem, s>>>(c, tj, n_tiles, tpb, n_groups, bpitch, xcap);
    count_launch();
    return true;
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
    // float64: taps as kernel parameters (uniform-register DFMA operands) when they fit and every row has the same pad
    FirTapsD<(sizeof(T) == 8 ? NF : 1), (sizeof(T) == 8 ? FIR_TAPS_CPD : 2)> PU;
    int uc_a = -1;
    if constexpr (sizeof(T) == 8) {
        if ((c.in_stride & 1) == 0 && fill_fir_taps(PU, c.bank_host_f64, NF, c.taps))
            uc_a = (int)(((reinterpret_cast<uintptr_t>(c.in) / sizeof(T)) + (uintptr_t)(int64_t)(c.first - c.hist_len)) & 1u);
    } else {
        PU.c[0][0] = PU.c[0][1] = 0.0;
    }
    launch_pdl(k, (unsigned)blocks, (unsigned)NT, smem, s, c, n_tiles, tpb, n_groups, cp, xlen, uc_a, PU);
    count_launch();
}

template <int M, int NF, int R, int CPF>
bool launch_fir_f32x2c(const FirCall& c, cudaStream_t s) {
    constexpr int NT = 128;
    constexpr int MAXE = M * (R - 1) - ((M * (R - 1)) & 1);
    constexpr int NCH = (MAXE + 4 + 3) / 4;
    constexpr int NS = ((M & 1) && R > 1) ? 2 : 1;
    constexpr int TJ = NT * R;
    const int cp = ((c.taps + 3 + (NS - 1) + 3) / 4) * 4;
    // pair indices reach (4*n_iter + 4) / 2 with n_iter = (taps + 3 + (NS-1) + 3) / 4: keep a safety margin of 8 floats
    if (!c.bank_host_f32 || cp + 16 > CPF) return false;
    if (c.in_stride != 0 && (c.in_stride & 3) != 0) return false;  // rows must share the 16-byte alignment pad
    const int a_param = (int)(((reinterpret_cast<uintptr_t>(c.in) >> 2) + (uintptr_t)(int64_t)(c.first - c.hist_len)) & 3u);
    static_assert(sizeof(FirTapsParam<CPF>) + 256 <= 32764, "kernel parameters are limited to 32 KB");
    FirTapsParam<CPF> P;
    float* flat = reinterpret_cast<float*>(&P);
    for (int sc = 0; sc < 2; ++sc)
        for (int i = 0; i < CPF; ++i) {
            const int k = i - 4 - sc;
            flat[sc * CPF + i] = (k >= 0 && k < c.taps) ? c.bank_host_f32[k] : 0.f;
        }
    const int xlen = M * R * (NT - 1) + (cp / 4 + NCH + 1) * 4;
    const size_t smem = 16 + (size_t)(2 * xlen) * sizeof(float);
    const int n_tiles = (c.n_pos + TJ - 1) / TJ;
    int n_groups = 1;
    const int tpb = pick_tiles_per_block(n_tiles, c.n_streams, &n_groups);
    int dev = 0;
    cudaGetDevice(&dev);
    auto k = fir_f32x2c_kernel<M, NF, R, NT, CPF>;
    static size_t configured[64] = {0};
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(n_groups + 1) * c.n_streams;
    k<<<(unsigned)blocks, NT, smem, s>>>(c, n_tiles, tpb, n_groups, xlen, a_param, P);
    count_launch();
    return true;
}

template <int M, int NF, int R>
void launch_fir_f32x2(const FirCall& c, cudaStream_t s) {
    constexpr int NT = 128;
    if constexpr (NF == 1) {  // taps as kernel parameters (uniform-register FFMA2 operands) when they fit 8 / 24 KB
        static const bool no_const = [] { const char* e = gar::tune_env("GAR_NO_CONST_TAPS"); return e && e[0] && e[0] != '0'; }();
        if (!no_const && c.n_pos * (int64_t)c.n_streams >= 4096 &&
            (launch_fir_f32x2c<M, NF, R, 1024>(c, s) || launch_fir_f32x2c<M, NF, R, 3072>(c, s)))
            return;
    }
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

// tiles per block: long runs amortise the per-block filter load and hide the TMA prefetch under the FMAs,
// but keep >= ~3 blocks per resident slot in flight for balance
int pick_tiles_per_block(int n_tiles, int n_streams, int* n_groups) {
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


bool tensor_fir_enabled() { return g_fir_mma; }

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
    X(double, DT_F64, 1, 4, 2, "fir_f64_up4_r2")  \
    /* path-B engines at /5 /6 /8 and x5 x6 x8 (8k <-> 48k, 44.1k -> 8.82k ...). A thread's window starts M*R samples after its \
       neighbour's and is read with 16-byte loads: M*R must be even (float64) / a multiple of 4 (float32) */ \
    X(double, DT_F64, 5, 1, 2, "fir_f64_s5_r2")   \
    X(double, DT_F64, 6, 1, 3, "fir_f64_s6_r3")   \
    X(double, DT_F64, 8, 1, 2, "fir_f64_s8_r2")   \
    X(double, DT_F64, 1, 5, 2, "fir_f64_up5_r2")  \
    X(double, DT_F64, 1, 6, 2, "fir_f64_up6_r2")  \
    X(double, DT_F64, 1, 8, 2, "fir_f64_up8_r2")  \
    X(float, DT_F32, 5, 1, 4, "fir_f32_s5_r4")    \
    X(float, DT_F32, 6, 1, 6, "fir_f32_s6_r6")    \
    X(float, DT_F32, 8, 1, 3, "fir_f32_s8_r3")    \
    X(float, DT_F32, 1, 5, 4, "fir_f32_up5_r4")   \
    X(float, DT_F32, 1, 6, 4, "fir_f32_up6_r4")

// float64 streaming-size / flush calls (a handful of tiles at most): 2-3 positions per thread instead of 6-7, so the one
// tile that is the whole critical path is three times shorter and the positions spread over more threads; same tap order
// per output, bit-identical to the large-tile variants
#define GAR_FIR_SMALL_VARIANTS(X)                 \
    X(double, DT_F64, 1, 2, 2, "fir_f64_up2_r2")  \
    X(double, DT_F64, 2, 1, 3, "fir_f64_s2_r3")   \
    X(double, DT_F64, 3, 1, 2, "fir_f64_s3_r2")

void set_tensor_fir(bool on) { g_fir_mma = on; }

// shared memory a register-tiled variant needs for this call (filter + two window buffers): long filters of large factors
// (/8 VeryHigh: ~10 000 taps) do not fit and take the shared-memory / generic kernels
template <typename T, int M, int NF, int R>
static bool fir_tiled_fits(const FirCall& c) {
    constexpr int NT = 128, VEC = VecOf<T>::N, NCH = (M * (R - 1) + VEC - 1) / VEC + 1;
    const int cp = ((c.taps + VEC - 1 + VEC - 1) / VEC) * VEC;
    const int xlen = M * R * (NT - 1) + (cp / VEC + NCH + 1) * VEC;
    return 16 + (size_t)(NF * cp + 2 * xlen) * sizeof(T) <= 227 * 1024;
}

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

// float32 input / output folded into the float64 tensor-core FIR kernels (K1m / K2m): will launch_fir take the call with
// in_f32 / out_f32 set? (a dry run of the same dispatch)
bool fir_mma_io32_takes(const FirCall& c) {
    // Long filters stage every input sample 4-5 times (the window of a tile is mostly taps, and it is staged per tap chunk), so
    // widening in shared memory costs more than the one-pass cast launch once the call is large: measured on the 913-tap /2
    // stage of 48k -> 16k x 256 rows x 10 s, 3.5 against 2.8 + 0.25 ms. Small calls (launch-latency-bound) still gain:
    // 8 channels x 10 s of 96k -> 48k VeryHigh, 0.316 against 0.344 ms.
    if (c.in_f32 && c.taps > 600 && (int64_t)c.n_pos * c.n_streams > (4 << 20)) return false;
    return c.n_pos > 0 && (!c.in_f32 || (c.in_stride & 3) == 0) && launch_fir_mma(c, nullptr, true) != nullptr;
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
    if (c.in_f32 || c.out_f32) return nullptr;  // only the tensor-core kernels take float32 I/O (the engine asked fir_mma_io32_takes first)
#define X(M, NF, R, NAME)                                       \
    if (dtype == DT_F32 && c.stride == M && c.nf == NF) {       \
        launch_fir_f32x2<M, NF, R>(c, s);                       \
        return NAME;                                            \
    }
    GAR_FIR_X2_VARIANTS(X)
#undef X
    static const int64_t small_max = [] { const char* e = gar::tune_env("GAR_FIR_SMALL_MAX"); return e ? std::atoll(e) : 4096ll; }();
    if ((int64_t)c.n_pos * c.n_streams <= small_max) {
#define X(T, DT, M, NF, R, NAME)                           \
    if (dtype == DT && c.stride == M && c.nf == NF) {      \
        launch_fir_tiled<T, M, NF, R>(c, s);               \
        return NAME;                                       \
    }
        GAR_FIR_SMALL_VARIANTS(X)
#undef X
    }
#define X(T, DT, M, NF, R, NAME)                                                          \
    if (dtype == DT && c.stride == M && c.nf == NF && fir_tiled_fits<T, M, NF, R>(c)) {   \
        launch_fir_tiled<T, M, NF, R>(c, s);                                              \
        return NAME;                                                                      \
    }
    GAR_FIR_VARIANTS(X)
#undef X
    static const bool smem_on = [] { const char* e = gar::tune_env("GAR_NO_FIR_SMEM"); return !(e && e[0] && e[0] != '0'); }();
    if (smem_on && g_fir_mma) {  // (gar_set_tensor_fir(0) = the "simple kernels" A/B switch also restores the generic kernel here)
        if (dtype == DT_F32 ? launch_fir_smem<float>(c, s) : launch_fir_smem<double>(c, s))
            return dtype == DT_F32 ? "fir_f32_smem" : "fir_f64_smem";
    }
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


}  // namespace gar
