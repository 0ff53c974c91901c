// kernels_chain.cu — K5: the batched x2 -> polyphase chain of one engine.Resampler (resampler.go:97-121, :182-227; the chaining of
// constant.go:299-345) as ONE persistent launch per Process call. The intermediate-rate samples never round-trip to HBM and
// no full-size intermediate buffer exists: they live in a ring of R chunks per row (a few tens of MB for the whole batch)
// that stays resident in the 126 MB L2.
//
//   * Work items of the two stages come from one ordered queue (an atomic counter):  slot s = the x2 items of chunk s
//     (8 rows x a few 128-position MMA tiles each, the arithmetic of K1m), then the polyphase items of chunk s - D
//     (64 outputs x up to 256 rows each, the arithmetic of K3p: register-resident coefficient matrices, TMA producer warp).
//     Chunk c of the x2 stage = intermediate samples [c*S, (c+1)*S); chunk c of the polyphase stage = the 64-output tiles
//     whose sample windows END inside chunk c.
//   * A polyphase item waits until the x2 items of chunks c-2 .. c have all finished (per-chunk completion counters,
//     red.release / ld.acquire at gpu scope); an x2 item of chunk c waits until the polyphase chunks that read the ring slot
//     it overwrites (<= c-R+2) have finished. An item only ever waits for items EARLIER in the queue, and items are claimed in
//     order by running blocks, so the schedule cannot deadlock whatever the number of co-resident blocks. The lag D and the
//     ring depth R = 2D + 2 are sized so that a dependency is two waves of items old when it is checked: the waits are
//     almost never taken.
//   * Both kinds of blocks share an SM (two 288-thread blocks per SM): the polyphase items' coefficient gathers hide under
//     the x2 items' MMAs.
//   * The first H samples of the ring are mirrored behind its end, so that a window that starts near the end of the ring is
//     contiguous for the TMA bulk copy.
// Results are bit-identical to the two stand-alone launches (same cores, mma_cores.cuh).
#include "mma_cores.cuh"

namespace gar {
namespace {

struct ChainGeom {
    // x2 stage: 8 warps x 4 MMA tiles x 4 positions = 128 positions = 256 intermediate samples per tile
    int32_t nk, blen, xlen_u, pitch_u;
    int32_t n_tiles_u, n_rg_u;            // tiles per row, 8-row groups
    int32_t tiles_per_chunk, tiles_per_item, items_u_chunk;
    // ring
    int32_t S, R, RS, H;                  // chunk length, chunks in the ring, ring length R*S, mirrored head (samples)
    // polyphase stage: 8 warp tasks x 8 outputs = 64 outputs per tile, stages of RB rows
    int32_t span, pitch_p, kp, n_tiles_p, nrb, n_rg_p;
    int32_t items_p_chunk;                // polyphase items reserved per slot (max over the chunks)
    int32_t NC, D, slot_items;            // chunks, queue lag of the polyphase items, items per slot
    int32_t carry_rows;                   // rows per carry item
    int32_t n_carry_items;                // per stage
    int32_t total_items;
};

constexpr int WS_QUEUE = 0, WS_HDR = 4;  // int32 words: queue head, 3 spare; then udone[NC], pdone[NC]
constexpr int CH_TJ = 128, CH_TO = 64;

// number of 64-output tiles whose windows (with the staging over-read) end before intermediate sample (c+1)*S
__host__ __device__ inline int chain_p_tile_hi(const ChainGeom& g, const int c, const int hp, const int64_t L, const int64_t at0,
                                                const int64_t step, const int n_out) {
    if (c < 0) return 0;
    if (c >= g.NC - 1) return g.n_tiles_p;
    const int64_t X = (int64_t)(c + 1) * g.S + hp - g.kp - 8;  // largest admissible window start d_n (virtual-input index)
    if (X < 0) return 0;
    const int64_t num = (((X + 1) * L) << 16) - at0;
    int64_t n_hi = num <= 0 ? 0 : (num + step - 1) / step;
    if (n_hi > n_out) n_hi = n_out;
    return (int)(n_hi / CH_TO);
}

__host__ __device__ inline int chain_u_items(const ChainGeom& g, const int c) {
    const int tiles_c = min(g.tiles_per_chunk, g.n_tiles_u - c * g.tiles_per_chunk);
    return tiles_c <= 0 ? 0 : g.n_rg_u * ((tiles_c + g.tiles_per_item - 1) / g.tiles_per_item);
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, const int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread: wait until *p >= target. A dependency is an item claimed earlier by a running block, so the wait is finite; the
// bound (~2 s) turns a scheduling bug into a loud launch failure instead of a hung device.
__device__ __forceinline__ void wait_count(const int* p, const int target) {
    if (target <= 0 || ld_acquire_gpu(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < target) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void bar_sync_mma_warps() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int NK, int RB, int NST>
__global__ void __launch_bounds__(288, 2) chain_up2_poly_kernel(const FirCall cu, const PolyCall cp, const ChainGeom g,
                                                                int* __restrict__ ws) {
    constexpr int MT = 4, SH = 1, WA = 4, JT = 4, BOFF = 3;  // x2 stage: M = 1, NF = 2
    constexpr int NTASK = 8, RN = 8, NT8 = RB / 8;
    static_assert(RB == 32 || RB == 16, "a stage is 32 or 16 rows");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* pfull = reinterpret_cast<uint64_t*>(smem_raw);        // [NST] polyphase stage filled
    uint64_t* pempty = pfull + NST;                                 // [NST] polyphase stage released (8 MMA warps)
    uint64_t* ubar = reinterpret_cast<uint64_t*>(smem_raw + 64);    // [2] x2 window buffers
    volatile int* s_item = reinterpret_cast<volatile int*>(smem_raw + 96);  // [2] claimed queue index (double-buffered)
    double* Bs = reinterpret_cast<double*>(smem_raw + 128);         // [2][blen] zero-padded x2 bank, staged once per block
    double* un = Bs + 2 * g.blen;                                   // x2 windows [2][8][pitch_u]  |  polyphase stages [NST][RB][pitch_p]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* udone = ws + WS_HDR;
    int* pdone = udone + g.NC;
    double* const ring = static_cast<double*>(cu.out);
    const int64_t ring_stride = cu.out_stride;

    {
        const double* __restrict__ bank = static_cast<const double*>(cu.bank);
        for (int idx = tid; idx < 2 * g.blen; idx += 288) {
            const int p = idx / g.blen, k = idx - p * g.blen - BOFF;
            Bs[idx] = (k >= 0 && k < cu.taps) ? bank[p * cu.taps + k] : 0.0;
        }
    }
    if (tid == 0) {
        for (int b = 0; b < NST; ++b) {
            mbar_init(pfull + b, 1);
            mbar_init(pempty + b, NTASK);
        }
        mbar_init(ubar, 1);
        mbar_init(ubar + 1, 1);
    }
    uint32_t uph0 = 0u, uph1 = 0u;  // parities of the x2 window barriers
    uint32_t stage_ctr = 0u;        // polyphase stages used so far by this block (all warps count alike)
    const int first_slot_item = g.n_carry_items;
    const int first_pcarry_item = first_slot_item + (g.NC + g.D) * g.slot_items;

    for (int round = 0;; ++round) {
        if (tid == 0) s_item[round & 1] = atomicAdd(ws + WS_QUEUE, 1);
        __syncthreads();  // also: the previous item is finished by every warp (shared memory reusable), first round: set-up done
        const int item = s_item[round & 1];
        if (item >= g.total_items) break;

        if (item < first_slot_item) {
            // ---------------- carried tail of the x2 stage (dft_stage.go:199-203): rows of one group ----------------
            const int r0 = item * g.carry_rows;
            for (int r = r0; r < min(cu.n_streams, r0 + g.carry_rows); ++r)
                carry_row(static_cast<const double*>(cu.hist) + (int64_t)r * cu.hist_stride, cu.hist_len,
                          static_cast<const double*>(cu.in) + (int64_t)r * cu.in_stride, cu.n_in,
                          static_cast<double*>(cu.hist_out) + (int64_t)r * cu.hist_out_stride, cu.drop, cu.new_hist_len);
            continue;
        }
        if (item >= first_pcarry_item) {
            // ---------------- carried tail of the polyphase stage (polyphase_stage.go:296-304), read from the ring ----------------
            if (tid == 0)
                for (int c = max(0, g.NC - 3); c < g.NC; ++c) wait_count(udone + c, chain_u_items(g, c));
            __syncthreads();
            const int r0 = (item - first_pcarry_item) * g.carry_rows;
            for (int r = r0; r < min(cp.n_streams, r0 + g.carry_rows); ++r) {
                const double* __restrict__ hist = static_cast<const double*>(cp.hist) + (int64_t)r * cp.hist_stride;
                const double* __restrict__ rrow = ring + (int64_t)r * ring_stride;
                double* __restrict__ ho = static_cast<double*>(cp.hist_out) + (int64_t)r * cp.hist_out_stride;
                for (int i = tid; i < cp.new_hist_len; i += 288) {
                    const int64_t v = (int64_t)cp.drop + i;  // index into hist_p ++ mid
                    double x = 0.0;
                    if (v < cp.hist_len) x = hist[v];
                    else if (v - cp.hist_len < cp.n_in) x = __ldcg(rrow + (v - cp.hist_len) % g.RS);
                    ho[i] = x;
                }
            }
            continue;
        }
        const int si = item - first_slot_item;
        const int slot = si / g.slot_items, w = si - slot * g.slot_items;

        if (w < g.items_u_chunk) {
            // =============================== x2 item: chunk `slot`, 8 rows, nt tiles ===============================
            const int c = slot;
            if (c >= g.NC) continue;
            const int tiles_c = min(g.tiles_per_chunk, g.n_tiles_u - c * g.tiles_per_chunk);
            const int rg = w % g.n_rg_u, tg = w / g.n_rg_u;
            const int nt = min(g.tiles_per_item, tiles_c - tg * g.tiles_per_item);
            if (nt <= 0) continue;
            // the ring slot (and, for the first tiles of a ring cycle, the mirror) is free once its readers have finished
            if (tid == 0) {
                for (int j = max(0, c - g.R - 1); j <= c - g.R + 2; ++j)
                    wait_count(pdone + j, g.n_rg_p * (chain_p_tile_hi(g, j, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out) -
                                                      chain_p_tile_hi(g, j - 1, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out)));
            }
            __syncthreads();
            if (warp < 8) {
                const int sbase = rg * 8;
                const int tile0 = c * g.tiles_per_chunk + tg * g.tiles_per_item;
                const bool rows_bulk = (cu.in_stride & 1) == 0 && sbase + 8 <= cu.n_streams;
                const int64_t total = (int64_t)cu.hist_len + cu.n_in;
                const int nq = g.nk + (MT - 1) * SH;
                const int xbuf = 8 * g.pitch_u;
                const double* __restrict__ col0_in = static_cast<const double*>(cu.in) + (int64_t)sbase * cu.in_stride;
                auto tile_geom = [&](const int it, int& jb0, int& len, int& a, int& wlen) -> bool {
                    jb0 = (tile0 + it) * CH_TJ;
                    const int npos_t = min(CH_TJ, cu.n_pos - jb0);
                    const int len_full = ((npos_t + JT - 1) / JT * JT - 1) + 4 * g.nk + 4;
                    len = max(0, min(g.xlen_u, len_full));
                    a = 0;
                    wlen = 0;
                    if (!rows_bulk || len == 0) return false;
                    const int64_t g0 = (int64_t)jb0 - cu.hist_len;  // index into `in` of the window start
                    if (g0 < 0) return false;
                    a = (int)((reinterpret_cast<uintptr_t>(col0_in + g0) & 15u) >> 3);  // start `a` samples early: aligned sources
                    wlen = (len + a + 1) & ~1;
                    if (g0 - a >= 0 && g0 - a + wlen <= cu.n_in && wlen <= g.pitch_u) return true;
                    a = 0;
                    return false;
                };
                auto issue = [&](const int it, const int buf) {  // one thread
                    int jb0, len, a, wlen;
                    if (!tile_geom(it, jb0, len, a, wlen)) return;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(ubar + buf, (uint32_t)(8 * wlen * sizeof(double)));
                    for (int r = 0; r < 8; ++r)
                        bulk_g2s(un + buf * xbuf + r * g.pitch_u, col0_in + (int64_t)r * cu.in_stride + ((int64_t)jb0 - cu.hist_len - a),
                                 (uint32_t)(wlen * sizeof(double)), ubar + buf);
                };
                double acc[MT][2];
                double Areg[WA];
                for (int it = 0; it < nt; ++it) {
                    const int buf = it & 1;
                    double* __restrict__ Xs = un + buf * xbuf;
                    int jb0, len, a, wlen;
                    const bool bulk = tile_geom(it, jb0, len, a, wlen);
                    if (it > 0) bar_sync_mma_warps();  // everyone is done with the windows this iteration's prefetch overwrites
                    if (tid == 0) {
                        if (it == 0) issue(0, 0);
                        if (it + 1 < nt) issue(it + 1, buf ^ 1);
                    }
                    if (bulk) {
                        const uint32_t ph = buf ? uph1 : uph0;
                        while (!mbar_try_wait(ubar + buf, ph)) {
                        }
                        if (buf) uph1 ^= 1u;
                        else uph0 ^= 1u;
                    } else {  // edge tile (the carried tail, the end of the rows, a ragged last row group): element copies
                        for (int r = warp; r < 8; r += 8) {
                            double* __restrict__ dst = Xs + r * g.pitch_u;
                            if (sbase + r >= cu.n_streams) {
                                for (int i = lane; i < len; i += 32) dst[i] = 0.0;
                                continue;
                            }
                            const int64_t row = sbase + r;
                            const int64_t v0 = jb0;  // element i of the window is v[v0 + i]
                            const int i1 = (int)min((int64_t)len, max((int64_t)0, (int64_t)cu.hist_len - v0));
                            const int i2 = (int)min((int64_t)len, max((int64_t)i1, total - v0));
                            const double* __restrict__ hsrc = static_cast<const double*>(cu.hist) + row * cu.hist_stride + v0;
                            const double* __restrict__ isrc = static_cast<const double*>(cu.in) + row * cu.in_stride + (v0 - cu.hist_len);
                            for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
                            for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
                            for (int i = i2 + lane; i < len; i += 32) dst[i] = 0.0;
                        }
                        cp_async_wait_all();
                        bar_sync_mma_warps();
                    }
                    const int npos_t = min(CH_TJ, cu.n_pos - jb0);
                    if (warp * MT * JT < npos_t) {
#pragma unroll
                        for (int b = 0; b < MT; ++b) acc[b][0] = acc[b][1] = 0.0;
                        const double* __restrict__ xw = Xs + (lane >> 2) * g.pitch_u + (lane & 3) + a + 4 * (warp * MT * SH);
                        const double* __restrict__ aw = Bs + ((lane >> 2) % 2) * g.blen + BOFF + (lane & 3) - ((lane >> 2) / 2);
                        fir_mma_warp_tiles<MT, SH>(acc, Areg, xw, aw, g.nk, nq, 0, nq);
                        // D[row = lane/4][cols 2*(lane%4), +1]: intermediate sample (jb0 + ..)*2 + row of rows sbase + col, into the ring
                        const int r8 = lane >> 2;
                        const int ob = (int)(((int64_t)jb0 * 2) % g.RS);  // tiles are 256-sample aligned and RS is a multiple of 256
                        const bool mirror = ob < g.H;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 2 * (lane & 3) + e;
                            if (sbase + col >= cu.n_streams) continue;
                            double* __restrict__ orow = ring + (int64_t)(sbase + col) * ring_stride + ob;
#pragma unroll
                            for (int b = 0; b < MT; ++b) {
                                const int jl = (warp * MT + b) * JT;  // tile-local position of MMA tile b
                                if (jb0 + jl + r8 / 2 < cu.n_pos) {
                                    orow[jl * 2 + r8] = acc[b][e];
                                    if (mirror) orow[g.RS + jl * 2 + r8] = acc[b][e];
                                }
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");  // the ring is read by TMA (async proxy) in other blocks
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_release_gpu(udone + c, 1);
            }
            continue;
        }

        // =============================== polyphase item: chunk slot - D, one 64-output tile x nrb row stages ===============================
        const int c = slot - g.D;
        if (c < 0 || c >= g.NC) continue;
        const int tl = chain_p_tile_hi(g, c - 1, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out);
        const int ntile = chain_p_tile_hi(g, c, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out) - tl;
        const int wp = w - g.items_u_chunk;
        if (ntile <= 0 || wp >= ntile * g.n_rg_p) continue;
        const int tile = tl + wp % ntile, rgp = wp / ntile;
        if (tid == 0)
            for (int j = max(0, c - 2); j <= c; ++j) wait_count(udone + j, chain_u_items(g, j));
        __syncthreads();

        const int rows_base = rgp * RB * g.nrb;
        const int64_t L = cp.L;
        const int n0 = tile * CH_TO;
        const int n1 = min(cp.n_out, n0 + CH_TO);
        const int64_t d_base = ((cp.at0 + (int64_t)n0 * cp.step) >> 16) / L;  // first staged sample = window of output n0
        const int64_t d_last = ((cp.at0 + (int64_t)(n1 - 1) * cp.step) >> 16) / L;
        const int span_t = min((int)(d_last - d_base) + g.kp + 4, g.span);
        const int64_t total = (int64_t)cp.hist_len + cp.n_in;
        const int64_t gi = d_base - cp.hist_len;  // intermediate-sample index of the first staged sample
        const int nj = min(g.nrb, (cp.n_streams - rows_base + RB - 1) / RB);
        // stage kind, the same pure function on both sides of the pipeline: a full row block whose span lies inside the
        // produced intermediate samples is moved by TMA, started `a` samples early (ring rows are 16-byte aligned, RS is even)
        auto stage_is_bulk = [&](const int row0, int& a, int& wlen) -> bool {
            a = 0;
            wlen = 0;
            if (row0 + RB > cp.n_streams || gi < 0) return false;
            a = (int)(gi & 1);
            wlen = (span_t + a + 1) & ~1;
            if (gi - a >= 0 && gi - a + wlen <= cp.n_in && wlen <= g.pitch_p) return true;
            a = 0;
            return false;
        };

        if (warp == NTASK) {
            // ---------------- producer warp ----------------
            asm volatile("fence.proxy.async.global;" ::: "memory");
            const int m_bulk = gi >= 1 ? (int)((gi - (gi & 1)) % g.RS) : 0;
            for (int j = 0; j < nj; ++j) {
                const uint32_t u = stage_ctr + (uint32_t)j;
                const int buf = (int)(u % NST);
                const uint32_t k = u / NST;
                const int row0 = rows_base + j * RB;
                double* xs = un + buf * RB * g.pitch_p;
                if (k >= 1) {  // the MMA warps have released the stage's previous contents
                    while (!mbar_try_wait(pempty + buf, (k - 1) & 1u)) {
                    }
                    __syncwarp();
                }
                int a, wlen;
                if (stage_is_bulk(row0, a, wlen)) {
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(pfull + buf, (uint32_t)(RB * wlen * sizeof(double)));
                    }
                    __syncwarp();
                    if (lane < RB)
                        bulk_g2s(xs + lane * g.pitch_p, ring + (int64_t)(row0 + lane) * ring_stride + m_bulk,
                                 (uint32_t)(wlen * sizeof(double)), pfull + buf);
                } else {  // edge stage (carried tail, end of the rows, ragged last row block): element copies by this warp
                    const int i1 = (int)min((int64_t)span_t, max((int64_t)0, (int64_t)cp.hist_len - d_base));
                    const int i2 = (int)min((int64_t)span_t, max((int64_t)i1, total - d_base));
                    const int m0 = (int)((d_base + i1 - cp.hist_len) % g.RS);  // ring index of element i1 (>= 0 whenever i1 < i2)
                    for (int r = 0; r < RB; ++r) {
                        const int64_t row = row0 + r;
                        double* __restrict__ dst = xs + r * g.pitch_p;
                        if (row >= cp.n_streams) {
                            for (int i = lane; i < span_t; i += 32) dst[i] = 0.0;
                            continue;
                        }
                        const double* __restrict__ hsrc = static_cast<const double*>(cp.hist) + row * cp.hist_stride + d_base;
                        const double* __restrict__ rsrc = ring + row * ring_stride + m0 - i1;
                        for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
                        for (int i = i1 + lane; i < i2; i += 32) dst[i] = __ldcg(rsrc + i);
                        for (int i = i2 + lane; i < span_t; i += 32) dst[i] = 0.0;
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(pfull + buf)) : "memory");
                }
            }
        } else {
            // ---------------- MMA warps: task = 8 outputs nf .. nf+7 ----------------
            const int nf = n0 + warp * RN;
            const int i = lane >> 2;  // this lane's output row of the MMA tile (polyphase_stage.go:260-264)
            const int64_t at = cp.at0 + (int64_t)(nf + i) * cp.step;
            const int64_t fullp = at >> 16;
            const int64_t dv = fullp / L;
            const int ph = (int)(fullp - dv * L);
            const int base = __shfl_sync(0xffffffffu, (int)(dv - d_base), 0);  // window offset of the task's first output
            const int o_i = (int)(dv - d_base) - base;
            const bool live = nf + i < n1;
            const int nks = g.kp >> 2;
            double A[NK];
            poly_gather_coeffs<NK>(A, cp, ph, o_i, (double)(int)(at & 0xFFFF) * (1.0 / 65536.0), live, nks, lane);
            for (int j = 0; j < nj; ++j) {
                const uint32_t u = stage_ctr + (uint32_t)j;
                const int buf = (int)(u % NST);
                const uint32_t k = u / NST;
                const int row0 = rows_base + j * RB;
                const double* __restrict__ xs = un + buf * RB * g.pitch_p;
                int apad, wlen;
                stage_is_bulk(row0, apad, wlen);
                while (!mbar_try_wait(pfull + buf, k & 1u)) {
                }
                __syncwarp();
                if (nf < n1) {
                    double acc[NT8][2];
#pragma unroll
                    for (int t = 0; t < NT8; ++t) acc[t][0] = acc[t][1] = 0.0;
                    const double* __restrict__ bp = xs + (lane >> 2) * g.pitch_p + base + apad + (lane & 3);
                    poly_mma_stage<NK, NT8>(acc, A, bp, g.pitch_p, nks);
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(pempty + buf)) : "memory");
                    if (live) {
#pragma unroll
                        for (int t = 0; t < NT8; ++t) {
                            const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                            if (s0 < cp.n_streams) (static_cast<double*>(cp.out) + s0 * cp.out_stride)[nf + i] = acc[t][0];
                            if (s0 + 1 < cp.n_streams) (static_cast<double*>(cp.out) + (s0 + 1) * cp.out_stride)[nf + i] = acc[t][1];
                        }
                    }
                } else {
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(pempty + buf)) : "memory");
                }
            }
        }
        stage_ctr += (uint32_t)nj;
        __syncthreads();  // every warp has read its last stage (the ring reads of this item are complete)
        if (tid == 0) {
            __threadfence();
            red_release_gpu(pdone + c, 1);
        }
    }
}

template <int NK, int RB, int NST>
bool launch_chain_t(const FusedCall& c, cudaStream_t s, ChainWs* wsp, const int variant) {
    const double r = (double)c.step / ((double)c.L * 65536.0);
    ChainGeom g{};
    // ---- polyphase geometry: K3p's (launch_poly_rows_pipe_t) ----
    const int omax = (int)std::ceil(7 * r) + 1;
    g.kp = ((omax + c.t2 + 3) / 4) * 4;
    if (g.kp > 4 * NK || g.kp > 2 * c.t2 + 8) return false;
    g.span = (int)std::ceil((CH_TO - 1) * r) + 1 + g.kp + 8;
    g.pitch_p = ((g.span + 2 + 15) / 16) * 16 + 4;
    g.n_tiles_p = (c.n_out + CH_TO - 1) / CH_TO;
    const int n_rb = (c.n_streams + RB - 1) / RB;
    g.nrb = 1;
    while (g.nrb < 256 / RB && g.nrb * 2 <= n_rb) g.nrb *= 2;
    g.n_rg_p = (c.n_streams + RB * g.nrb - 1) / (RB * g.nrb);
    const size_t p_bytes = (size_t)NST * RB * g.pitch_p * sizeof(double);
    // ---- x2 geometry: K1m's (launch_fir_mma_t<1, 2>, 8 warps x 4 tiles, whole window, two buffers) ----
    const int kpu = c.t1 + 3;
    g.nk = (kpu + 3) / 4;
    g.blen = (4 * g.nk + 3 + 5) & ~1;
    g.blen = ((g.blen + 7) & ~15) + 8;  // the two phase filters 8 doubles apart modulo 16 banks
    g.xlen_u = (CH_TJ - 1) + 4 * g.nk + 4 * 3 + 10;
    g.pitch_u = ((g.xlen_u + 15) / 16) * 16 + 4;
    const size_t u_bytes = (size_t)2 * 8 * g.pitch_u * sizeof(double);
    const size_t smem = 128 + (size_t)2 * g.blen * sizeof(double) + std::max(p_bytes, u_bytes);
    if (smem > 113 * 1024) return false;  // two blocks per SM
    g.n_tiles_u = (c.np + CH_TJ - 1) / CH_TJ;
    g.n_rg_u = (c.n_streams + 7) / 8;

    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_cta = 2 * n_sm;
    // ---- chunking: the largest chunk whose ring (R = 2D + 2 chunks, D = queue lag worth two waves of items) fits the budget ----
    static const int64_t ring_budget = [] { const char* e = gar::tune_env("GAR_CHAIN_RING_MB"); return (int64_t)(e ? std::atoi(e) : 40) << 20; }();
    static const int tpi_env = [] { const char* e = gar::tune_env("GAR_CHAIN_TPI"); return e ? std::atoi(e) : 2; }();
    static const int lag_items_pct = [] { const char* e = gar::tune_env("GAR_CHAIN_LAG_PCT"); return e ? std::atoi(e) : 200; }();
    const int64_t n_mid = 2 * (int64_t)c.np;
    g.H = ((g.pitch_p + 255) / 256) * 256;
    bool found = false;
    for (int S = 8192; S >= 1024; S >>= 1) {
        g.S = S;
        g.tiles_per_chunk = S / 256;
        g.tiles_per_item = std::max(1, std::min(tpi_env, g.tiles_per_chunk));
        g.items_u_chunk = g.n_rg_u * ((g.tiles_per_chunk + g.tiles_per_item - 1) / g.tiles_per_item);
        g.NC = (int)((n_mid + S - 1) / S);
        const int p_tiles_max = (int)std::ceil((double)S / r / CH_TO) + 2;
        const int slot_est = g.items_u_chunk + p_tiles_max * g.n_rg_p;
        g.D = std::max(1, std::min(24, (int)(((int64_t)n_cta * lag_items_pct / 100 + slot_est - 1) / slot_est)));
        g.R = 2 * g.D + 2;
        g.RS = g.R * S;
        const int64_t ring_bytes = (int64_t)(g.RS + g.H) * c.n_streams * 8;
        if (ring_bytes <= ring_budget || S == 1024) {
            found = ring_bytes <= std::max<int64_t>(ring_budget, 96ll << 20);
            break;
        }
    }
    if (!found || g.NC < 2 * g.R || c.new_hp > g.S || c.hp > g.S || g.H > g.S) return false;
    // polyphase items per slot: the exact maximum over the chunks
    int pmax = 0;
    {
        int prev = 0;
        for (int cc = 0; cc < g.NC; ++cc) {
            const int hi = chain_p_tile_hi(g, cc, c.hp, c.L, c.at0, c.step, c.n_out);
            pmax = std::max(pmax, hi - prev);
            prev = hi;
        }
    }
    g.items_p_chunk = pmax * g.n_rg_p;
    g.slot_items = g.items_u_chunk + g.items_p_chunk;
    g.carry_rows = 8;
    g.n_carry_items = (c.n_streams + g.carry_rows - 1) / g.carry_rows;
    g.total_items = 2 * g.n_carry_items + (g.NC + g.D) * g.slot_items;

    // ---- workspace: counters + ring (rows 16-byte aligned) ----
    const size_t cnt_bytes = (((size_t)(WS_HDR + 2 * g.NC) * 4) + 255) & ~(size_t)255;
    const int64_t ring_stride = ((int64_t)g.RS + g.H + 1) & ~int64_t(1);
    const size_t need = cnt_bytes + (size_t)ring_stride * (size_t)c.n_streams * 8;
    if (!wsp->dev || wsp->bytes < need) {
        if (wsp->dev) {
            cudaDeviceSynchronize();
            cudaFree(wsp->dev);
            wsp->dev = nullptr;
            wsp->bytes = 0;
        }
        const size_t nb = need + need / 8;
        if (cudaMalloc(&wsp->dev, nb) != cudaSuccess) {
            cudaGetLastError();
            wsp->dev = nullptr;
            return false;
        }
        wsp->bytes = nb;
        // stale ring contents are only ever multiplied by zero coefficients (staging over-read): they must be finite
        cudaMemsetAsync(wsp->dev, 0, nb, s);
    }
    // the counters sit behind the ring so that the ring rows keep the allocation's alignment
    double* ring = static_cast<double*>(wsp->dev);
    int* ws = reinterpret_cast<int*>(static_cast<char*>(wsp->dev) + (size_t)ring_stride * (size_t)c.n_streams * 8);
    cudaMemsetAsync(ws, 0, cnt_bytes, s);

    FirCall cu{};
    cu.hist = c.hist_u; cu.hist_stride = c.hist_u_stride; cu.hist_len = c.hu;
    cu.in = c.in; cu.in_stride = c.in_stride; cu.n_in = c.n_in;
    cu.out = ring; cu.out_stride = ring_stride;
    cu.hist_out = c.hist_u_out; cu.hist_out_stride = c.hist_u_out_stride;
    cu.drop = c.drop_u; cu.new_hist_len = c.new_hu;
    cu.bank = c.bank_u; cu.taps = c.t1; cu.stride = 1; cu.nf = 2; cu.first = 0; cu.n_pos = c.np; cu.n_streams = c.n_streams;
    PolyCall cp{};
    cp.hist = c.hist_p; cp.hist_stride = c.hist_p_stride; cp.hist_len = c.hp;
    cp.in = ring; cp.in_stride = ring_stride; cp.n_in = 2 * c.np;
    cp.out = c.out; cp.out_stride = c.out_stride;
    cp.hist_out = c.hist_p_out; cp.hist_out_stride = c.hist_p_out_stride;
    cp.drop = c.drop_p; cp.new_hist_len = c.new_hp;
    cp.bank_a = c.bank_a; cp.bank_b = c.bank_b; cp.bank_c = c.bank_c; cp.bank_d = c.bank_d;
    cp.taps = c.t2; cp.L = c.L; cp.at0 = c.at0; cp.step = c.step; cp.n_out = c.n_out; cp.interp = c.interp; cp.n_streams = c.n_streams;

    static size_t configured[64][4] = {{0}};
    size_t& conf = configured[dev & 63][variant];
    if (smem > conf) {
        cudaFuncSetAttribute(chain_up2_poly_kernel<NK, RB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conf = smem;
    }
    const int blocks = std::min(n_cta, g.total_items);
    chain_up2_poly_kernel<NK, RB, NST><<<(unsigned)blocks, 288, smem, s>>>(cu, cp, g, ws);
    count_launch();
    return true;
}

}  // namespace

static bool g_chain = [] {
    const char* e = gar::tune_env("GAR_NO_CHAIN");
    return !(e && e[0] && e[0] != '0');
}();
void set_chain_kernel(bool on) { g_chain = on; }
bool chain_kernel_enabled() { return g_chain; }

// K5 dispatch: float64 batches of at least 32 lock-step rows whose x2 and polyphase stages both run on the tensor cores
// (the K1m / K3p domain), calls long enough for a few ring cycles
bool launch_chain_up2_poly(const FusedCall& c, cudaStream_t s, ChainWs* ws) {
    if (!g_chain || !ws || !tensor_fir_enabled() || !tiled_polyphase_enabled()) return false;
    if (c.in_f32 || c.out_f32 || c.n_streams < 32 || c.t1 < 16 || c.np <= 0 || c.n_out <= 0) return false;
    if ((int64_t)c.np * c.n_streams < (1 << 22) || c.L > 4096 || c.t2 > 1024) return false;
    if (2 * (int64_t)c.np > 0x7ffffff0LL) return false;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.0) || r > 8.0) return false;
    return launch_chain_t<20, 32, 2>(c, s, ws, 0) || launch_chain_t<20, 16, 3>(c, s, ws, 1) ||
           launch_chain_t<28, 32, 2>(c, s, ws, 2) || launch_chain_t<28, 16, 3>(c, s, ws, 3);
}

}  // namespace gar
