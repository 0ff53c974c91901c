// kernels_chain.cu — K5: the batched x2 -> polyphase chain of one engine.Resampler (resampler.go:97-121, :182-227; the chaining of
// constant.go:299-345) as ONE persistent launch per Process call. The intermediate-rate samples never round-trip to HBM and
// no full-size intermediate buffer exists: they live in a ring of R chunks per row (a few tens of MB for the whole batch)
// that stays resident in the 126 MB L2.
//
//   * Work items. x2 item = 8 rows x 1-2 MMA tiles of 128 positions (the arithmetic of K1m); polyphase item = one 64-output
//     tile x up to 128 rows (the arithmetic of K3p: register-resident coefficient matrices). Chunk c of the x2 stage =
//     intermediate samples [c*S, (c+1)*S); chunk c of the polyphase stage = the 64-output tiles whose sample windows END inside
//     chunk c.
//   * Dataflow schedule. A polyphase item is READY when the x2 items of every chunk <= c have finished; an x2 item may
//     overwrite its ring slot when the polyphase chunks <= c-R+2 (the readers of that slot and of the mirror) have finished.
//     Completion is counted per chunk in global memory: red.release.gpu by the last of the 8 MMA warps to finish an item,
//     relaxed polls + one acquire fence on the reader's side. x2 items are claimed in order by fetch-and-add and stay
//     PENDING in their block until their slot is free; polyphase items are owned statically (item i belongs to block
//     i mod grid) and a block serves its own ready polyphase items first — they are what frees the ring — so the schedule
//     cannot deadlock whatever the number of co-resident blocks (induction over the chunk index, R > 2). A bounded spin turns
//     a scheduling bug into a launch error instead of a hung device.
//   * A block is 8 MMA warps + 1 scheduler / producer warp that claims items, checks their dependencies and streams their
//     sample windows into shared-memory stages by TMA ahead of the MMA warps, across item boundaries; two blocks share an
//     SM. The scheduler claims LATE (at most one issued stage unread): whatever queues inside a block lengthens the life of
//     its samples in the ring, and the ring has to hold (bytes per microsecond) x (microseconds from the x2 store to the last
//     polyphase read) — Little's law is what bounds this kernel (profiles/r2_chain_k5.md).
//   * Input windows are loaded with an L2 evict-first policy and outputs stored with streaming stores: what stays in L2 is
//     the ring.
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
    int32_t items_p_chunk;                // polyphase queue entries per chunk (max over the chunks; the rest are no-ops)
    int32_t NC;                           // chunks
    int32_t carry_rows;                   // rows per carry item
    int32_t n_carry_items;                // per stage
    int32_t ahead;                        // stages the scheduler may be ahead of the MMA warps when it claims an item
    int32_t stage_elems;                  // doubles per shared-memory stage: max(8 * pitch_u, RB * pitch_p)
};

constexpr int WS_QU = 0, WS_HDR = 32;  // int32 words: head of the x2 queue (alone in its 128-byte line); then udone[NC], pdone[NC]
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

// polls are relaxed loads (an acquire load costs an L1 invalidation, CCTL.IVALL, every time); the acquire fence follows once,
// when the value polled for has been seen
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// TMA bulk copy with an L2 evict-first policy: the input rows stream through L2 once, the ring should stay
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void red_release_gpu(int* p, const int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Descriptor of a claimed work item, handed from the scheduler warp to the MMA warps through a small ring in shared memory
struct ChainDesc {
    int32_t kind, c, a0, a1, a2, pad[3];  // kind: 0 exit, 1 x2 item (c, row group, first tile, tiles), 2 polyphase item (c, tile, row group),
};                                        //       3 / 4 carried tails of the x2 / polyphase stage (first row)
constexpr int CH_ND = 4;

// Block = 8 MMA warps + 1 scheduler / producer warp, no block-wide barrier after the set-up:
//   scheduler warp: takes the next ready item (see the file header), publishes its descriptor and streams its
//     sample windows into a ring of NST shared-memory stages by TMA bulk copies (x2 item: one 8-row window per 128-position
//     tile; polyphase item: one RB-row window block per stage) — it runs ahead of the MMA warps across item boundaries, so
//     the queue atomics, the dependency checks and the copy latencies hide under the MMAs of the items before;
//   MMA warps: take the descriptors in order, wait for each stage (mbarrier full), run the K1m / K3p core on it, release
//     it (mbarrier empty), store, and count the item as finished (one release-increment per item, by the last warp).
template <int NK, int RB, int NST>
__global__ void __launch_bounds__(288, 2) chain_up2_poly_kernel(const FirCall cu, const PolyCall cp, const ChainGeom g,
                                                                int* __restrict__ ws) {
    constexpr int MT = 4, SH = 1, WA = 4, JT = 4, BOFF = 3;  // x2 stage: M = 1, NF = 2
    constexpr int NTASK = 8, RN = 8, NT8 = RB / 8;
    static_assert(RB == 32 || RB == 16, "a stage is 32 or 16 rows");
    static_assert(NST <= 4, "barrier area");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sfull = reinterpret_cast<uint64_t*>(smem_raw);        // [NST] stage filled (producer arrival + TMA bytes)
    uint64_t* sempty = sfull + 4;                                   // [NST] stage released (8 MMA warps)
    uint64_t* dfull = sempty + 4;                                   // [CH_ND] descriptor published
    uint64_t* dempty = dfull + CH_ND;                               // [CH_ND] descriptor read (8 MMA warps)
    ChainDesc* desc = reinterpret_cast<ChainDesc*>(smem_raw + 128); // [CH_ND]
    int* icnt = &desc[0].pad[0];                                    // desc[ds].pad[0]: MMA warps that have finished items of slot ds
    double* Bs = reinterpret_cast<double*>(smem_raw + 256);         // [2][blen] zero-padded x2 bank, staged once per block
    double* stg = Bs + 2 * g.blen;                                  // [NST][stage_elems]: x2 windows [8][pitch_u] | polyphase [RB][pitch_p]

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
            mbar_init(sfull + b, 1);
            mbar_init(sempty + b, NTASK);
        }
        for (int b = 0; b < CH_ND; ++b) {
            mbar_init(dfull + b, 1);
            mbar_init(dempty + b, NTASK);
            desc[b].pad[0] = 0;
        }
    }
    __syncthreads();

    // ---- geometry shared by both sides (pure functions of the descriptor) ----
    const int64_t u_total = (int64_t)cu.hist_len + cu.n_in;
    const int64_t p_total = (int64_t)cp.hist_len + cp.n_in;
    const int64_t L = cp.L;
    // x2 tile `tile` of row group starting at row sbase: window length, alignment pad, TMA or element copies
    auto u_tile_geom = [&](const int sbase, const int tile, int& jb0, int& len, int& a, int& wlen) -> bool {
        jb0 = tile * CH_TJ;
        const int npos_t = min(CH_TJ, cu.n_pos - jb0);
        const int len_full = ((npos_t + JT - 1) / JT * JT - 1) + 4 * g.nk + 4;
        len = max(0, min(g.xlen_u, len_full));
        a = 0;
        wlen = 0;
        if ((cu.in_stride & 1) != 0 || sbase + 8 > cu.n_streams || len == 0) return false;
        const int64_t g0 = (int64_t)jb0 - cu.hist_len;  // index into `in` of the window start
        if (g0 < 0) return false;
        a = (int)((reinterpret_cast<uintptr_t>(static_cast<const double*>(cu.in) + (int64_t)sbase * cu.in_stride + g0) & 15u) >> 3);
        wlen = (len + a + 1) & ~1;
        if (g0 - a >= 0 && g0 - a + wlen <= cu.n_in && wlen <= g.pitch_u) return true;
        a = 0;
        return false;
    };
    // polyphase tile: first staged virtual-input sample and staged span
    auto p_tile_geom = [&](const int tile, int& n0, int& n1, int64_t& d_base, int& span_t) {
        n0 = tile * CH_TO;
        n1 = min(cp.n_out, n0 + CH_TO);
        d_base = ((cp.at0 + (int64_t)n0 * cp.step) >> 16) / L;
        const int64_t d_last = ((cp.at0 + (int64_t)(n1 - 1) * cp.step) >> 16) / L;
        span_t = min((int)(d_last - d_base) + g.kp + 4, g.span);
    };
    // a full row block whose span lies inside the produced intermediate samples is moved by TMA, started `a` samples early
    // (ring rows are 16-byte aligned and RS is even, so the pad is the parity of the sample index)
    auto p_stage_is_bulk = [&](const int row0, const int64_t gi, const int span_t, int& a, int& wlen) -> bool {
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
        // =========================================== scheduler / producer warp ===========================================
        uint32_t st = 0u, dk = 0u;
        uint64_t pol_stream;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        int u_ok = -1, p_ok = -1;  // every x2 / polyphase chunk up to here is known to be finished
        const int n_u_items = g.n_carry_items + g.NC * g.items_u_chunk;  // x2 queue: carried tails, then the chunks in order
        const int n_p_items = g.NC * g.items_p_chunk + g.n_carry_items;  // polyphase queue: the chunks in order, then the tails
        auto publish = [&](const int kind, const int c, const int a0, const int a1, const int a2) {
            const int ds = (int)(dk % CH_ND);
            if (dk >= CH_ND) {
                while (!mbar_try_wait(dempty + ds, ((dk / CH_ND) - 1u) & 1u)) {
                }
            }
            __syncwarp();
            if (lane == 0) {
                desc[ds].kind = kind; desc[ds].c = c; desc[ds].a0 = a0; desc[ds].a1 = a1; desc[ds].a2 = a2;
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(dfull + ds)) : "memory");
            }
            ++dk;
        };
        // non-blocking: advance the watermarks over the chunks whose completion counters have reached their targets
        auto u_done_upto = [&](const int c) -> bool {  // x2 chunks <= c finished (their ring samples are visible)
            const int before = u_ok;
            while (u_ok < c && ld_relaxed_gpu(udone + u_ok + 1) >= NTASK * chain_u_items(g, u_ok + 1)) ++u_ok;
            if (u_ok != before) fence_acquire_gpu();
            return u_ok >= c;
        };
        auto p_done_upto = [&](const int c) -> bool {  // polyphase chunks <= c finished (their ring reads are complete)
            const int before = p_ok;
            while (p_ok < c &&
                   ld_relaxed_gpu(pdone + p_ok + 1) >=
                       NTASK * g.n_rg_p * (chain_p_tile_hi(g, p_ok + 1, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out) -
                                           chain_p_tile_hi(g, p_ok, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out)))
                ++p_ok;
            if (p_ok != before) fence_acquire_gpu();
            return p_ok >= c;
        };
        auto stage_acquire = [&](int& buf) {  // the MMA warps have released the stage's previous contents
            buf = (int)(st % NST);
            const uint32_t k = st / NST;
            if (k >= 1u) {
                while (!mbar_try_wait(sempty + buf, (k - 1u) & 1u)) {
                }
            }
            __syncwarp();
        };
        // Polyphase items are owned statically (item i belongs to block i mod gridDim.x: no atomics, no contention when a chunk
        // becomes ready, and equal shares of these equal-cost items); x2 items are claimed by fetch-and-add. A claimed x2 item
        // whose ring slot is still being read stays pending while the block keeps serving its own polyphase items — the items
        // that free the ring — so the scheme cannot deadlock (induction over the chunk index, R > 2).
        int my_p = (int)blockIdx.x;  // next polyphase-queue item of this block
        int pending_u = -1;          // claimed x2 item waiting for its ring slot
        bool u_exhausted = false;
        for (;;) {
            // claim late: a new item is claimed only when at most `ahead` issued stages are still unread. An item claimed
            // earlier would only queue behind them, and everything queued lengthens the life of its samples in the ring.
            if (st > (uint32_t)g.ahead) {
                const uint32_t u = st - 1u - (uint32_t)g.ahead;  // this stage use must have been released
                while (!mbar_try_wait(sempty + u % NST, (u / NST) & 1u)) {
                }
            }
            int kind = -1, item = 0;
            const long long t0 = clock64();
            for (;;) {
                if (my_p < n_p_items) {
                    const int c_need = my_p < g.NC * g.items_p_chunk ? my_p / g.items_p_chunk : g.NC - 1;
                    if (u_done_upto(c_need)) {
                        kind = 2;
                        item = my_p;
                        my_p += (int)gridDim.x;
                        break;
                    }
                }
                if (pending_u < 0 && !u_exhausted) {
                    int got = 0;
                    if (lane == 0) got = atomicAdd(ws + WS_QU, 1);
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (got < n_u_items) pending_u = got;
                    else u_exhausted = true;
                }
                if (pending_u >= 0) {
                    // the ring slot (and, for the first tiles of a ring cycle, the mirror) is free once its readers have finished
                    const int c_u = pending_u < g.n_carry_items ? -1 : (pending_u - g.n_carry_items) / g.items_u_chunk;
                    if (c_u - g.R + 2 < 0 || p_done_upto(c_u - g.R + 2)) {
                        kind = 1;
                        item = pending_u;
                        pending_u = -1;
                        break;
                    }
                }
                if (my_p >= n_p_items && pending_u < 0 && u_exhausted) {
                    kind = 0;
                    break;
                }
                // nothing is ready: other blocks are working on the items this one waits for
                __nanosleep(100);
                if (clock64() - t0 > 20000000000ll) __trap();  // ~10 s: a scheduling bug must not hang the device
            }
            if (kind == 0) {
                publish(0, 0, 0, 0, 0);
                break;
            }
            if (kind == 1) {
                if (item < g.n_carry_items) {  // carried tail of the x2 stage: reads the input only
                    publish(3, 0, item * g.carry_rows, 0, 0);
                    continue;
                }
                // ---------------- x2 item: chunk c, 8 rows, nt tiles ----------------
                const int c = (item - g.n_carry_items) / g.items_u_chunk, w = (item - g.n_carry_items) - c * g.items_u_chunk;
                const int tiles_c = min(g.tiles_per_chunk, g.n_tiles_u - c * g.tiles_per_chunk);
                const int rg = w % g.n_rg_u, tg = w / g.n_rg_u;
                const int nt = min(g.tiles_per_item, tiles_c - tg * g.tiles_per_item);
                if (nt <= 0) continue;
                const int sbase = rg * 8, tile0 = c * g.tiles_per_chunk + tg * g.tiles_per_item;
                publish(1, c, rg, tile0, nt);
                for (int it = 0; it < nt; ++it) {
                    int buf;
                    stage_acquire(buf);
                    double* xs = stg + (size_t)buf * g.stage_elems;
                    int jb0, len, a, wlen;
                    if (u_tile_geom(sbase, tile0 + it, jb0, len, a, wlen)) {
                        if (lane == 0) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            mbar_expect_tx(sfull + buf, (uint32_t)(8 * wlen * sizeof(double)));
                        }
                        __syncwarp();
                        if (lane < 8)
                            bulk_g2s_stream(xs + lane * g.pitch_u,
                                            static_cast<const double*>(cu.in) + (int64_t)(sbase + lane) * cu.in_stride + ((int64_t)jb0 - cu.hist_len - a),
                                            (uint32_t)(wlen * sizeof(double)), sfull + buf, pol_stream);
                    } else {  // edge tile (the carried tail, the end of the rows, a ragged last row group): element copies
                        for (int r = 0; r < 8; ++r) {
                            double* __restrict__ dst = xs + r * g.pitch_u;
                            if (sbase + r >= cu.n_streams) {
                                for (int i = lane; i < len; i += 32) dst[i] = 0.0;
                                continue;
                            }
                            const int64_t row = sbase + r;
                            const int64_t v0 = jb0;  // element i of the window is v[v0 + i]
                            const int i1 = (int)min((int64_t)len, max((int64_t)0, (int64_t)cu.hist_len - v0));
                            const int i2 = (int)min((int64_t)len, max((int64_t)i1, u_total - v0));
                            const double* __restrict__ hsrc = static_cast<const double*>(cu.hist) + row * cu.hist_stride + v0;
                            const double* __restrict__ isrc = static_cast<const double*>(cu.in) + row * cu.in_stride + (v0 - cu.hist_len);
                            for (int i = lane; i < i1; i += 32) dst[i] = hsrc[i];
#pragma unroll 4
                            for (int i = i1 + lane; i < i2; i += 32) cp_async_elem(dst + i, isrc + i);
                            for (int i = i2 + lane; i < len; i += 32) dst[i] = 0.0;
                        }
                        cp_async_wait_all();
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sfull + buf)) : "memory");
                    }
                    ++st;
                }
                continue;
            }
            // ---------------- polyphase queue ----------------
            if (item >= g.NC * g.items_p_chunk) {  // carried tail of the polyphase stage: the last samples of the ring
                publish(4, 0, (item - g.NC * g.items_p_chunk) * g.carry_rows, 0, 0);
                continue;
            }
            // polyphase item: chunk c, one 64-output tile x nrb row stages
            const int c = item / g.items_p_chunk, wp = item - c * g.items_p_chunk;
            const int tl = chain_p_tile_hi(g, c - 1, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out);
            const int ntile = chain_p_tile_hi(g, c, cp.hist_len, cp.L, cp.at0, cp.step, cp.n_out) - tl;
            if (ntile <= 0 || wp >= ntile * g.n_rg_p) continue;
            const int tile = tl + wp % ntile, rgp = wp / ntile;
            asm volatile("fence.proxy.async.global;" ::: "memory");  // the ring was written through the generic proxy
            publish(2, c, tile, rgp, 0);
            int n0, n1, span_t;
            int64_t d_base;
            p_tile_geom(tile, n0, n1, d_base, span_t);
            const int64_t gi = d_base - cp.hist_len;  // intermediate-sample index of the first staged sample
            const int rows_base = rgp * RB * g.nrb;
            const int nj = min(g.nrb, (cp.n_streams - rows_base + RB - 1) / RB);
            const int m_bulk = gi >= 1 ? (int)((gi - (gi & 1)) % g.RS) : 0;
            for (int j = 0; j < nj; ++j) {
                int buf;
                stage_acquire(buf);
                double* xs = stg + (size_t)buf * g.stage_elems;
                const int row0 = rows_base + j * RB;
                int a, wlen;
                if (p_stage_is_bulk(row0, gi, span_t, a, wlen)) {
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(sfull + buf, (uint32_t)(RB * wlen * sizeof(double)));
                    }
                    __syncwarp();
                    if (lane < RB)
                        bulk_g2s(xs + lane * g.pitch_p, ring + (int64_t)(row0 + lane) * ring_stride + m_bulk,
                                 (uint32_t)(wlen * sizeof(double)), sfull + buf);
                } else {  // edge stage (carried tail, end of the rows, ragged last row block): element copies by this warp
                    const int i1 = (int)min((int64_t)span_t, max((int64_t)0, (int64_t)cp.hist_len - d_base));
                    const int i2 = (int)min((int64_t)span_t, max((int64_t)i1, p_total - d_base));
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
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sfull + buf)) : "memory");
                }
                ++st;
            }
        }
        return;
    }

    // =================================================== MMA warps ===================================================
    uint32_t st = 0u;
    const int ctid = tid;  // 0 .. 255
    // The item is counted once, by the LAST of the 8 MMA warps to finish it: the others synchronise with it at block scope
    // (acq_rel shared-memory atomic), its red.release at gpu scope is cumulative over their stores. One MEMBAR.GPU per item
    // instead of eight (they had the MMA warps stalled ~5 % of the time). The slot counters are never reset: use m of a slot
    // (descriptor number dk = m * CH_ND + slot) ends when its counter reaches 8 * (m + 1).
    auto item_done = [&](int* counter, const int ds, const uint32_t dk) {
        __syncwarp();
        if (lane == 0) {
            int prev;
            asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(icnt + ds * (int)(sizeof(ChainDesc) / 4))) : "memory");
            if (prev == (int)(NTASK * (dk / CH_ND + 1u)) - 1 && counter) red_release_gpu(counter, NTASK);
        }
    };
    for (uint32_t dk = 0u;; ++dk) {
        const int ds = (int)(dk % CH_ND);
        while (!mbar_try_wait(dfull + ds, (dk / CH_ND) & 1u)) {
        }
        const int kind = desc[ds].kind, c = desc[ds].c, a0 = desc[ds].a0, a1 = desc[ds].a1, a2 = desc[ds].a2;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(dempty + ds)) : "memory");
        if (kind == 0) break;
        if (kind == 3) {
            // ---------------- carried tail of the x2 stage (dft_stage.go:199-203) ----------------
            for (int r = a0; r < min(cu.n_streams, a0 + g.carry_rows); ++r) {
                const double* __restrict__ hist = static_cast<const double*>(cu.hist) + (int64_t)r * cu.hist_stride;
                const double* __restrict__ in = static_cast<const double*>(cu.in) + (int64_t)r * cu.in_stride;
                double* __restrict__ ho = static_cast<double*>(cu.hist_out) + (int64_t)r * cu.hist_out_stride;
                for (int i = ctid; i < cu.new_hist_len; i += 256) ho[i] = vload(hist, cu.hist_len, in, cu.n_in, cu.drop + i);
            }
            item_done(nullptr, ds, dk);
            continue;
        }
        if (kind == 4) {
            // ---------------- carried tail of the polyphase stage (polyphase_stage.go:296-304), read from the ring ----------------
            for (int r = a0; r < min(cp.n_streams, a0 + g.carry_rows); ++r) {
                const double* __restrict__ hist = static_cast<const double*>(cp.hist) + (int64_t)r * cp.hist_stride;
                const double* __restrict__ rrow = ring + (int64_t)r * ring_stride;
                double* __restrict__ ho = static_cast<double*>(cp.hist_out) + (int64_t)r * cp.hist_out_stride;
                for (int i = ctid; i < cp.new_hist_len; i += 256) {
                    const int64_t v = (int64_t)cp.drop + i;  // index into hist_p ++ mid
                    double x = 0.0;
                    if (v < cp.hist_len) x = hist[v];
                    else if (v - cp.hist_len < cp.n_in) x = __ldcg(rrow + (v - cp.hist_len) % g.RS);
                    ho[i] = x;
                }
            }
            item_done(nullptr, ds, dk);
            continue;
        }
        if (kind == 1) {
            // =============================== x2 item ===============================
            const int sbase = a0 * 8, tile0 = a1, nt = a2;
            const int nq = g.nk + (MT - 1) * SH;
            double acc[MT][2];
            double Areg[WA];
            const double* __restrict__ aw = Bs + ((lane >> 2) % 2) * g.blen + BOFF + (lane & 3) - ((lane >> 2) / 2);
            for (int it = 0; it < nt; ++it) {
                const int buf = (int)(st % NST);
                const uint32_t k = st / NST;
                ++st;
                const double* __restrict__ Xs = stg + (size_t)buf * g.stage_elems;
                int jb0, len, a, wlen;
                u_tile_geom(sbase, tile0 + it, jb0, len, a, wlen);
                while (!mbar_try_wait(sfull + buf, k & 1u)) {
                }
                __syncwarp();
                const int npos_t = min(CH_TJ, cu.n_pos - jb0);
                const bool active = warp * MT * JT < npos_t;
                if (active) {
#pragma unroll
                    for (int b = 0; b < MT; ++b) acc[b][0] = acc[b][1] = 0.0;
                    const double* __restrict__ xw = Xs + (lane >> 2) * g.pitch_u + (lane & 3) + a + 4 * (warp * MT * SH);
                    fir_mma_warp_tiles<MT, SH>(acc, Areg, xw, aw, g.nk, nq, 0, nq);
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sempty + buf)) : "memory");
                if (active) {
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
            item_done(udone + c, ds, dk);  // (the generic -> async proxy fence for the TMA reads of the ring sits on the reader's side)
            continue;
        }
        // =============================== polyphase item: task = 8 outputs nf .. nf+7 ===============================
        {
            const int tile = a0, rgp = a1;
            int n0, n1, span_t;
            int64_t d_base;
            p_tile_geom(tile, n0, n1, d_base, span_t);
            const int64_t gi = d_base - cp.hist_len;
            const int rows_base = rgp * RB * g.nrb;
            const int nj = min(g.nrb, (cp.n_streams - rows_base + RB - 1) / RB);
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
                const int buf = (int)(st % NST);
                const uint32_t k = st / NST;
                ++st;
                const int row0 = rows_base + j * RB;
                const double* __restrict__ xs = stg + (size_t)buf * g.stage_elems;
                int apad, wlen;
                p_stage_is_bulk(row0, gi, span_t, apad, wlen);
                while (!mbar_try_wait(sfull + buf, k & 1u)) {
                }
                __syncwarp();
                if (nf < n1) {
                    double acc[NT8][2];
#pragma unroll
                    for (int t = 0; t < NT8; ++t) acc[t][0] = acc[t][1] = 0.0;
                    const double* __restrict__ bp = xs + (lane >> 2) * g.pitch_p + base + apad + (lane & 3);
                    poly_mma_stage<NK, NT8>(acc, A, bp, g.pitch_p, nks);
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sempty + buf)) : "memory");
                    if (live) {
#pragma unroll
                        for (int t = 0; t < NT8; ++t) {
                            const int64_t s0 = (int64_t)row0 + t * 8 + 2 * (lane & 3);
                            // streaming stores: the outputs should not push the ring out of L2
                            if (s0 < cp.n_streams) __stcs(static_cast<double*>(cp.out) + s0 * cp.out_stride + (nf + i), acc[t][0]);
                            if (s0 + 1 < cp.n_streams) __stcs(static_cast<double*>(cp.out) + (s0 + 1) * cp.out_stride + (nf + i), acc[t][1]);
                        }
                    }
                } else {
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sempty + buf)) : "memory");
                }
            }
            item_done(pdone + c, ds, dk);  // the stages of this item have been read: its ring samples may be overwritten
        }
    }
}

template <int NK, int RB, int NST>
bool launch_chain_t(const FusedCall& c, cudaStream_t s, ChainWs* wsp, const int variant, const bool dry) {
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
    // rows per polyphase item: the coefficient registers are gathered once per item, but a long item keeps its ring samples
    // alive for longer (the ring has to hold everything between the x2 item that wrote a sample and the last polyphase item
    // that reads it)
    static const int prows_env = [] { const char* e = gar::tune_env("GAR_CHAIN_PROWS"); return e ? std::atoi(e) : 128; }();
    g.nrb = 1;
    while (g.nrb < prows_env / RB && g.nrb * 2 <= n_rb) g.nrb *= 2;
    g.n_rg_p = (c.n_streams + RB * g.nrb - 1) / (RB * g.nrb);
    const size_t p_elems = (size_t)RB * g.pitch_p;
    // ---- x2 geometry: K1m's (launch_fir_mma_t<1, 2>, 8 warps x 4 tiles, whole window, two buffers) ----
    const int kpu = c.t1 + 3;
    g.nk = (kpu + 3) / 4;
    g.blen = (4 * g.nk + 3 + 5) & ~1;
    g.blen = ((g.blen + 7) & ~15) + 8;  // the two phase filters 8 doubles apart modulo 16 banks
    g.xlen_u = (CH_TJ - 1) + 4 * g.nk + 4 * 3 + 10;
    g.pitch_u = ((g.xlen_u + 15) / 16) * 16 + 4;
    g.stage_elems = (int32_t)std::max(p_elems, (size_t)8 * g.pitch_u);
    const size_t smem = 256 + (size_t)2 * g.blen * sizeof(double) + (size_t)NST * g.stage_elems * sizeof(double);
    if (smem > 113 * 1024) return false;  // two blocks per SM
    g.n_tiles_u = (c.np + CH_TJ - 1) / CH_TJ;
    g.n_rg_u = (c.n_streams + 7) / 8;

    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_cta = 2 * n_sm;
    // ---- chunking: R chunks of S intermediate samples per row within the L2 budget. What is in flight at any time is
    // (blocks) x (shared-memory stages) x (one 8-row x 256-sample tile) = 14 MB of intermediate samples on 148 SMs; the ring
    // holds that plus the chunks waiting for their polyphase items.
    static const int64_t ring_budget = [] { const char* e = gar::tune_env("GAR_CHAIN_RING_MB"); return (int64_t)(e ? std::atoi(e) : 40) << 20; }();
    static const int tpi_env = [] { const char* e = gar::tune_env("GAR_CHAIN_TPI"); return e ? std::atoi(e) : 2; }();
    static const int s_env = [] { const char* e = gar::tune_env("GAR_CHAIN_S"); return e ? std::atoi(e) : 0; }();
    const int64_t n_mid = 2 * (int64_t)c.np;
    g.H = ((g.pitch_p + 255) / 256) * 256;
    bool found = false;
    for (int S = 4096; S >= 1024; S >>= 1) {
        if (s_env && S != s_env) continue;
        g.S = S;
        g.R = (int)std::min<int64_t>(64, ring_budget / ((int64_t)S * c.n_streams * 8));
        if (g.R < 8 && S > 1024) continue;
        if (g.R < 6) break;
        found = true;
        break;
    }
    if (!found) return false;
    g.tiles_per_chunk = g.S / 256;
    g.tiles_per_item = std::max(1, std::min(tpi_env, g.tiles_per_chunk));
    g.items_u_chunk = g.n_rg_u * ((g.tiles_per_chunk + g.tiles_per_item - 1) / g.tiles_per_item);
    g.NC = (int)((n_mid + g.S - 1) / g.S);
    g.RS = g.R * g.S;
    g.R = std::min(g.R, std::max(8, g.NC / 2));
    g.RS = g.R * g.S;
    if (g.NC < 2 * g.R || c.new_hp > g.S || c.hp > g.S || g.H > g.S) return false;
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
    static const int ahead_env = [] { const char* e = gar::tune_env("GAR_CHAIN_AHEAD"); return e ? std::atoi(e) : 1; }();
    g.ahead = std::max(0, std::min(NST - 1, ahead_env));
    g.carry_rows = 8;
    g.n_carry_items = (c.n_streams + g.carry_rows - 1) / g.carry_rows;

    if (dry) return true;  // geometry only: the engine asks before it sizes its inter-stage buffers
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
    const int blocks = n_cta;
    chain_up2_poly_kernel<NK, RB, NST><<<(unsigned)blocks, 288, smem, s>>>(cu, cp, g, ws);
    count_launch();
    return true;
}

}  // namespace

// test hook (gar_debug_chain_tile_hi): the chunk -> tile assignment of the polyphase items, for the CPU tests of its invariants
int chain_debug_tile_hi(int S, int kp, int NC, int n_tiles_p, int c, int hp, int64_t L, int64_t at0, int64_t step, int n_out) {
    ChainGeom g{};
    g.S = S; g.kp = kp; g.NC = NC; g.n_tiles_p = n_tiles_p;
    return chain_p_tile_hi(g, c, hp, L, at0, step, n_out);
}

// 0: never, 1: every eligible call, 2 (default): eligible calls whose full-size intermediate buffer would exceed the engine's
// inter-stage memory budget (the engine decides; see Engine::run)
static int g_chain_mode = [] {
    const char* e = gar::tune_env("GAR_CHAIN_MODE");
    return e ? std::atoi(e) : 2;
}();
void set_chain_kernel(int mode) { g_chain_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
int chain_kernel_mode() { return g_chain_mode; }

// K5 dispatch: float64 batches of at least 32 lock-step rows whose x2 and polyphase stages both run on the tensor cores
// (the K1m / K3p domain), calls long enough for a few ring cycles. dry: eligibility only (no allocation, no launch).
bool launch_chain_up2_poly(const FusedCall& c, cudaStream_t s, ChainWs* ws, bool dry) {
    if (g_chain_mode == 0 || (!ws && !dry) || !tensor_fir_enabled() || !tiled_polyphase_enabled()) return false;
    if (c.in_f32 || c.out_f32 || c.n_streams < 32 || c.t1 < 16 || c.np <= 0 || c.n_out <= 0) return false;
    if ((int64_t)c.np * c.n_streams < (1 << 21) || c.L > 4096 || c.t2 > 1024) return false;
    if (2 * (int64_t)c.np > 0x7ffffff0LL) return false;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.0) || r > 8.0) return false;
    // coefficient registers for K <= 80 or <= 112; three stages of 16 rows (one stage also holds an 8-row x2 window)
    return launch_chain_t<20, 16, 3>(c, s, ws, 0, dry) || launch_chain_t<28, 16, 3>(c, s, ws, 1, dry);
}

}  // namespace gar
