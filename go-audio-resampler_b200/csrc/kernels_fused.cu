// kernels_fused.cu — x2 up-sampler + polyphase stage of one engine.Resampler in one launch (resampler.go:97-121,142-175):
// K4 (generic, one thread per output), K4r / K3r (rational ratios, lanes = periods, barrier-free item pipeline), the
// coefficient-tile builder, and launch_fused_up2_poly.
#include "device_common.cuh"

namespace gar {
namespace {

// Carried tails of a fused x2 -> polyphase call (one block per row): the x2 stage's new tail is a plain copy,
// the polyphase stage's new tail needs the last few intermediate samples, recomputed here with the same
// strictly sequential chain as the tile core (bit-identical in float64).
// v[g] of the x2 stage's virtual input hist_u ++ in, where `in` may hold float32 samples (FusedCall::in_f32)
template <typename T>
__device__ __forceinline__ T fused_vin(const FusedCall& c, const T* __restrict__ hist_u, const int64_t row, const int g) {
    if (g < 0) return T(0);
    if (g < c.hu) return hist_u[g];
    const int i = g - c.hu;
    if (i >= c.n_in) return T(0);
    if (sizeof(T) == 8 && c.in_f32) return (T)(static_cast<const float*>(c.in) + row * c.in_stride)[i];
    return (static_cast<const T*>(c.in) + row * c.in_stride)[i];
}

template <typename T>
__device__ __forceinline__ void fused_carry_tails_rt(const FusedCall& c, const int64_t row, T* scratch, const int scratch_cap) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
    {
        T* hu_out = static_cast<T*>(c.hist_u_out) + row * c.hist_u_out_stride;
        block_copy4(c.new_hu, [&](int i) { return fused_vin<T>(c, hist_u, row, c.drop_u + i); }, [&](int i, T v) { hu_out[i] = v; });
    }
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
        block_copy4(span, [&](int i) { return fused_vin<T>(c, hist_u, row, p_lo + i); }, [&](int i, T v) { xb[i] = v; });
        block_copy4(2 * c.t1, [&](int i) { return bank_u[i]; }, [&](int i, T v) { bk[i] = v; });
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
                    for (int t = 0; t < c.t1; ++t) acc = fma(fused_vin<T>(c, hist_u, row, (j >> 1) + t), bkg[t], acc);
                    v = acc;
                } else {
                    double acc = 0;
                    for (int t = 0; t < c.t1; ++t)
                        acc = fma((double)fused_vin<T>(c, hist_u, row, (j >> 1) + t), (double)bkg[t], acc);
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
                                                            const int bank_pitch /*0: read banks through L1*/,
                                                            const int rowcap /*> 0: gather the outputs' rows by TMA*/,
                                                            const int cpitch) {
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
    T* crow = pbank;                              // [rowcap][cpitch] the coefficient row of every output (optional)

    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const int tid = threadIdx.x;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
    const int n_mid = c.np * NF;

    pdl_trigger_if_small();
    if (tile == n_tiles) {  // ---- carried tails ----
        pdl_wait();
        fused_carry_tails_rt<T>(c, row, xs, xlen + hpf + MT);
        return;
    }

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    __syncthreads();

    // ---- 0. this block's outputs are known from the geometry alone. Streaming-size launches (rowcap > 0): warp 0 gathers the
    // coefficient row of EVERY output into shared memory with one TMA bulk copy per row, all in flight at once, while the
    // block stages its samples and runs the x2 core — the per-tap loads of the polyphase loop otherwise pay an L2 round trip
    // per unrolled group (the source view put 40 % of a chunk launch there). Otherwise the rows are prefetched into L1. ----
    const int64_t Lq = (int64_t)c.L << 16;
    const int64_t lo = tile == 0 ? 0 : (int64_t)c.hp + (int64_t)tile * ms;
    const int64_t hi = (int64_t)c.hp + (int64_t)(tile + 1) * ms;
    auto first_n = [&](const int64_t d) -> int64_t {  // smallest n with div_n >= d
        const int64_t need_at = d * Lq - c.at0;
        return need_at <= 0 ? 0 : (need_at + c.step - 1) / c.step;
    };
    const int64_t n_lo = min((int64_t)c.n_out, first_n(lo));
    const int64_t n_hi = tile == n_tiles - 1 ? (int64_t)c.n_out : min((int64_t)c.n_out, first_n(hi));
    const bool gather = rowcap > 0 && n_hi - n_lo <= rowcap;  // (INTERP: rows of the interleaved bank, [tap]{a,b,c,d}, always aligned)
    auto row_pad = [&](const T* rowp) -> int {  // elements between the 16-byte aligned address below the row and the row
        return (int)((reinterpret_cast<uintptr_t>(rowp) & 15u) / sizeof(T));
    };
    if (gather) {  // every thread issues the copies of its own outputs; each warp adds its byte count to the barrier
        uint32_t bytes = 0;
        for (int64_t n = n_lo + tid; n < n_hi; n += NT) {
            const int64_t full = (c.at0 + n * c.step) >> 16;
            const T* rowp = INTERP ? static_cast<const T*>(c.bank_il) + (full % c.L) * c.t2 * 4
                                   : static_cast<const T*>(c.bank_a) + (full % c.L) * c.t2;
            const int pad = INTERP ? 0 : row_pad(rowp);
            const uint32_t nb = INTERP ? (uint32_t)(c.t2 * 4 * sizeof(T)) : (uint32_t)(((c.t2 + pad) * sizeof(T) + 15) & ~(size_t)15);
            bulk_g2s(crow + (size_t)(n - n_lo) * cpitch, rowp - pad, nb, bar + 1);
            bytes += nb;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        if ((tid & 31) == 0 && bytes)
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"(bytes) : "memory");
    } else if (bank_pitch == 0 && n_hi - n_lo <= 4 * NT) {
        for (int64_t n = n_lo + tid; n < n_hi; n += NT) {
            const int64_t full = (c.at0 + n * c.step) >> 16;
            const int64_t co = (full % c.L) * c.t2;
            prefetch_row_l1(static_cast<const T*>(c.bank_a) + co, c.t2);
            if (INTERP) {
                prefetch_row_l1(static_cast<const T*>(c.bank_b) + co, c.t2);
                prefetch_row_l1(static_cast<const T*>(c.bank_c) + co, c.t2);
                prefetch_row_l1(static_cast<const T*>(c.bank_d) + co, c.t2);
            }
        }
    }

    pdl_wait();  // up to here: shared memory, geometry and the constant banks only

    // ---- 1. stage the x2 stage's input window ----
    const int p0 = (tile * ms) >> 1;  // first position of the tile (ms is even)
    const int tp = min(TP, c.np - p0);
    const int need = tp > 0 ? tp - 1 + c.t1 : 0;
    int a = 0;
    bool bulk = false;
    {
        const int gi = p0 - c.hu;
        if (gi >= 0 && need > 0 && !(sizeof(T) == 8 && c.in_f32)) {  // float32 samples are converted by guarded loads
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
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(words * sizeof(T)));
            bulk_g2s(xs, in + gi, (uint32_t)(words * sizeof(T)), bar);
        }
        for (int i = words + tid; i < xlen; i += NT) xs[i] = T(0);
    } else {
        block_copy4(xlen, [&](int i) { return i < need ? fused_vin<T>(c, hist_u, row, p0 + i) : T(0); }, [&](int i, T v) { xs[i] = v; });
    }
    block_copy4(NF * cp, [&](int i) {
        const int p = i / cp, k = i % cp - a;
        return (k >= 0 && k < c.t1) ? bank_u[p * c.t1 + k] : T(0);
    }, [&](int i, T v) { cs[i] = v; });
    // polyphase side: carried tail right-aligned in front of tile 0 (the tile itself starts 16-byte aligned),
    // optional a-bank copy (odd pitch: conflict-free rows)
    const int front = tile == 0 ? hpf : 0;
    if (tile == 0) block_copy4(c.hp, [&](int i) { return hist_p[i]; }, [&](int i, T v) { vp[hpf - c.hp + i] = v; });
    if (bank_pitch > 0) {
        const T* __restrict__ ba = static_cast<const T*>(c.bank_a);
#pragma unroll 8
        for (int i = tid; i < c.L * c.t2; i += NT) pbank[(i / c.t2) * bank_pitch + (i % c.t2)] = ba[i];  // 8 loads in flight
    }
    __syncthreads();
    if (gather && tid == 0)  // the single arrival: every warp's expect_tx above is ordered before it by the barrier
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
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
    // virtual vp index d lives at shared-memory element d - vbase
    const int64_t vbase = tile == 0 ? (int64_t)c.hp - hpf : lo;
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
    float* __restrict__ out32 = static_cast<float*>(c.out) + row * c.out_stride;  // FusedCall::out_f32
    const bool o32 = sizeof(T) == 8 && c.out_f32;
    const T* __restrict__ ga = static_cast<const T*>(c.bank_a);
    const T* __restrict__ gb = static_cast<const T*>(c.bank_b);
    const T* __restrict__ gc = static_cast<const T*>(c.bank_c);
    const T* __restrict__ gd = static_cast<const T*>(c.bank_d);
    (void)n_mid;
    if (gather) {
        while (!mbar_try_wait(bar + 1, 0)) {
        }
    }
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
            if (gather) {  // the gathered row of the interleaved bank: same operation order
                const T* __restrict__ sl = crow + (size_t)(n - n_lo) * cpitch;
                for (int k = 0; k < c.t2; ++k) {
                    const T coef = fma(x, fma(x, fma(x, sl[4 * k + 3], sl[4 * k + 2]), sl[4 * k + 1]), sl[4 * k]);
                    if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h[k], (double)coef, acc1);
                    else acc0 = fma((double)h[k], (double)coef, acc0);
                }
            } else {
                for (int k = 0; k < c.t2; ++k) {
                    const T coef = fma(x, fma(x, fma(x, gd[co + k], gc[co + k]), gb[co + k]), ga[co + k]);
                    if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h[k], (double)coef, acc1);
                    else acc0 = fma((double)h[k], (double)coef, acc0);
                }
            }
        } else {
            const T* __restrict__ ca = bank_pitch > 0 ? pbank + phase * bank_pitch : ga + (int64_t)phase * c.t2;
            if (gather) ca = crow + (size_t)(n - n_lo) * cpitch + row_pad(ca);
            for (int k = 0; k < c.t2; ++k) {
                if ((k & 1) && sizeof(T) == 4) acc1 = fma((double)h[k], (double)ca[k], acc1);
                else acc0 = fma((double)h[k], (double)ca[k], acc0);
            }
        }
        if (o32) out32[n] = (float)(acc0 + acc1);  // float32(v): constant.go:195-197
        else out[n] = (T)(acc0 + acc1);
    }
}

// =============================================================================================
// K4s — fused x2 up-sampler + polyphase stage for IRRATIONAL ratios on one / a few rows (BASELINE config 5b as a single
// stream; polyphase_stage.go:260-288 with cubic coefficient interpolation).
//
// With a fractional phase step no two outputs of a stream share their coefficients, and a thread-per-output kernel reads
// four bank rows per output through L1/L2 (4 scattered loads per FMA-tap: K4 measured 3.5 % of the FMA peak on config 5b).
// But outputs that share the INTEGER phase share the four rows a/b/c/d[phase] and differ only in the fraction x. So a
// block takes a large tile (about 15.5 outputs per phase), sorts its outputs by phase in shared memory (counting sort:
// histogram, scan, scatter), and a HALF-WARP processes one phase at a time: the phase's rows, interleaved as [tap]{a,b,c,d},
// are copied into the half-warp's shared-memory slot with 16-byte cp.async (from a pre-interleaved copy of the banks), the
// lanes are the outputs of that phase, and per tap every lane does 2 broadcast LDS.128 (coefficients) + 1 LDS.64 (sample)
// for 4 DFMAs: three Horner steps and the dot-product step, in exactly K4's operation order (bit-identical results).
// The x2 stage runs as in K4 (register-tiled core into a shared-memory tile), in NPASS passes per tile because the tile
// is several times larger than one pass of NT*R positions; the coefficient slots alias the x2 stage's dead input window.
// =============================================================================================
struct SortGeom {
    int32_t n_tiles, ms, cp, xlen, hpf, npass, ocap;
    const void* bank_il;  // [L][t2][4] = a,b,c,d per tap
};

template <int R, int NT>
__global__ void __launch_bounds__(NT, 1) fused_up2_poly_sorted_kernel(const FusedCall c, const SortGeom g) {
    using T = double;
    using V = double2;
    constexpr int VEC = 2, NF = 2;
    constexpr int TP = NT * R;    // positions per pass
    constexpr int MT = TP * NF;   // intermediate samples per pass
    constexpr int NHW = NT / 16;  // half-warps per block

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* cs = reinterpret_cast<T*>(smem_raw + 16);  // [2][cp]
    T* xs = cs + NF * g.cp;                       // [xlen]
    T* vp = xs + g.xlen;                          // [hpf + npass*MT]
    T* slots = vp + g.hpf + g.npass * MT;         // [NHW][2][t2*4]: the rows of two adjacent phases per half-warp
    int* hist = reinterpret_cast<int*>(slots + (size_t)NHW * 2 * c.t2 * 4);  // [L]
    int* start = hist + c.L;                                                 // [L + 1]
    int* cursor = start + c.L + 1;                                           // [L]
    unsigned short* key = reinterpret_cast<unsigned short*>(cursor + c.L);   // [ocap] phase of output i (unsorted)
    unsigned short* s_ph = key + g.ocap;                                     // sorted: phase,
    unsigned short* s_n = s_ph + g.ocap;                                     //   output index - n_lo,
    unsigned short* s_div = s_n + g.ocap;                                    //   window start - vbase,
    unsigned short* s_x = s_div + g.ocap;                                    //   fraction bits

    const int n_tiles = g.n_tiles, ms = g.ms, hpf = g.hpf;
    const int tile = blockIdx.x % (n_tiles + 1);
    const int64_t row = blockIdx.x / (n_tiles + 1);
    const int tid = threadIdx.x;
    const T* __restrict__ hist_u = static_cast<const T*>(c.hist_u) + row * c.hist_u_stride;
    const T* __restrict__ in = static_cast<const T*>(c.in) + row * c.in_stride;
    const T* __restrict__ hist_p = static_cast<const T*>(c.hist_p) + row * c.hist_p_stride;
    const T* __restrict__ bank_u = static_cast<const T*>(c.bank_u);
    pdl_wait();  // launched as a programmatic dependent: its blocks may be resident before the previous launch has drained

    if (tile == n_tiles) {  // ---- carried tails ----
        fused_carry_tails_rt<T>(c, row, xs, g.xlen + hpf + g.npass * MT);
        return;
    }

    // ---- 1. stage the x2 stage's input window of the whole tile ----
    const int n_mid = c.np * NF;
    const int m0 = tile * ms;                             // first intermediate sample of the tile (even)
    const int need_mid = min(ms + c.t2 - 1, n_mid - m0);  // intermediate samples the tile's outputs read
    const int p0 = m0 >> 1;
    const int need_pos = (need_mid + 1) >> 1;
    const int npass = (need_pos + TP - 1) / TP;
    const int need = need_pos > 0 ? need_pos - 1 + c.t1 : 0;
    int a = 0;
    bool bulk = false;
    {
        const int gi = p0 - c.hu;
        if (gi >= 0 && need > 0) {
            const uintptr_t addr = reinterpret_cast<uintptr_t>(in + gi);
            const int mis = (int)((addr & 15u) / sizeof(T));
            const int words = ((need + mis + VEC - 1) / VEC) * VEC;
            if (gi - mis >= 0 && gi - mis + words <= c.n_in && words <= g.xlen) {
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
        for (int i = words + tid; i < g.xlen; i += NT) xs[i] = T(0);
    } else {
        for (int i = tid; i < g.xlen; i += NT) xs[i] = i < need ? vload(hist_u, c.hu, in, c.n_in, p0 + i) : T(0);
    }
    for (int i = tid; i < NF * g.cp; i += NT) {
        const int p = i / g.cp, k = i % g.cp - a;
        cs[i] = (k >= 0 && k < c.t1) ? bank_u[p * c.t1 + k] : T(0);
    }
    const int front = tile == 0 ? hpf : 0;
    if (tile == 0)
        for (int i = tid; i < c.hp; i += NT) vp[hpf - c.hp + i] = hist_p[i];

    // ---- 2. phase histogram of the tile's outputs (independent of the samples: overlaps the bulk copy) ----
    const int64_t Lq = (int64_t)c.L << 16;
    const int64_t lo = tile == 0 ? 0 : (int64_t)c.hp + (int64_t)tile * ms;
    const int64_t hi = (int64_t)c.hp + (int64_t)(tile + 1) * ms;
    auto first_n = [&](const int64_t d) -> int64_t {  // smallest n with div_n >= d
        const int64_t need_at = d * Lq - c.at0;
        return need_at <= 0 ? 0 : (need_at + c.step - 1) / c.step;
    };
    const int64_t n_lo = min((int64_t)c.n_out, first_n(lo));
    const int64_t n_hi = tile == n_tiles - 1 ? (int64_t)c.n_out : min((int64_t)c.n_out, first_n(hi));
    const int cnt = (int)(n_hi - n_lo);  // <= ocap by construction (launcher)
    const int64_t vbase = tile == 0 ? (int64_t)c.hp - hpf : lo;
    for (int i = tid; i < c.L; i += NT) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += NT) {
        const int64_t full = (c.at0 + (n_lo + i) * c.step) >> 16;
        const int ph = (int)(full % c.L);
        key[i] = (unsigned short)ph;
        atomicAdd(&hist[ph], 1);
    }
    if (bulk) {
        while (!mbar_try_wait(bar, 0)) {
        }
    }

    // ---- 3. x2 FIR core -> intermediate tile in shared memory ----
    for (int q = 0; q < npass; ++q) {
        T res[R][NF];
        fir_tile_accumulate<T, 1, NF, R>(xs + R * (q * NT + tid), cs, g.cp, c.t1, a, res);
        T* mp = vp + front + (size_t)R * NF * (q * NT + tid);
#pragma unroll
        for (int w = 0; w < R * NF / VEC; ++w) reinterpret_cast<V*>(mp)[w] = vec_pack(&res[0][0] + w * VEC);
    }
    __syncthreads();  // tile complete, histogram complete

    // ---- 4. counting sort by phase: sorted position -> (phase, output, window start, fraction) ----
    {   // exclusive prefix sum of the histogram: warp scans + a scan of the warp totals (the serial form — thread p adding up
        // hist[0 .. p) — was ~200 dependent shared-memory loads long for the last phases)
        __shared__ int wtot[NT / 32 + 1];
        const int lane = tid & 31, wid = tid >> 5;
        int carry = 0;
        for (int b0 = 0; b0 < c.L; b0 += NT) {
            const int p = b0 + tid;
            const int v = p < c.L ? hist[p] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wtot[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                const int w = lane < NT / 32 ? wtot[lane] : 0;
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += t;
                }
                if (lane < NT / 32) wtot[lane] = wi - w;  // exclusive offset of warp `lane`
                if (lane == 31) wtot[NT / 32] = wi;       // total of this round
            }
            __syncthreads();
            const int excl = carry + wtot[wid] + incl - v;
            if (p < c.L) {
                start[p] = excl;
                cursor[p] = excl;
            }
            carry += wtot[NT / 32];
            __syncthreads();
        }
        if (tid == 0) start[c.L] = carry;
    }
    __syncthreads();
    for (int i = tid; i < cnt; i += NT) {
        const int ph = key[i];
        const int64_t at = c.at0 + (n_lo + i) * c.step;
        const int64_t div = ((at >> 16) - ph) / c.L;  // exact: (full - phase) is a multiple of L
        const int pos = atomicAdd(&cursor[ph], 1);
        s_ph[pos] = (unsigned short)ph;
        s_n[pos] = (unsigned short)i;
        s_div[pos] = (unsigned short)(div - vbase);
        s_x[pos] = (unsigned short)(at & 0xFFFF);
    }
    __syncthreads();

    // ---- 5. groups of 16 phase-sorted outputs per half-warp; the rows of two adjacent phases per slot load ----
    const int hw = tid >> 4, l16 = tid & 15;
    const unsigned hmask = 0xFFFFu << (tid & 16);
    T* slot = slots + (size_t)hw * 2 * c.t2 * 4;
    const T* __restrict__ il = static_cast<const T*>(g.bank_il);
    T* __restrict__ out = static_cast<T*>(c.out) + row * c.out_stride;
    const int n_groups = (cnt + 15) >> 4;
    for (int gi = hw; gi < n_groups; gi += NHW) {
        const int j = gi * 16 + l16;
        const bool active = j < cnt;
        const int my_ph = active ? (int)s_ph[j] : -1;
        const int pf = s_ph[gi * 16], pl = s_ph[min(cnt, gi * 16 + 16) - 1];
        T x = 0;
        const T* h = vp;
        int64_t n = 0;
        if (active) {
            x = (T)(int)s_x[j] * (T)(1.0 / 65536.0);
            h = vp + s_div[j];
            n = n_lo + s_n[j];
        }
        for (int pw = pf; pw <= pl; pw += 2) {
            const int chunks16 = min(2, pl - pw + 1) * c.t2 * 2;  // 16-byte pieces: rows pw (and pw + 1) are contiguous
            const T* src = il + (size_t)pw * c.t2 * 4;
            for (int ch = l16; ch < chunks16; ch += 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + ch * 2)), "l"(src + ch * 2) : "memory");
            cp_async_wait_all();
            __syncwarp(hmask);
            if (my_ph == pw || my_ph == pw + 1) {
                const T* sl = slot + (size_t)(my_ph - pw) * c.t2 * 4;
                double acc = 0;
#pragma unroll 4
                for (int k = 0; k < c.t2; ++k) {
                    const V ab = *reinterpret_cast<const V*>(sl + 4 * k);
                    const V cd = *reinterpret_cast<const V*>(sl + 4 * k + 2);
                    const T coef = fma(x, fma(x, fma(x, cd.y, cd.x), ab.y), ab.x);
                    acc = fma(h[k], coef, acc);
                }
                out[n] = acc;
            }
            __syncwarp(hmask);  // every lane is done with the slot before it is overwritten
        }
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
    pdl_wait();  // launched as a programmatic dependent (its coefficient tile may come from the launch just before it)

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
            // (the constant-bank / uniform-register form of the core, fir_tile_accumulate_uc, needs warp-convergent code:
            //  inside this work-queue loop ptxas keeps the taps in vector registers and nothing is gained — measured)
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

}  // namespace

// A/B toggle (GAR_NO_RAT=1 or set_tiled_polyphase(false)): fall back to the one-thread-per-output kernels (no K4r / K3r / K3i)
static bool g_fused_rat = [] {
    const char* e = gar::tune_env("GAR_NO_RAT");
    return !(e && e[0] && e[0] != '0');
}();
void set_tiled_polyphase(bool on) { g_fused_rat = on; }


bool tiled_polyphase_enabled() { return g_fused_rat; }

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
    if (!INTERP && (int64_t)n_tiles * c.n_streams >= 64) {  // (64: a single 10 s stream has 288 tiles — it ran 4x slower through L1)
        const int pitch = c.t2 | 1;
        if (((size_t)c.L * pitch + words) * sizeof(T) + 16 <= 100 * 1024) {
            bank_pitch = pitch;
            words += (size_t)c.L * pitch;
        }
    }
    // streaming-size launches: one TMA bulk copy per output gathers its coefficient row (rows padded to 16-byte multiples,
    // pitch = 2 elements of 8 bytes mod 16 so that the lanes' rows start in different banks)
    int rowcap = 0, cpitch = 0;
    // (measured, warm caches: a Flush launch of one tile 14.3 -> 12.0 us; a 4096-frame chunk of 18 tiles is better off with the
    //  L1 prefetch, 12.2 against 13.3 us, so the gather is for launches of a few tiles only)
    if (!INTERP && bank_pitch == 0 && (int64_t)n_tiles * c.n_streams <= 4) {
        const double r = (double)c.step / ((double)c.L * 65536.0);
        const int cap = (int)((double)(MT + c.hp + 2) / r) + 4;
        int pitch = ((c.t2 + VEC + VEC - 1) / VEC) * VEC;  // room for the alignment pad
        while ((pitch * (int)sizeof(T)) % 128 != 16) pitch += VEC;
        if ((words + (size_t)cap * pitch) * sizeof(T) + 16 <= 200 * 1024) {
            rowcap = cap;
            cpitch = pitch;
            words += (size_t)cap * pitch;
        }
    }
    if (INTERP && sizeof(T) == 8 && c.bank_il && (int64_t)n_tiles * c.n_streams <= 4) {
        // interpolated coefficients: rows of the interleaved bank (32 bytes per tap); as many outputs as 128 KB of rows hold —
        // a Flush has a few dozen to a few hundred (a tile with more reads its coefficients through L1 as before)
        int pitch = 4 * c.t2;
        while ((pitch * 8) % 128 != 16) pitch += 2;
        const int cap = (int)((128 * 1024) / ((size_t)pitch * 8));
        if (cap >= 16 && (words + (size_t)cap * pitch) * sizeof(T) + 16 <= 200 * 1024) {
            rowcap = cap;
            cpitch = pitch;
            words += (size_t)cap * pitch;
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
    launch_pdl(k, (unsigned)blocks, (unsigned)NT, smem, s, c, n_tiles, ms, cp, xlen, hpf, bank_pitch, rowcap, cpitch);
    count_launch();
    return true;
}

// K4s launcher: float64, interpolated coefficients, fewer rows than the tensor-core batch kernels take.
template <int NT>
static bool launch_fused_sorted_nt(const FusedCall& c, const void* bank_il, cudaStream_t s) {
    constexpr int R = 6, VEC = 2;
    constexpr int NCH = (R - 1 + VEC - 1) / VEC + 1;
    constexpr int MT = NT * R * 2;
    const double r = (double)c.step / ((double)c.L * 65536.0);  // intermediate samples per output
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_mid = c.np * 2;
    const int hpf = (c.hp + VEC - 1) / VEC * VEC;
    const int cp = ((c.t1 + VEC - 1 + VEC - 1) / VEC) * VEC;
    auto geom = [&](int ms, SortGeom& g) -> size_t {
        g.ms = ms;
        g.npass = ((ms + c.t2) / 2 + NT * R - 1) / (NT * R);
        g.xlen = R * (g.npass * NT - 1) + (cp / VEC + NCH + 1) * VEC;
        g.ocap = ((int)((double)(ms + c.hp + 2) / r) + 4 + 7) & ~7;  // tile 0 also covers the carried tail
        g.n_tiles = (n_mid + ms - 1) / ms;
        return 16 + ((size_t)2 * cp + g.xlen + hpf + (size_t)g.npass * MT + (size_t)(NT / 16) * 2 * c.t2 * 4) * sizeof(double) +
               ((size_t)3 * c.L + 1) * sizeof(int) + (size_t)5 * g.ocap * sizeof(unsigned short) + 16;
    };
    // One block per SM (the block is 16 or 8 warps wide). Tile = an even share of the row per SM, so that one wave of
    // blocks covers the call; at least ~4 outputs per phase (fewer leave the half-warp groups spread over many phases),
    // at most what fits shared memory — then several balanced waves.
    SortGeom g{};
    g.cp = cp; g.hpf = hpf; g.bank_il = bank_il;
    const int64_t total_mid = (int64_t)n_mid * c.n_streams;
    const int slots = std::max(1, sms - c.n_streams);  // every row adds one carry block to the wave
    int ms = (int)((total_mid + slots - 1) / slots);
    ms = std::max(ms, (int)(4.0 * c.L * r));
    ms = std::max((ms + 1) & ~1, 8 * c.t2);
    size_t smem = geom(ms, g);
    if (smem > 227 * 1024) {
        int lo = 8 * c.t2, hi = ms;  // largest tile that fits
        if (geom(lo, g) > 227 * 1024) return false;
        while (hi - lo > 2) {
            const int mid = ((lo + hi) / 2) & ~1;
            if (geom(mid, g) <= 227 * 1024) lo = mid; else hi = mid;
        }
        // balanced waves: the tile count of every row rounded up to a whole number of waves
        int tiles = (n_mid + lo - 1) / lo;
        const int64_t waves = ((int64_t)(tiles + 1) * c.n_streams + sms - 1) / sms;
        tiles = (int)std::max<int64_t>(tiles, waves * sms / c.n_streams - 1);
        ms = std::max(((n_mid + tiles - 1) / tiles + 1) & ~1, 8 * c.t2);
        if (ms > lo) ms = lo;
        smem = geom(ms, g);
    }
    if (g.ocap > 65535 || smem > 227 * 1024 || hpf + g.npass * MT + c.hp > 65535) return false;
    auto k = fused_up2_poly_sorted_kernel<R, NT>;
    static size_t configured[64] = {0};
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    const int64_t blocks = (int64_t)(g.n_tiles + 1) * c.n_streams;
    launch_pdl(k, (unsigned)blocks, (unsigned)NT, smem, s, c, g);
    count_launch();
    return true;
}

static bool launch_fused_sorted(const FusedCall& c, const void* bank_il, cudaStream_t s) {
    if (!bank_il || c.np <= 0 || c.n_out <= 0 || c.hp > 1024 || c.L > 512 || c.t2 > 200 || c.t2 < 2) return false;
    const double r = (double)c.step / ((double)c.L * 65536.0);
    if (!(r > 0.05) || r > 16.0) return false;
    // 16 warps per block when the coefficient slots of 32 half-warps leave room for a useful tile, else 8 warps
    if ((size_t)32 * 2 * c.t2 * 32 <= 100 * 1024 && launch_fused_sorted_nt<512>(c, bank_il, s)) return true;
    return launch_fused_sorted_nt<256>(c, bank_il, s);
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
        const char* e = gar::tune_env("GAR_RAT_OPT");
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
        // A call of only a few blocks per SM (one long row: 8k->192k's last stage is 1810 tiles) is governed by quantisation:
        // its time is the number of tiles the busiest SM gets = ceil(blocks / SMs) x tiles per block (two blocks share an
        // SM; a lone block hides latency worse: +15 %). 604 three-tile blocks: 5 x 3 = 15 tile times; 259 seven-tile
        // blocks: 2 x 7 = 14. (44.1k->48k as one 10 s row, 375 tiles, stays at one tile per block: 3 x 1.)
        if (total_tiles <= slots * 24) {
            const int64_t nsm = slots / 2;
            double best = 1e30;
            for (int64_t t = 1; t <= std::min<int64_t>(RAT_MAXT, g.n_tiles); ++t) {
                const int64_t groups = ((g.n_tiles + t - 1) / t + 1) * c.n_streams;  // + the carry block of every row
                const int64_t per_sm = (groups + nsm - 1) / nsm;
                const double cost = (double)per_sm * (double)t * (per_sm < 2 ? 1.15 : 1.0);
                if (cost < best - 1e-9) {
                    best = cost;
                    tpb = t;
                }
            }
        }
        g.tiles_per_block = (int32_t)tpb;
        g.n_groups = (g.n_tiles + g.tiles_per_block - 1) / g.tiles_per_block;
    }
    if (const char* e = gar::tune_env("GAR_RAT_TPB")) {  // tuning override
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
                cudaDeviceSynchronize();  // tiles may still be read by launches on any stream of this handle
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
    launch_pdl(k, (unsigned)blocks, (unsigned)nthreads, smem, s, c, g);
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

bool launch_rat_poly_only_f64(const FusedCall& c, cudaStream_t s, RatCache* cache) { return launch_rat<double, false>(c, s, cache); }

// float64 x2 -> polyphase pairs that run as the two tensor-core launches (K1m + K3m / K3p) rather than as one fused launch
bool up2_poly_runs_as_tensor_pair(const FusedCall& c) {
    const bool rational = !c.interp && ((c.step | c.at0) & 0xFFFF) == 0;
    static const int rational_min_rows = [] { const char* e = gar::tune_env("GAR_TENSOR_MIN_ROWS"); return e ? std::atoi(e) : 32; }();
    return tensor_fir_enabled() && g_fused_rat && (int64_t)c.np * c.n_streams >= 32768 &&
           c.n_streams >= (rational ? rational_min_rows : 8);
}

const char* launch_fused_up2_poly(const FusedCall& c, int dtype, cudaStream_t s, RatCache* cache) {
    if (c.n_streams <= 0) return "none";
    if (dtype == DT_F64 && (c.in_f32 || c.out_f32)) {  // float32 I/O folded into the kernel: the generic K4 only
        if (c.interp) return launch_fused_t<double, true>(c, s) ? "fused_up2_poly_f64_interp_io32" : nullptr;
        return launch_fused_t<double, false>(c, s) ? "fused_up2_poly_f64_io32" : nullptr;
    }
    if (dtype == DT_F32) {
        if (c.interp) return launch_fused_t<float, true>(c, s) ? "fused_up2_poly_f32_interp" : nullptr;
        return launch_fused_t<float, false>(c, s) ? "fused_up2_poly_f32" : nullptr;
    }
    // Tensor-core path = two launches: the x2 stage (K1m) and the polyphase stage (K3m) on the FP64 tensor cores.
    //  * rational ratios: batches of >= 32 lock-step rows. Measured on 21 M input samples of 44.1k->48k (TFLOP/s, tensor path
    //    against the fused K4r): 8 rows 9.4 / 17.6, 16 rows 14.6 / 17.6, 24 rows 17.7 / 17.6, 32 rows 19.9 / 17.8, 48 rows
    //    20.4 / 17.7, 64 rows 23.0, 256 rows x 10 s 28.1 — K3m / K3p amortise their coefficient matrices over the rows, K4r
    //    is the better kernel for a few long rows (GAR_TENSOR_MIN_ROWS overrides the threshold);
    //  * irrational ratios: batches of >= 8 rows (K3m / K3i evaluate the interpolated coefficients once per batch; the fused
    //    one-thread-per-output kernel is 10x slower there).
    if (up2_poly_runs_as_tensor_pair(c)) return nullptr;
    if (!c.interp && launch_rat<double, true>(c, s, cache)) return "fused_up2_rat_f64";
    // a large lock-step batch that the rational kernel does not cover runs as two launches: the stand-alone x2 kernel and
    // K3i (lanes = rows, interpolated coefficients evaluated once per batch) beat the one-thread-per-output fused kernel
    {
        const double r = (double)c.step / ((double)c.L * 65536.0);
        if (g_fused_rat && c.n_streams >= 8 && (int64_t)c.n_out * c.n_streams >= 16384 && r > 0.0 && r <= 8.0 &&
            c.L <= 4096 && c.t2 <= 1024)
            return nullptr;
    }
    // irrational ratio on one / a few rows: phase-sorted tiles (K4s) once the call is large enough to fill them
    if (c.interp && g_fused_rat && c.bank_il && (int64_t)c.n_out >= 16 * (int64_t)c.L &&
        launch_fused_sorted(c, c.bank_il, s))
        return "fused_up2_poly_sorted_f64";
    if (c.interp && c.bank_il && c.n_streams <= 4 && (int64_t)c.np * 2 <= 512 && c.n_out > 88) {
        // A Flush-size call with interpolated coefficients and more outputs than the fused kernel can gather coefficient rows for
        // (its shared memory holds ~90 rows of the interleaved bank beside the tile): two launches — the x2 kernel, then the
        // stand-alone polyphase kernel, which gathers the rows of 128 outputs per block — are faster than one launch that reads
        // four scattered coefficients per tap (config 5b's Flush: 6.6 + 5.7 against 21.4 us)
        return nullptr;
    }
    if (c.interp) return launch_fused_t<double, true>(c, s) ? "fused_up2_poly_f64_interp" : nullptr;
    return launch_fused_t<double, false>(c, s) ? "fused_up2_poly_f64" : nullptr;
}


}  // namespace gar
