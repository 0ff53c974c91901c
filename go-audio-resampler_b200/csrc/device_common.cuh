// device_common.cuh — device helpers shared by the kernel translation units (vector pack/unpack, the virtual input
// v = hist ++ in, mbarrier / TMA bulk copy / cp.async wrappers, the register-tiled FIR core, the DMMA wrapper) and the
// few host-side pieces they share (launch counter, A/B switches, tiles-per-block heuristic).
#pragma once
#include <type_traits>
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace gar {

// host side, defined in kernels_misc.cu / kernels_fir.cu / kernels_fused.cu
void count_launch();
bool tensor_fir_enabled();        // K1m / K2m / K3m (FP64 tensor cores)
bool tiled_polyphase_enabled();   // K4r / K3r / K3i / K3m
int pick_tiles_per_block(int n_tiles, int n_streams, int* n_groups);
// K3r: the rational-ratio kernel without the x2 stage (kernels_fused.cu), used by launch_poly
bool launch_rat_poly_only_f64(const FusedCall& c, cudaStream_t s, RatCache* cache);

// Programmatic dependent launch (sm_90+): a kernel launched with this attribute may start its blocks while the previous
// kernel of the stream is still running (as soon as that kernel's blocks have called pdl_trigger() or exited); it must call
// pdl_wait() before it touches anything the previous kernel reads or writes. The streaming-size kernels use it to overlap
// their prologue (barrier set-up, geometry, TMA gather / prefetch of the constant coefficient banks, instruction fetch)
// with the tail of the launch before them: a Flush of the 8k->192k pipeline is ten dependent launches of a few microseconds.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

namespace {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// only small grids let their successor in early: its waiting blocks would otherwise take slots from this kernel's own blocks
__device__ __forceinline__ void pdl_trigger_if_small() {
    if (gridDim.x <= 296) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

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

// Pull the 128-byte lines of row[0 .. n) into L1 (streaming-size launches: the per-tap coefficient loads of the polyphase
// loops otherwise pay one L2 round trip per unrolled group; issued before the staging / FIR phases they land in time)
template <typename T>
__device__ __forceinline__ void prefetch_row_l1(const T* row, const int n) {
    const uintptr_t b = reinterpret_cast<uintptr_t>(row) & ~(uintptr_t)127, e = reinterpret_cast<uintptr_t>(row + n);
    for (uintptr_t q = b; q < e; q += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
}

// dst(i) = src(i) for i in [0, n), block-wide, FOUR loads in flight per thread before the first store: a plain
// `for (i = tid; ...) dst[i] = load(i)` issues load, waits, stores, loads again — one full memory round trip per
// iteration, which is what a streaming-size launch (a few hundred elements per block) spends most of its time on.
template <class Load, class Store>
__device__ __forceinline__ void block_copy4(const int n, Load load, Store store) {
    const int nt = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * nt) {
        decltype(load(0)) v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * nt < n) v[j] = load(i0 + j * nt);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * nt < n) store(i0 + j * nt, v[j]);
    }
}

template <typename T>
__device__ __forceinline__ void carry_row(const T* hist, int hist_len, const T* in, int n_in, T* hist_out, int drop,
                                          int new_len) {
    block_copy4(new_len, [&](int i) { return vload(hist, hist_len, in, n_in, drop + i); },
                [&](int i, T v) { hist_out[i] = v; });
}

// calls f(integral_constant<N>) for the N in [LO, HI] equal to n, f(integral_constant<0>) when n is outside the range
template <int HI, int LO, class F>
__device__ __forceinline__ void dispatch_count(const int n, F&& f) {
    if (n == HI) f(std::integral_constant<int, HI>{});
    else if constexpr (HI > LO) dispatch_count<HI - 1, LO>(n, f);
    else f(std::integral_constant<int, 0>{});
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


// ---- float64 FIR core with the taps in the CONSTANT bank (kernel parameters) --------------------------------------------
// Every lane of a warp multiplies by the same tap, so when the filter travels as a kernel parameter ptxas loads each tap into a
// UNIFORM register (LDCU.64 c[0x0][UR+imm]) and issues DFMA R, R, UR, R: two register-file operands instead of three, and no
// coefficient LDS at all. tools/probe_dfma_const.cu: 35.6 against 29.7 TFLOP/s for this core's operand pattern.
// Layout: c[p][i] = bank[p][i - 2] (two leading zeros absorb the alignment pad a in {0, 1}), zero after the last tap.
template <int NFM, int CPD>
struct FirTapsD {
    double c[NFM][CPD];
};
constexpr int FIR_TAPS_CPD = 288;  // taps per phase + 6 must fit (the quality presets need at most 267 + pad)

// host side: fill the parameter block from a host copy of the bank ([nf][taps] doubles); false when it does not fit
template <int NFM, int CPD>
inline bool fill_fir_taps(FirTapsD<NFM, CPD>& P, const double* bank, int nf, int taps) {
    if (!bank || nf > NFM || taps + 6 > CPD) return false;
    for (int p = 0; p < NFM; ++p)
        for (int i = 0; i < CPD; ++i) {
            const int k = i - 2;
            P.c[p][i] = (p < nf && k >= 0 && k < taps) ? bank[(size_t)p * taps + k] : 0.0;
        }
    return true;
}

// Same arithmetic and the same summation order as fir_tile_accumulate<double, M, NF, R> (bit-identical results). The tap
// cursors cur[p] walk the parameter bank on their own so that their address arithmetic stays in uniform registers.
template <int M, int NF, int R, int NFM, int CPD>
__device__ __forceinline__ void fir_tile_accumulate_uc(const double* __restrict__ xt, const FirTapsD<NFM, CPD>& P,
                                                       const int taps, const int a, double (&res)[R][NF]) {
    constexpr int VEC = 2;
    constexpr int NCH = (M * (R - 1) + VEC - 1) / VEC + 1;
    constexpr int WREG = NCH * VEC;
    const int n_iter = (taps + a + VEC - 1) / VEC;
    double xr[WREG];
    double acc[R][NF];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < NF; ++p) acc[r][p] = 0.0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) vec_unpack(*reinterpret_cast<const double2*>(xt + ch * VEC), xr + ch * VEC);
    const double* cur[NF];
#pragma unroll
    for (int p = 0; p < NF; ++p) cur[p] = &P.c[p][2 - a];
    auto step = [&](const int u, const int it) {
        double cv[NF][VEC];
#pragma unroll
        for (int p = 0; p < NF; ++p) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) cv[p][i] = cur[p][i];
            cur[p] += VEC;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const double xv = xr[(u * VEC + i + M * r) % WREG];
#pragma unroll
                for (int p = 0; p < NF; ++p) acc[r][p] = fma(xv, cv[p][i], acc[r][p]);
            }
        vec_unpack(*reinterpret_cast<const double2*>(xt + (it + NCH) * VEC), xr + (u % NCH) * VEC);
    };
    int it0 = 0;
    for (; it0 + NCH <= n_iter; it0 += NCH) {
#pragma unroll
        for (int u = 0; u < NCH; ++u) step(u, it0 + u);
    }
#pragma unroll
    for (int u = 0; u < NCH; ++u)
        if (it0 + u < n_iter) step(u, it0 + u);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < NF; ++p) res[r][p] = acc[r][p];
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}


}  // namespace
}  // namespace gar
