// mma_cores.cuh — the per-warp FP64 tensor-core cores (PTX mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) of the batched kernels, as
// inlined device functions so that the persistent chain kernel (K5, kernels_chain.cu) runs EXACTLY the arithmetic of the
// stand-alone kernels K1m (fir_mma_f64_kernel, kernels_fir.cu) and K3p (poly_rows_pipe_kernel, kernels_poly.cu): same
// fragments, same order of accumulation, bit-identical results.
#pragma once
#include "device_common.cuh"

namespace gar {
namespace {

// ---- x2 / decimator FIR as a block-Toeplitz contraction (dft_stage.go:229-338, :499-551) -----------------------------------
// One warp owns MT consecutive MMA tiles; tile b sees the sample window shifted by b*SH k-steps, so ONE B fragment (xw[4*q])
// and ONE A fragment (aw[4*q], kept in the rotating window Areg) feed MT MMAs per step q. k-steps [qa, qb) of nq = nk +
// (MT-1)*SH; acc is cleared by the caller.
template <int MT, int SH>
__device__ __forceinline__ void fir_mma_warp_tiles(double (&acc)[MT][2], double (&Areg)[(MT - 1) * SH + 1],
                                                   const double* __restrict__ xw, const double* __restrict__ aw, const int nk,
                                                   const int nq, const int qa, const int qb) {
    constexpr int WA = (MT - 1) * SH + 1;
    constexpr int OFF = (MT - 1) * SH;
    // MODE 0: every step has all MT tiles inside the filter; MODE 2: the first group (q0 = 0), where which tiles have started
    // is known at compile time; MODE 1: run-time checks (ramp-down, short filters)
    auto steps = [&](const int q0, auto mode) {
        constexpr int MODE = decltype(mode)::value;
#pragma unroll
        for (int u = 0; u < WA; ++u) {
            const int q = MODE == 2 ? u : q0 + u;
            if (MODE != 1 || q < qb) {
                Areg[u] = (MODE != 1 || q < nk) ? aw[4 * q] : 0.0;
                const double bf = xw[4 * q];
#pragma unroll
                for (int b = 0; b < MT; ++b) {
                    const int kk = q - b * SH;
                    const bool on = MODE == 0 ? true : MODE == 2 ? u - b * SH >= 0 : (kk >= 0 && kk < nk);
                    if (on) dmma884(acc[b][0], acc[b][1], Areg[((u - b * SH) % WA + WA) % WA], bf);
                }
            }
        }
    };
    // ramp-down group with D = nk - q0 known at compile time
    auto steps_tail = [&](const int q0, auto d_tag) {
        constexpr int D = decltype(d_tag)::value - OFF;
#pragma unroll
        for (int u = 0; u < WA; ++u) {
            if (u < D + OFF) {
                Areg[u] = u < D ? aw[4 * (q0 + u)] : 0.0;
                const double bf = xw[4 * (q0 + u)];
#pragma unroll
                for (int b = 0; b < MT; ++b)
                    if (u - b * SH < D) dmma884(acc[b][0], acc[b][1], Areg[((u - b * SH) % WA + WA) % WA], bf);
            }
        }
    };
    const int q_steady = min(nk, qb) - WA;  // last q0 of an all-valid group
    for (int q0 = qa; q0 < qb; q0 += WA) {
        if (q0 >= OFF && q0 <= q_steady) steps(q0, std::integral_constant<int, 0>{});
        else if (q0 == 0 && q_steady >= 0) steps(q0, std::integral_constant<int, 2>{});
        else if (qb == nq && WA <= 25 && q0 >= OFF && nk - q0 + OFF >= 1 && nk - q0 < WA)
            dispatch_count<WA - 1 + OFF, 1>(nk - q0 + OFF, [&](auto t) {
                if constexpr (decltype(t)::value != 0) steps_tail(q0, t);
            });
        else steps(q0, std::integral_constant<int, 1>{});
    }
}

// ---- polyphase stage (polyphase_stage.go:186-312) as coefficient-matrix x row-window MMAs ------------------------------------
// Lane l of a warp task (8 adjacent outputs) keeps A[i = l/4][4*kk + l%4] of its output i for every k-step in registers:
// coef(phase, x)[k] = a + x(b + x(c + x d)) at k = 4*kk + l%4 - o_i, zero outside the filter (polyphase_stage.go:266-283).
template <int NK>
__device__ __forceinline__ void poly_gather_coeffs(double (&A)[NK], const PolyCall& c, const int ph, const int o_i,
                                                   const double x, const bool live, const int nks, const int lane) {
    const double* __restrict__ ga = static_cast<const double*>(c.bank_a);
    const double* __restrict__ gb = static_cast<const double*>(c.bank_b);
    const double* __restrict__ gc = static_cast<const double*>(c.bank_c);
    const double* __restrict__ gd = static_cast<const double*>(c.bank_d);
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
        const int k = 4 * kk + (lane & 3) - o_i;
        double v = 0.0;
        if (kk < nks && live && k >= 0 && k < c.taps) {
            const int co = ph * c.taps + k;
            v = ga[co];
            if (c.interp) v = fma(x, fma(x, fma(x, gd[co], gc[co]), gb[co]), v);
        }
        A[kk] = v;
    }
}

// acc[t] += A x (the window of rows 8t .. 8t+7 of a staged row block); bp = this lane's B-fragment base
template <int NK, int NT8>
__device__ __forceinline__ void poly_mma_stage(double (&acc)[NT8][2], const double (&A)[NK], const double* __restrict__ bp,
                                               const int pitch, const int nks) {
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
        if (kk < nks) {
#pragma unroll
            for (int t = 0; t < NT8; ++t) dmma884(acc[t][0], acc[t][1], A[kk], bp[t * 8 * pitch + 4 * kk]);
        }
    }
}

}  // namespace
}  // namespace gar
