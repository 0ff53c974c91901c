// kernels.cuh — launch descriptors shared by the host engine and the device kernels.
//
// Data model (B200-first): every primitive stage call sees a
// *virtual input*  v = hist ++ in  per stream row, where `hist` is the stage's
// carried tail (device resident, ping-pong buffered) and `in` is either the
// caller's chunk, the previous stage's output, or a shared zero row (flush).
// All integer geometry (how many outputs, where they start, what is carried) is
// computed on the host by the exact state machine of the reference and passed
// in these descriptors, so kernels are pure functions of (v, bank, descriptor).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace gar {

// out[(j*nf + p)] = sum_k v[first + j*stride + k] * bank[p*taps + k],  j < n_pos, p < nf
//   integer up-sampler  (dft_stage.go:156-338):  stride = 1, nf = factor
//   integer decimator   (dft_stage.go:488-554):  stride = M, nf = 1, first = decimPhase
struct FirCall {
    const void* hist;     int64_t hist_stride;  int32_t hist_len;
    const void* in;       int64_t in_stride;    int32_t n_in;        // in_stride == 0: shared (zero) row
    void* out;            int64_t out_stride;                          // already offset to the write position
    void* hist_out;       int64_t hist_out_stride;                     // new tail = v[drop .. hist_len+n_in)
    int32_t drop;         int32_t new_hist_len;
    const void* bank;     // device, [nf][taps]
    const float* bank_host_f32;  // optional host copy of a float32 bank: float32 decimators pass their taps as KERNEL
                                 // PARAMETERS (constant bank) so that the FMAs take them from uniform registers
    const double* bank_host_f64; // the same for float64 banks (vector-FMA FIR stages)
    int32_t taps;         int32_t stride;       int32_t nf;
    int32_t first;        int32_t n_pos;
    int32_t n_streams;    // rows processed by this launch (row r of every pointer = base + r*stride)
    int32_t in_f32;       // `in` / `out` hold float32 samples (strides in float32 elements); only the float64 tensor-core FIR
    int32_t out_f32;      // kernels (K1m / K2m) take them (fir_mma_io32_takes)
};

// out[n] = sum_k v[div_n + k] * (a + x(b + x(c + x d)))[phase_n][k]   (polyphase_stage.go:186-312)
//   at_n = at0 + n*step; full = at_n >> 16; div = full / L; phase = full % L; x = (at_n & 0xFFFF) / 65536
struct PolyCall {
    const void* hist;     int64_t hist_stride;  int32_t hist_len;
    const void* in;       int64_t in_stride;    int32_t n_in;
    void* out;            int64_t out_stride;
    void* hist_out;       int64_t hist_out_stride;
    int32_t drop;         int32_t new_hist_len;
    const void* bank_a;   const void* bank_b;   const void* bank_c;   const void* bank_d;  // [L][taps]
    const void* bank_il;  // optional: the four banks interleaved per tap, [L][taps][4] = a,b,c,d (flush-size launches gather rows of it)
    int32_t taps;         int32_t L;
    int64_t at0;          int64_t step;
    int32_t n_out;        int32_t interp;       // interp: (step & 0xFFFF) != 0 || (at0 & 0xFFFF) != 0
    int32_t n_streams;
    int32_t out_f32;      // `out` holds float32 samples (out_stride in float32 elements); only the tensor-core kernels (K3m / K3p) take it
};

// K4: x2 up-sampler fused with the polyphase stage that follows it inside one engine.Resampler
// (resampler.go:97-121,142-175). The intermediate-rate samples of a tile live only in shared memory.
//   mid[j]  = sum_t vu[(j>>1) + t] * bank_u[j&1][t]            vu = hist_u ++ in,      j < 2*np
//   out[n]  = sum_k vp[div_n + k] * coef(phase_n, x_n)[k]      vp = hist_p ++ mid
struct FusedCall {
    // x2 stage
    const void* hist_u;   int64_t hist_u_stride;   int32_t hu;
    const void* in;       int64_t in_stride;       int32_t n_in;
    void* hist_u_out;     int64_t hist_u_out_stride;
    int32_t drop_u;       int32_t new_hu;
    const void* bank_u;   int32_t t1;              int32_t np;       // positions; 2*np intermediate samples
    const double* bank_u_host;  // optional host copy of a float64 x2 bank: passed as kernel parameters (uniform-register taps)
    // polyphase stage
    const void* hist_p;   int64_t hist_p_stride;   int32_t hp;
    void* hist_p_out;     int64_t hist_p_out_stride;
    int32_t drop_p;       int32_t new_hp;
    const void* bank_a;   const void* bank_b;      const void* bank_c;   const void* bank_d;
    const void* bank_il;  // optional: the four banks interleaved per tap, [L][t2][4] = a,b,c,d (K4s); null if absent
    int32_t t2;           int32_t L;
    int64_t at0;          int64_t step;
    int32_t n_out;        int32_t interp;
    void* out;            int64_t out_stride;
    int32_t n_streams;
    // float64 engines behind a float32 API (ProcessFloat32Into, constant.go:161-199): `in` / `out` hold float32 samples and the
    // kernel converts on load / store instead of two cast launches (streaming-size calls; strides are in float32 elements)
    int32_t in_f32;       int32_t out_f32;
};

// Per polyphase stage, device-resident cache of the K4r kernel's coefficient tiles: for every start phase
// F0 = at0 >> 16 (< L) the phase filters of a period, pre-shifted and grouped the way the kernel's warps read
// them (see fused_up2_rat_kernel). A tile is built by a small kernel the first time its F0 is seen and then
// bulk-copied into shared memory by every block. Owned by the engine; invalidated when a bank is replaced.
struct RatCache {
    void* dev = nullptr;        // [L][tile_bytes]
    size_t tile_bytes = 0;
    int key = 0;                // geometry the tiles were built for (dtype/S/RN/tp/pitch hash)
    std::vector<uint8_t> built; // [L]
};

// Per polyphase stage, device workspace of the persistent chain kernel K5 (kernels_chain.cu): the L2-resident ring of
// intermediate-rate samples and the work-queue / completion counters. Owned by the engine, grown on demand.
struct ChainWs {
    void* dev = nullptr;
    size_t bytes = 0;
};

// cubic.go:33-90 — indices/phases precomputed on the host by the exact float64 recurrence
struct CubicCall {
    const void* hist;     int64_t hist_stride;                         // 3 previous samples (zeros at start)
    const void* in;       int64_t in_stride;    int32_t n_in;
    void* out;            int64_t out_stride;
    void* hist_out;       int64_t hist_out_stride;
    const int32_t* idx;   const double* phase;                         // per output: input index, phase x
    int32_t n_out;        int32_t n_streams;
};

enum Dtype : int { DT_F64 = 0, DT_F32 = 1 };

// Every tuning / A-B environment variable of the kernels (GAR_MMA_*, GAR_K3M_*, GAR_RAT_*, GAR_NO_*, GAR_TENSOR_MIN_ROWS,
// GAR_L2_SLICE_MB ...) is read through this one switch: they are ignored unless GAR_DEBUG_TUNING=1 is set, so a stray
// variable in a production environment cannot change kernel selection. The supported A/B switches are the API calls
// (gar_set_fusion, gar_set_tiled_polyphase, gar_set_tensor_fir, gar_set_slice_budget).
const char* tune_env(const char* name);

// launchers (kernels_fir.cu, kernels_poly.cu, kernels_fused.cu, kernels_misc.cu). Return the name of the kernel variant used.
const char* launch_fir(const FirCall& c, int dtype, cudaStream_t s);
const char* launch_poly(const PolyCall& c, int dtype, cudaStream_t s, RatCache* cache);
const char* launch_cubic(const CubicCall& c, int dtype, cudaStream_t s);
// returns nullptr when the pair cannot be fused (caller falls back to the two stand-alone launches)
const char* launch_fused_up2_poly(const FusedCall& c, int dtype, cudaStream_t s, RatCache* cache);
// K5: the x2 stage and the polyphase stage of a large lock-step batch as ONE persistent launch (intermediate samples in an
// L2-resident ring); false when the call is outside its domain (the caller runs the two stand-alone launches)
// dry: eligibility only (geometry; no allocation, no launch)
bool launch_chain_up2_poly(const FusedCall& c, cudaStream_t s, ChainWs* ws, bool dry = false);
// process-wide policy for K5: 0 never, 1 every eligible call, 2 (default) eligible calls whose full-size intermediate buffer
// would exceed the engine's inter-stage memory budget
void set_chain_kernel(int mode);
int chain_kernel_mode();
int chain_debug_tile_hi(int S, int kp, int NC, int n_tiles_p, int c, int hp, int64_t L, int64_t at0, int64_t step, int n_out);
// float32 I/O folded into the two tensor-core launches of a batched x2 -> polyphase chain: will launch_fir / launch_poly take the
// call with in_f32 / out_f32 set? (the engine asks before it decides against the cast launches)
bool fir_mma_io32_takes(const FirCall& c);
bool up2_poly_runs_as_tensor_pair(const FusedCall& c);  // launch_fused_up2_poly leaves the pair to K1m + K3m / K3p (geometry only)
bool poly_rows_pipe_out32_takes(const PolyCall& c);
// carry only (a call that produced no output but appended to the tail)
void launch_carry(const void* hist, int64_t hist_stride, int32_t hist_len, const void* in, int64_t in_stride,
                  int32_t n_in, void* hist_out, int64_t hist_out_stride, int32_t drop, int32_t new_len,
                  int32_t n_streams, int dtype, cudaStream_t s);
// row-wise dtype casts between I/O and compute precision (constant.go:171-178,195-197)
void launch_cast(const void* src, int64_t src_stride, int src_dtype, void* dst, int64_t dst_stride, int dst_dtype,
                 int32_t n, int32_t n_rows, cudaStream_t s);
// N1 boundary kernels (cmd/resample-wav/main.go:358-520, convenience.go:261-282,463-486):
//   deinterleave: planar[ch][i] = T(double(interleaved[i*C+ch]) * inv_max)   (inv_max == 0: plain cast, float I/O)
//   interleave:   interleaved[i*C+ch] = int(clamp(double(planar[ch][i]), -1, 1) * max_val)  (truncating; max_val == 0: cast)
// fmt: 0 f64, 1 f32, 2 int16, 3 int32, 4 int64 containers
// round_f32: values pass through float32 on the way (float32 engines that compute in float64, see gar_handle::wide_f32)
void launch_deinterleave(const void* in, int fmt, int channels, int64_t n_frames, void* planar, int64_t stride,
                         int dtype, double inv_max, int round_f32, cudaStream_t s);
void launch_interleave(const void* planar, int64_t stride, int dtype, int channels, int64_t n_frames, void* out, int fmt,
                       double max_val, int round_f32, cudaStream_t s);
// dependent-FMA probe; returns elapsed ms for `iters` iterations, flops in *flops
float run_fma_probe(int dtype, int iters, double* flops, cudaStream_t s);
// kernels launched by this library in this process (optionally resetting the counter)
long long launch_count(bool reset);
// process-wide A/B switch for the register-tiled polyphase kernels (K4r / K3r / K3i); on by default
void set_tiled_polyphase(bool on);
// process-wide A/B switch for the FP64 tensor-core FIR kernels (DMMA); on by default
void set_tensor_fir(bool on);
// name the tiled FIR variant that launch_fir would pick (no launch)
const char* fir_variant_name(int dtype, int stride, int nf, int taps, int64_t n_pos, int n_streams);

}  // namespace gar
