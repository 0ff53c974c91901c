// engine.hpp — host runtime of the B200 resampling engine.
//
// Split of responsibilities:
//   * integers on the host: the streaming state machine of every stage (history
//     length, fixed-point phase accumulator, decimation phase) is advanced on the
//     CPU with the reference's exact integer logic, because sample counts, phase
//     positions and flush lengths must be bit-exact and never depend on sample
//     values;
//   * samples on the device: carried tails, inter-stage buffers and all arithmetic.
// A call is first *planned* (a short list of primitive stage launches with their
// geometry), then executed asynchronously on a CUDA stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "design.hpp"
#include "kernels.cuh"

namespace gar {

struct StageState {      // per stream, per primitive stage
    int64_t hist_len = 0;
    int64_t at = 0;           // POLY (polyphase_stage.go:39-41)
    int64_t decim_phase = 0;  // DECIM (dft_stage.go:381)
    double cubic_phase = 0;   // CUBIC (cubic.go:17)
    int parity = 0;           // which ping-pong tail buffer is current
    bool operator==(const StageState& o) const {
        return hist_len == o.hist_len && at == o.at && decim_phase == o.decim_phase && cubic_phase == o.cubic_phase &&
               parity == o.parity;
    }
};

struct StreamState {
    std::vector<StageState> st;
    std::vector<int64_t> samples_in, samples_out;  // per engine (resampler.go:348-353)
    bool same_as(const StreamState& o) const { return st == o.st; }
};

// buffer ids used by ops
enum : int { BUF_EXT_IN = -1, BUF_ZERO = -2, BUF_OUT = -3 };

struct Op {
    int stage = -1;          // -1: plain copy src -> dst
    int src_buf = BUF_EXT_IN;
    int64_t src_off = 0, n_in = 0;
    int dst_buf = BUF_OUT;
    int64_t dst_off = 0, n_out = 0;
    // geometry of the stage call
    int64_t hist_len = 0, first = 0, drop = 0, new_hist_len = 0;
    int parity_in = 0;
    bool interp = false;
    // cubic: tables (host) for this op
    int64_t table_off = 0;
};

struct Plan {
    std::vector<Op> ops;
    int64_t n_out = 0;                  // samples delivered to BUF_OUT
    std::vector<int64_t> buf_need;      // elements needed per internal buffer
    std::vector<int32_t> cubic_idx;     // cubic tables, concatenated
    std::vector<double> cubic_phase;
};

struct StageDev {
    void* bank[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> bank_h64;  // float64 engines, FIR stages: host copy of bank[0] as uploaded (kernel-parameter taps)
    std::vector<float> bank_h32;  // float32 engines, FIR stages: host copy of bank[0] as uploaded (kernel-parameter taps)
    void* bank_il = nullptr;  // POLY stages with a fractional step: a,b,c,d interleaved per tap [L][taps][4] (K4s)
    void* hist[2] = {nullptr, nullptr};
    int64_t hist_cap = 0;
    const char* kernel = "";
    RatCache rat_cache;  // POLY stages: coefficient tiles of the fused rational-ratio kernel
    ChainWs chain_ws;    // POLY stages: ring + counters of the persistent x2 -> polyphase chain kernel (K5)
};

class Engine {
  public:
    Engine() = default;
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    // returns gar_status
    // round_banks_f32: float64 compute on float32-rounded coefficients (float32 engines in wide mode)
    int init(const Chain& chain, int rows, int compute_dtype, int device, std::string& err, bool round_banks_f32 = false);

    const Chain& chain() const { return chain_; }
    int rows() const { return rows_; }
    int compute_dtype() const { return dtype_; }
    int device() const { return device_; }
    cudaStream_t stream() const { return stream_; }
    const StreamState& state(int row) const { return streams_[(size_t)row]; }
    const StageDev& stage_dev(int s) const { return dev_[(size_t)s]; }
    int64_t launches() const { return launches_; }
    void set_fuse(bool on) { fuse_ = on; }
    void reset_launches() { launches_ = 0; }
    int64_t device_bytes() const { return device_bytes_; }
    // distinct kernel variants this engine has launched so far (names are static strings)
    const std::vector<const char*>& kernels_used() const { return kernels_used_; }

    // Planning (pure integer; mutates `st`). flush=false: Process(n_in). flush=true: Flush().
    // pipeline_mode: path A semantics (constant.go) vs single engine (resampler.go).
    void plan(StreamState& st, int64_t n_in, bool flush, Plan& out) const;

    // Execute on rows [row0,row0+count) which must share identical state. d_in/d_out point at row row0,
    // compute dtype. Enqueues on `s`; commits the state. Returns status; *n_out per row.
    // io32: d_in / d_out hold float32 samples although the engine computes in float64 (strides in float32 elements); only
    // valid when io32_foldable() said so — the single fused launch converts on load / store (no cast launches).
    int run(int row0, int count, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out, int64_t out_stride,
            int64_t out_cap, bool flush, cudaStream_t s, int64_t* n_out, std::string& err, int io32 = 0);
    // A Process call of n_in samples on rows in row0's state is ONE fused x2 -> polyphase launch of streaming size, which can
    // take float32 input / output directly (constant.go:161-199 ProcessFloat32Into without the two cast passes).
    bool io32_foldable(int row0, int64_t n_in, bool flush) const;
    // io32 = 2: a large batched call whose first and last stage calls run on the float64 tensor-core kernels (K1m x2 / K2m
    // decimators / K3p polyphase) can take float32 input / output as well: K1m / K2m widen their sample windows in shared
    // memory, the last kernel narrows on the store
    bool pair32_foldable(int row0, int count, int64_t n_in, int64_t in_stride) const;
    // Time slicing of multi-stage calls: a Process call whose inter-stage buffers would exceed `bytes` is run as a
    // sequence of shorter Process calls (identical samples and counts: every stage is greedy), so that the intermediate-rate
    // buffers stay small (and, with an L2-sized budget, the intermediate-rate streams stay L2-resident). 0 disables.
    void set_slice_budget(int64_t bytes) { slice_budget_ = bytes; }
    int64_t slice_budget() const { return slice_budget_; }
    // samples of one row a time slice may hold for `count` rows (0: no slicing needed for n_in)
    int64_t slice_length(int row0, int count, int64_t n_in) const;

    // Advance only the integer state of one row (geometry-only use; no samples move).
    int64_t advance(int row, int64_t n_in, bool flush);

    // Find the maximal run of rows starting at row0 (within [row0,row_end)) sharing row0's state.
    int lockstep_run(int row0, int row_end) const;

    void reset_state();
    int set_bank(int stage, int which, const double* coef, int64_t n, std::string& err);

    // staging helpers (grow-only device scratch in bytes), slot 0..3
    void* scratch(int slot, size_t bytes, std::string& err);
    // cross-stream ordering for callers that touch the scratch slots outside run() (capi.cu)
    void begin_on(cudaStream_t s) { order_before(s); }
    void end_on(cudaStream_t s) { order_after(s); }

  private:
    void name_kernels();
    int ensure_internal(const Plan& p, cudaStream_t s, std::string& err);
    int ensure_hist(int stage, int64_t need, cudaStream_t s, std::string& err);
    void order_before(cudaStream_t s);
    void order_after(cudaStream_t s);
    int upload_bank(int stage, int which, const std::vector<double>& v, std::string& err);
    int upload_interleaved(int stage, std::string& err);

    Chain chain_;
    int rows_ = 0, dtype_ = DT_F64, device_ = 0;
    size_t esz_ = 8;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t order_ev_ = nullptr;      // recorded after every enqueue (cross-stream ordering of the shared state)
    cudaStream_t last_stream_ = nullptr;
    bool order_pending_ = false;
    std::vector<StreamState> streams_;
    std::vector<StageDev> dev_;
    std::vector<void*> ibuf_;         // internal buffers [rows][cap]
    std::vector<int64_t> ibuf_cap_;   // elements per row
    void* zeros_ = nullptr;
    int64_t zeros_cap_ = 0;
    void* scratch_[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap_[4] = {0, 0, 0, 0};
    int32_t* d_cubic_idx_ = nullptr;
    double* d_cubic_phase_ = nullptr;
    int64_t cubic_cap_ = 0;
    bool round_banks_f32_ = false;
    bool fuse_ = true;  // K4 fused x2 -> polyphase launches (GAR_NO_FUSE=1 disables, for A/B tests)
    int64_t launches_ = 0;
    // bytes of the largest inter-stage buffer per slice. Default 2 GiB: a memory-footprint guard only. L2-sized slices
    // (40 MiB) keep the intermediate stream out of HBM but measured 20 % slower on the batched chains (57 launches
    // instead of 5: every slice pays a launch ramp and tail), and those chains are FMA-bound, not HBM-bound.
    int64_t slice_budget_ = 2ll << 30;
    int run_once(int row0, int count, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out, int64_t out_stride,
                 int64_t out_cap, bool flush, cudaStream_t s, int64_t* n_out, std::string& err, int io32 = 0,
                 bool chain_required = false);  // chain_required: K5 or nothing (-1: not taken, nothing touched)
    std::vector<const char*> kernels_used_;
    void note_kernel(const char* name);
    int64_t device_bytes_ = 0;
};

}  // namespace gar
