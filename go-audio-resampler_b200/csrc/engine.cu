// engine.cu — host runtime (no device code here; built by nvcc for the CUDA runtime API).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace gar {

namespace {
inline bool cuda_ok(cudaError_t e, std::string& err, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
}  // namespace

Engine::~Engine() {
    if (device_ >= 0 && (stream_ || !dev_.empty())) cudaSetDevice(device_);
    for (auto& d : dev_) {
        for (auto& b : d.bank) if (b) cudaFree(b);
        if (d.bank_il) cudaFree(d.bank_il);
        for (auto& h : d.hist) if (h) cudaFree(h);
        if (d.rat_cache.dev) cudaFree(d.rat_cache.dev);
        if (d.chain_ws.dev) cudaFree(d.chain_ws.dev);
    }
    for (void* b : ibuf_) if (b) cudaFree(b);
    if (zeros_) cudaFree(zeros_);
    for (void* p : scratch_) if (p) cudaFree(p);
    if (d_cubic_idx_) cudaFree(d_cubic_idx_);
    if (d_cubic_phase_) cudaFree(d_cubic_phase_);
    if (order_ev_) cudaEventDestroy(order_ev_);
    if (stream_) cudaStreamDestroy(stream_);
}

// All engine state on the device (tails, inter-stage buffers, scratch, coefficient-tile caches) is shared by every call on the
// handle, whatever stream the caller enqueues on. Calls are ordered against each other with one event: recorded on the
// stream of the last enqueue, waited for by the next enqueue when it uses a different stream.
void Engine::order_before(cudaStream_t s) {
    if (order_pending_ && last_stream_ != s) cudaStreamWaitEvent(s, order_ev_, 0);
}
void Engine::order_after(cudaStream_t s) {
    if (!order_ev_) return;
    cudaEventRecord(order_ev_, s);
    last_stream_ = s;
    order_pending_ = true;
}

int Engine::upload_bank(int stage, int which, const std::vector<double>& v, std::string& err) {
    StageDev& d = dev_[(size_t)stage];
    d.rat_cache.built.assign(d.rat_cache.built.size(), 0);  // tiles derive from the bank
    if (d.bank[which]) {
        cudaFree(d.bank[which]);
        d.bank[which] = nullptr;
    }
    if (v.empty()) return 0;
    const size_t bytes = v.size() * esz_;
    if (!cuda_ok(cudaMalloc(&d.bank[which], bytes), err, "cudaMalloc(bank)")) return 4;
    device_bytes_ += (int64_t)bytes;
    if (dtype_ == DT_F32) {  // coefficients are designed in f64 and cast (polyphase_stage.go:149-152, dft_stage.go:98,464)
        std::vector<float> f(v.size());
        for (size_t i = 0; i < v.size(); ++i) f[i] = (float)v[i];
        if (!cuda_ok(cudaMemcpy(d.bank[which], f.data(), bytes, cudaMemcpyHostToDevice), err, "bank upload")) return 4;
        if (which == 0 && chain_.stages[(size_t)stage].kind != STAGE_POLY) d.bank_h32 = std::move(f);
    } else {
        if (!cuda_ok(cudaMemcpy(d.bank[which], v.data(), bytes, cudaMemcpyHostToDevice), err, "bank upload")) return 4;
        if (which == 0 && chain_.stages[(size_t)stage].kind != STAGE_POLY) d.bank_h64 = v;
    }
    return 0;
}

// a,b,c,d of a polyphase stage interleaved per tap (float64 engines, fractional step only): the phase-sorted kernel copies a
// phase's four rows into shared memory with 16-byte asynchronous copies
int Engine::upload_interleaved(int stage, std::string& err) {
    StageDev& d = dev_[(size_t)stage];
    const StageDesign& sd = chain_.stages[(size_t)stage];
    if (d.bank_il) {
        cudaFree(d.bank_il);
        d.bank_il = nullptr;
    }
    if (sd.kind != STAGE_POLY || dtype_ != DT_F64 || !sd.interp) return 0;
    const size_t n = sd.bank[0].size();
    for (int w = 1; w < 4; ++w)
        if (sd.bank[w].size() != n) return 0;
    std::vector<double> il(n * 4);
    for (size_t i = 0; i < n; ++i)
        for (int w = 0; w < 4; ++w) il[i * 4 + (size_t)w] = sd.bank[w][i];
    if (!cuda_ok(cudaMalloc(&d.bank_il, il.size() * 8), err, "cudaMalloc(interleaved bank)")) {
        d.bank_il = nullptr;
        return 4;
    }
    device_bytes_ += (int64_t)(il.size() * 8);
    if (!cuda_ok(cudaMemcpy(d.bank_il, il.data(), il.size() * 8, cudaMemcpyHostToDevice), err, "bank upload")) return 4;
    return 0;
}

int Engine::init(const Chain& chain, int rows, int compute_dtype, int device, std::string& err, bool round_banks_f32) {
    chain_ = chain;
    round_banks_f32_ = round_banks_f32 && compute_dtype == DT_F64;
    if (round_banks_f32_)  // the coefficients a float32 engine uses (polyphase_stage.go:149-152, dft_stage.go:98,464), widened
        for (StageDesign& sd : chain_.stages)
            for (auto& b : sd.bank)
                for (double& v : b) v = (double)(float)v;
    rows_ = rows;
    dtype_ = compute_dtype;
    esz_ = compute_dtype == DT_F32 ? 4 : 8;
    device_ = device;
    if (device == -1) {  // geometry-only handle: integer state machine and banks on the host, nothing on a device
        dev_.assign(chain_.stages.size(), StageDev{});
        name_kernels();
        streams_.assign((size_t)rows, StreamState{});
        reset_state();
        return 0;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        err = "no CUDA device available (this engine has no CPU fallback)";
        device_ = -1;
        return 4;
    }
    if (device < 0 || device >= ndev) {
        err = "CUDA device ordinal out of range";
        device_ = -1;
        return 1;
    }
    if (!cuda_ok(cudaSetDevice(device), err, "cudaSetDevice")) return 4;
    if (!cuda_ok(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), err, "cudaStreamCreate")) return 4;
    if (!cuda_ok(cudaEventCreateWithFlags(&order_ev_, cudaEventDisableTiming), err, "cudaEventCreate")) return 4;

    const size_t S = chain_.stages.size();
    dev_.assign(S, StageDev{});
    for (size_t s = 0; s < S; ++s) {
        StageDesign& sd = chain_.stages[s];
        for (int w = 0; w < 4; ++w) {
            if (sd.kind == STAGE_POLY || w == 0) {
                int rc = upload_bank((int)s, w, sd.bank[w], err);
                if (rc) return rc;
            }
        }
        {
            int rc = upload_interleaved((int)s, err);
            if (rc) return rc;
        }
        int64_t cap = 8;
        switch (sd.kind) {
            case STAGE_UP: cap = sd.taps + 8; break;
            case STAGE_DECIM: cap = sd.taps + sd.factor + 8; break;
            case STAGE_POLY: cap = sd.taps + 600; break;
            case STAGE_CUBIC: cap = 4; break;
        }
        dev_[s].hist_cap = 0;
        int rc = ensure_hist((int)s, cap, stream_, err);
        if (rc) return rc;
    }
    name_kernels();
    ibuf_.assign(chain_.engines.size() * 2, nullptr);
    ibuf_cap_.assign(chain_.engines.size() * 2, 0);
    streams_.assign((size_t)rows, StreamState{});
    if (const char* e = gar::tune_env("GAR_NO_FUSE")) fuse_ = !(e[0] && e[0] != '0');
    if (const char* e = gar::tune_env("GAR_L2_SLICE_MB")) slice_budget_ = (int64_t)std::atoll(e) << 20;
    order_after(stream_);  // the tail clears above were enqueued on stream_: later calls on other streams wait for them
    reset_state();
    return 0;
}

void Engine::note_kernel(const char* name) {
    if (!name) return;
    for (const char* k : kernels_used_)
        if (k == name || std::strcmp(k, name) == 0) return;
    kernels_used_.push_back(name);
}

void Engine::name_kernels() {
    for (size_t s = 0; s < chain_.stages.size(); ++s) {
        const StageDesign& sd = chain_.stages[s];
        switch (sd.kind) {
            case STAGE_UP: dev_[s].kernel = fir_variant_name(dtype_, 1, sd.factor, sd.taps, 0, rows_); break;
            case STAGE_DECIM: dev_[s].kernel = fir_variant_name(dtype_, sd.factor, 1, sd.taps, 0, rows_); break;
            case STAGE_POLY: dev_[s].kernel = dtype_ == DT_F32 ? (sd.interp ? "poly_f32_interp" : "poly_f32")
                                                               : (sd.interp ? "poly_f64_interp" : "poly_f64"); break;
            default: dev_[s].kernel = dtype_ == DT_F32 ? "cubic_f32" : "cubic_f64"; break;
        }
    }
}

void Engine::reset_state() {
    const size_t S = chain_.stages.size(), E = chain_.engines.size();
    for (auto& st : streams_) {
        st.st.assign(S, StageState{});
        st.samples_in.assign(E, 0);
        st.samples_out.assign(E, 0);
        for (size_t s = 0; s < S; ++s)
            if (chain_.stages[s].kind == STAGE_CUBIC) st.st[s].hist_len = 3;  // cubic.go:18: zero-initialised 4-point window
    }
    if (device_ < 0) return;
    // Only the cubic stage keeps device state that Reset must clear (its zero-initialised 4-point window); everything else
    // is defined by the integer state above. The clear is enqueued on the handle's stream and ordered against the calls
    // before and after it by the engine's event — Reset never blocks the host.
    bool any = false;
    for (size_t s = 0; s < S; ++s)
        if (chain_.stages[s].kind == STAGE_CUBIC)
            for (auto& h : dev_[s].hist)
                if (h) {
                    if (!any) {
                        cudaSetDevice(device_);
                        order_before(stream_);
                        any = true;
                    }
                    cudaMemsetAsync(h, 0, (size_t)rows_ * (size_t)dev_[s].hist_cap * esz_, stream_);
                }
    if (any) order_after(stream_);
}

int Engine::set_bank(int stage, int which, const double* coef, int64_t n, std::string& err) {
    if (stage < 0 || stage >= (int)chain_.stages.size() || which < 0 || which > 3) {
        err = "stage/bank index out of range";
        return 1;
    }
    StageDesign& sd = chain_.stages[(size_t)stage];
    if ((int64_t)sd.bank[which].size() != n) {
        err = "bank size mismatch";
        return 1;
    }
    sd.bank[which].assign(coef, coef + n);
    if (device_ < 0) return 0;
    cudaSetDevice(device_);
    cudaDeviceSynchronize();  // the old bank and its coefficient tiles may be in use on any stream
    order_pending_ = false;
    const int rc = upload_bank(stage, which, sd.bank[which], err);
    return rc ? rc : upload_interleaved(stage, err);
}

// Grows both ping-pong tail buffers of a stage. Everything (clear, copy of the old tails) is enqueued on `s`, the stream
// the following kernels run on: the legacy default stream has no ordering with the engine's non-blocking streams.
int Engine::ensure_hist(int stage, int64_t need, cudaStream_t s, std::string& err) {
    StageDev& d = dev_[(size_t)stage];
    if (need <= d.hist_cap) return 0;
    const int64_t ncap = std::max<int64_t>(need + need / 2, 16);
    const size_t bytes = (size_t)rows_ * (size_t)ncap * esz_;
    void* nb[2] = {nullptr, nullptr};
    for (int p = 0; p < 2; ++p)
        if (!cuda_ok(cudaMalloc(&nb[p], bytes), err, "cudaMalloc(hist)")) {
            if (nb[0]) cudaFree(nb[0]);
            return 4;  // nothing published: the old buffers and capacity stay valid
        }
    if (d.hist[0]) cudaDeviceSynchronize();  // the old tails may still be in use on any stream
    for (int p = 0; p < 2; ++p) {
        cudaMemsetAsync(nb[p], 0, bytes, s);
        device_bytes_ += (int64_t)bytes;
        if (d.hist[p]) {
            cudaMemcpy2DAsync(nb[p], (size_t)ncap * esz_, d.hist[p], (size_t)d.hist_cap * esz_, (size_t)d.hist_cap * esz_,
                              (size_t)rows_, cudaMemcpyDeviceToDevice, s);
            cudaStreamSynchronize(s);
            cudaFree(d.hist[p]);
            device_bytes_ -= (int64_t)((size_t)rows_ * (size_t)d.hist_cap * esz_);
        }
        d.hist[p] = nb[p];
    }
    d.hist_cap = ncap;
    return 0;
}

void* Engine::scratch(int slot, size_t bytes, std::string& err) {
    if (bytes <= scratch_cap_[slot] && scratch_[slot]) return scratch_[slot];
    cudaSetDevice(device_);
    if (scratch_[slot]) {
        cudaDeviceSynchronize();
        cudaFree(scratch_[slot]);
        device_bytes_ -= (int64_t)scratch_cap_[slot];
        scratch_[slot] = nullptr;
        scratch_cap_[slot] = 0;
    }
    const size_t nb = std::max<size_t>(bytes + bytes / 4, 1 << 16);
    if (!cuda_ok(cudaMalloc(&scratch_[slot], nb), err, "cudaMalloc(scratch)")) return nullptr;
    scratch_cap_[slot] = nb;
    device_bytes_ += (int64_t)nb;
    return scratch_[slot];
}

int Engine::ensure_internal(const Plan& p, cudaStream_t s, std::string& err) {
    for (size_t b = 0; b < p.buf_need.size(); ++b) {
        if (p.buf_need[b] <= ibuf_cap_[b]) continue;
        if (ibuf_[b]) {
            cudaDeviceSynchronize();
            cudaFree(ibuf_[b]);
            device_bytes_ -= (int64_t)((size_t)rows_ * (size_t)ibuf_cap_[b] * esz_);
            ibuf_[b] = nullptr;
            ibuf_cap_[b] = 0;  // a failed allocation below must not leave a stale capacity behind
        }
        // rows padded to 16 bytes so that row starts stay TMA/vector aligned
        int64_t cap = p.buf_need[b] + p.buf_need[b] / 8 + 64;
        cap = (cap + 3) & ~int64_t(3);
        const size_t bytes = (size_t)rows_ * (size_t)cap * esz_;
        if (!cuda_ok(cudaMalloc(&ibuf_[b], bytes), err, "cudaMalloc(inter-stage buffer)")) return 4;
        ibuf_cap_[b] = cap;
        device_bytes_ += (int64_t)bytes;
    }
    int64_t zneed = 0;
    for (const Op& o : p.ops)
        if (o.src_buf == BUF_ZERO) zneed = std::max(zneed, o.n_in);
    if (zneed > zeros_cap_) {
        if (zeros_) {
            cudaDeviceSynchronize();
            cudaFree(zeros_);
            zeros_ = nullptr;
            zeros_cap_ = 0;
        }
        const int64_t ncap = zneed + 64;
        if (!cuda_ok(cudaMalloc(&zeros_, (size_t)ncap * 8), err, "cudaMalloc(zeros)")) {
            zeros_ = nullptr;
            return 4;
        }
        // cleared on the launch stream: the flush kernels that read the zero row are enqueued on `s` right after
        cudaMemsetAsync(zeros_, 0, (size_t)ncap * 8, s);
        zeros_cap_ = ncap;
    }
    if ((int64_t)p.cubic_idx.size() > cubic_cap_) {
        if (d_cubic_idx_ || d_cubic_phase_) {
            cudaDeviceSynchronize();
            if (d_cubic_idx_) cudaFree(d_cubic_idx_);
            if (d_cubic_phase_) cudaFree(d_cubic_phase_);
            d_cubic_idx_ = nullptr;
            d_cubic_phase_ = nullptr;
            cubic_cap_ = 0;
        }
        const int64_t ncap = (int64_t)p.cubic_idx.size() * 2;
        if (!cuda_ok(cudaMalloc(&d_cubic_idx_, (size_t)ncap * 4), err, "cudaMalloc(cubic idx)")) {
            d_cubic_idx_ = nullptr;
            return 4;
        }
        if (!cuda_ok(cudaMalloc(&d_cubic_phase_, (size_t)ncap * 8), err, "cudaMalloc(cubic phase)")) {
            cudaFree(d_cubic_idx_);
            d_cubic_idx_ = nullptr;
            d_cubic_phase_ = nullptr;
            return 4;
        }
        cubic_cap_ = ncap;  // published only after both allocations succeeded
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Planning: the reference's streaming state machines, integers only.
// ------------------------------------------------------------------------------------------------
void Engine::plan(StreamState& st, int64_t n_in, bool flush, Plan& P) const {
    const size_t E = chain_.engines.size();
    P.ops.clear();
    P.n_out = 0;
    P.buf_need.assign(E * 2, 0);
    P.cubic_idx.clear();
    P.cubic_phase.clear();

    auto note_dst = [&](int buf, int64_t end) {
        if (buf >= 0) P.buf_need[(size_t)buf] = std::max(P.buf_need[(size_t)buf], end);
    };

    // One primitive stage call: Process(n samples from src) appended at dst_off. Returns samples produced.
    auto stage_op = [&](int s, int src_buf, int64_t n, int dst_buf, int64_t dst_off) -> int64_t {
        if (n <= 0) return 0;  // every stage returns early on empty input without touching state
        const StageDesign& sd = chain_.stages[(size_t)s];
        StageState& ss = st.st[(size_t)s];
        Op op;
        op.stage = s;
        op.src_buf = src_buf;
        op.n_in = n;
        op.dst_buf = dst_buf;
        op.dst_off = dst_off;
        op.hist_len = ss.hist_len;
        op.parity_in = ss.parity;
        const int64_t total = ss.hist_len + n;
        op.drop = 0;
        op.new_hist_len = total;
        op.n_out = 0;
        switch (sd.kind) {
            case STAGE_UP: {  // dft_stage.go:169-204
                if (total >= sd.taps) {
                    const int64_t np = total - sd.taps + 1;
                    op.n_out = np * sd.factor;
                    op.drop = np;
                    op.new_hist_len = total - np;
                }
                break;
            }
            case STAGE_DECIM: {  // dft_stage.go:499-551
                if (total >= sd.taps) {
                    const int64_t nf = total - sd.taps + 1;
                    const int64_t M = sd.factor;
                    const int64_t cnt = nf > ss.decim_phase ? (nf - ss.decim_phase + M - 1) / M : 0;
                    if (cnt > 0) {  // cnt == 0: early return, nothing consumed (dft_stage.go:516-518)
                        op.n_out = cnt;
                        op.first = ss.decim_phase;
                        ss.decim_phase = (((ss.decim_phase - nf) % M) + M) % M;
                        op.drop = nf;
                        op.new_hist_len = total - nf;
                    }
                }
                break;
            }
            case STAGE_POLY: {  // polyphase_stage.go:186-312
                const int64_t num_in = total - sd.taps + 1;
                if (num_in > 0) {
                    const int64_t L = sd.factor;
                    const int64_t limit = (num_in * L) << 16;
                    const int64_t num_out = (limit - ss.at + sd.step - 1) / sd.step;
                    if (num_out > 0) {
                        op.n_out = num_out;
                        op.first = ss.at;
                        op.interp = ((sd.step | ss.at) & 0xFFFF) != 0;
                        const int64_t at_end = ss.at + num_out * sd.step;
                        const int64_t consumed = (at_end >> 16) / L;
                        if (consumed > 0 && consumed <= total) {  // :300-304
                            op.drop = consumed;
                            op.new_hist_len = total - consumed;
                        }
                        ss.at = at_end - ((consumed * L) << 16);  // unconditional (:307)
                    }
                }
                break;
            }
            case STAGE_CUBIC: {  // cubic.go:33-63 — exact float64 recurrence, host side
                op.table_off = (int64_t)P.cubic_idx.size();
                double ph = ss.cubic_phase;
                const double inc = 1.0 / sd.ratio;
                // one sequential float64 addition per output (this loop IS the cost of the preset): raw stores into tables sized
                // once — ceil(n * ratio) + 2 bounds the count — instead of two push_backs per output
                const size_t base = P.cubic_idx.size();
                const size_t cap = base + (size_t)std::ceil((double)n * sd.ratio) + 8;
                P.cubic_idx.resize(cap);
                P.cubic_phase.resize(cap);
                int32_t* __restrict__ pi = P.cubic_idx.data();
                double* __restrict__ pp = P.cubic_phase.data();
                size_t w = base;
                for (int64_t i = 0; i < n; ++i) {
                    while (ph < 1.0) {
                        if (w == P.cubic_idx.size()) {  // (cannot happen by the bound; keeps the loop safe against rounding)
                            P.cubic_idx.resize(w + w / 4 + 1024);
                            P.cubic_phase.resize(P.cubic_idx.size());
                            pi = P.cubic_idx.data();
                            pp = P.cubic_phase.data();
                        }
                        pi[w] = (int32_t)i;
                        pp[w] = ph;
                        ++w;
                        ph += inc;
                    }
                    ph -= 1.0;
                }
                P.cubic_idx.resize(w);
                P.cubic_phase.resize(w);
                ss.cubic_phase = ph;
                op.n_out = (int64_t)w - op.table_off;
                op.hist_len = 3;
                op.drop = n;
                op.new_hist_len = 3;
                break;
            }
        }
        ss.hist_len = op.new_hist_len;
        ss.parity ^= 1;
        note_dst(dst_buf, dst_off + op.n_out);
        P.ops.push_back(op);
        return op.n_out;
    };

    auto copy_op = [&](int src_buf, int64_t n, int dst_buf, int64_t dst_off) {
        Op op;
        op.stage = -1;
        op.src_buf = src_buf;
        op.n_in = n;
        op.dst_buf = dst_buf;
        op.dst_off = dst_off;
        op.n_out = n;
        note_dst(dst_buf, dst_off + n);
        P.ops.push_back(op);
    };

    // engine.Resampler.Process (resampler.go:182-227)
    auto engine_process = [&](size_t e, int src_buf, int64_t n, int dst_buf, int64_t& fill) {
        const EngineDesign& ed = chain_.engines[e];
        st.samples_in[e] += n;
        int64_t produced = 0;
        if (ed.n_stages == 0) {
            copy_op(src_buf, n, dst_buf, fill);
            produced = n;
        } else if (ed.n_stages == 1) {
            produced = stage_op(ed.first_stage, src_buf, n, dst_buf, fill);
        } else {
            const int mid = (int)(2 * e);
            const int64_t t = stage_op(ed.first_stage, src_buf, n, mid, 0);
            produced = stage_op(ed.first_stage + 1, mid, t, dst_buf, fill);
        }
        fill += produced;
        st.samples_out[e] += produced;
    };

    // engine.Resampler.Flush (resampler.go:275-322); each stage Flush = Process(zeros[taps]) if it holds history
    auto engine_flush = [&](size_t e, int dst_buf, int64_t& fill) {
        const EngineDesign& ed = chain_.engines[e];
        if (ed.has_cubic) return;
        int64_t produced = 0;
        if (ed.has_pre) {
            const int pre = ed.first_stage;
            if (st.st[(size_t)pre].hist_len > 0) {
                const int64_t z = chain_.stages[(size_t)pre].taps;
                if (ed.has_poly) {
                    const int mid = (int)(2 * e);
                    const int64_t t = stage_op(pre, BUF_ZERO, z, mid, 0);
                    produced += stage_op(pre + 1, mid, t, dst_buf, fill + produced);
                } else {
                    produced += stage_op(pre, BUF_ZERO, z, dst_buf, fill + produced);
                }
            }
        }
        if (ed.has_decim) {
            const int d = ed.first_stage;
            if (st.st[(size_t)d].hist_len > 0)
                produced += stage_op(d, BUF_ZERO, chain_.stages[(size_t)d].taps, dst_buf, fill + produced);
        }
        if (ed.has_poly) {
            const int p = ed.first_stage + 1;
            if (st.st[(size_t)p].hist_len > 0)
                produced += stage_op(p, BUF_ZERO, chain_.stages[(size_t)p].taps, dst_buf, fill + produced);
        }
        fill += produced;
        st.samples_out[e] += produced;
    };

    if (E == 0) {  // ratio within 0.001 of 1: empty pipeline, buffers[0] is also the final buffer (constant.go:255-294)
        if (!flush && n_in > 0) copy_op(BUF_EXT_IN, n_in, BUF_OUT, 0);
        P.n_out = flush ? 0 : n_in;
        return;
    }
    // constant.go:255-345 (process) and :360-386 (flush), front to back
    int cur_buf = BUF_EXT_IN;
    int64_t cur_n = flush ? 0 : n_in;
    for (size_t e = 0; e < E; ++e) {
        const int dst = (e + 1 == E) ? BUF_OUT : (int)(2 * e + 1);
        int64_t fill = 0;
        if (cur_n > 0) engine_process(e, cur_buf, cur_n, dst, fill);
        if (flush) engine_flush(e, dst, fill);
        cur_buf = dst;
        cur_n = fill;
        if (!flush && cur_n == 0) break;
    }
    P.n_out = cur_buf == BUF_OUT ? cur_n : 0;
}

int64_t Engine::advance(int row, int64_t n_in, bool flush) {
    Plan P;
    plan(streams_[(size_t)row], n_in, flush, P);
    return P.n_out;
}

int Engine::lockstep_run(int row0, int row_end) const {
    int n = 1;
    while (row0 + n < row_end && streams_[(size_t)(row0 + n)].same_as(streams_[(size_t)row0])) ++n;
    return n;
}

int64_t Engine::slice_length(int row0, int count, int64_t n_in) const {
    if (slice_budget_ <= 0 || count <= 0 || n_in < 32768 || chain_.stages.size() < 2) return 0;
    // inter-stage elements per input sample, from a dry run of the state machine on a probe length
    StreamState st = streams_[(size_t)row0];
    Plan P;
    const int64_t probe = std::min<int64_t>(n_in, 1 << 16);
    plan(st, probe, false, P);
    int64_t need = 0;
    for (int64_t b : P.buf_need) need = std::max(need, b);
    if (need <= 0) return 0;
    const double per_sample = (double)need / (double)probe * (double)esz_ * (double)count;  // bytes per input sample
    int64_t len = (int64_t)((double)slice_budget_ / per_sample);
    len = std::max<int64_t>(len & ~int64_t(4095), 16384);
    return len < n_in ? len : 0;
}

bool Engine::io32_foldable(int row0, int64_t n_in, bool flush) const {
    if (device_ < 0 || dtype_ != DT_F64 || !fuse_ || flush || n_in <= 0 || n_in > 32768) return false;
    StreamState st = streams_[(size_t)row0];
    Plan P;
    plan(st, n_in, false, P);
    if (P.ops.size() != 2) return false;
    const Op &a = P.ops[0], &b = P.ops[1];
    if (a.stage < 0 || b.stage != a.stage + 1 || a.src_buf != BUF_EXT_IN || b.dst_buf != BUF_OUT || b.src_buf != a.dst_buf ||
        a.dst_buf < 0 || a.n_out <= 0 || b.n_out <= 0 || b.n_in != a.n_out || b.dst_off != 0)
        return false;
    const StageDesign &su = chain_.stages[(size_t)a.stage], &sp = chain_.stages[(size_t)b.stage];
    return su.kind == STAGE_UP && su.factor == 2 && sp.kind == STAGE_POLY && sp.engine_index == su.engine_index &&
           sp.taps <= 200 && b.hist_len <= 1024;
}

bool Engine::pair32_foldable(int row0, int count, int64_t n_in, int64_t in_stride) const {
    if (device_ < 0 || dtype_ != DT_F64 || !fuse_ || n_in <= 32768 || count < 8 || (in_stride & 3) != 0) return false;
    if (slice_length(row0, rows_, n_in) > 0) return false;  // time-sliced calls keep the cast launches
    StreamState st = streams_[(size_t)row0];
    Plan P;
    plan(st, n_in, false, P);
    // the op that reads the caller's input must be an integer-factor FIR stage that the float64 tensor-core kernels take (K1m x2,
    // K2m /2 /3 /4), the op that writes the caller's output such a stage or a polyphase stage that K3p takes; whatever lies
    // between (the further x2 stages of an 8k -> 192k pipeline) runs on float64 inter-stage buffers as ever
    if (P.ops.empty()) return false;
    const Op &a = P.ops.front(), &b = P.ops.back();
    if (a.stage < 0 || b.stage < 0 || a.src_buf != BUF_EXT_IN || b.dst_buf != BUF_OUT || a.n_out <= 0 || b.n_out <= 0 || b.dst_off != 0)
        return false;
    for (size_t k = 0; k < P.ops.size(); ++k) {
        const Op& o = P.ops[k];
        if (o.stage < 0 || o.n_out <= 0 || (k > 0 && o.src_buf == BUF_EXT_IN) || (k + 1 < P.ops.size() && o.dst_buf == BUF_OUT)) return false;
    }
    auto fir_call = [&](const Op& o, bool in32, bool out32) -> bool {
        const StageDesign& sd = chain_.stages[(size_t)o.stage];
        FirCall fc{};
        fc.taps = sd.taps; fc.n_streams = count; fc.in_stride = in32 ? in_stride : 0; fc.in_f32 = in32; fc.out_f32 = out32;
        if (sd.kind == STAGE_UP) { fc.stride = 1; fc.nf = sd.factor; fc.n_pos = (int32_t)(o.n_out / sd.factor); }
        else if (sd.kind == STAGE_DECIM) { fc.stride = sd.factor; fc.nf = 1; fc.first = (int32_t)o.first; fc.n_pos = (int32_t)o.n_out; }
        else return false;
        return fir_mma_io32_takes(fc);
    };
    // an x2 -> polyphase pair that the ordinary dispatch would run as ONE fused launch (few rows: K4r / K4 / K4s are the faster
    // kernels there) keeps it, casts included
    for (size_t k = 0; k + 1 < P.ops.size(); ++k) {
        const Op &u = P.ops[k], &q = P.ops[k + 1];
        const StageDesign &su = chain_.stages[(size_t)u.stage], &sq = chain_.stages[(size_t)q.stage];
        if (q.stage == u.stage + 1 && su.kind == STAGE_UP && su.factor == 2 && sq.kind == STAGE_POLY && q.src_buf == u.dst_buf &&
            q.n_in == u.n_out && sq.engine_index == su.engine_index) {
            FusedCall f{};
            f.np = (int32_t)(u.n_out / 2); f.n_streams = count; f.at0 = q.first; f.step = sq.step; f.interp = q.interp ? 1 : 0;
            if (!up2_poly_runs_as_tensor_pair(f)) return false;
        }
    }
    const bool single = P.ops.size() == 1;
    if (!fir_call(a, true, single)) return false;
    if (single) return true;
    const StageDesign& sp = chain_.stages[(size_t)b.stage];
    if (sp.kind != STAGE_POLY) return fir_call(b, false, true);
    PolyCall pc{};
    pc.taps = sp.taps; pc.L = sp.factor; pc.at0 = b.first; pc.step = sp.step; pc.n_out = (int32_t)b.n_out;
    pc.interp = b.interp ? 1 : 0; pc.n_streams = count;
    return poly_rows_pipe_out32_takes(pc);
}

int Engine::run(int row0, int count, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out, int64_t out_stride,
                int64_t out_cap, bool flush, cudaStream_t s, int64_t* n_out, std::string& err, int io32) {
    if (io32) return run_once(row0, count, d_in, in_stride, n_in, d_out, out_stride, out_cap, flush, s, n_out, err, io32);
    // the slice length is derived from ALL rows of the handle, not from this call's row range: row groups of one batch call
    // must advance through the same sequence of stage calls to stay in lock step (same tail ping-pong parity)
    const int64_t slice = flush || device_ < 0 ? 0 : slice_length(row0, rows_, n_in);
    if (slice <= 0) return run_once(row0, count, d_in, in_stride, n_in, d_out, out_stride, out_cap, flush, s, n_out, err);
    // The call exceeds the inter-stage memory budget. A call over ALL rows of the handle whose x2 -> polyphase pair the
    // persistent chain kernel (K5) takes needs no full-size intermediate buffer at all: one launch instead of time slices.
    // (Row groups of a batch must make the same number of stage calls to stay in lock step, hence all rows or none.)
    // Time slices of 64 K samples or more cost little (a launch ramp and tail per slice) and the two tensor-core launches are
    // the faster kernels, so the default policy turns to K5 only when the budget forces shorter slices.
    if ((chain_kernel_mode() == 1 || (chain_kernel_mode() == 2 && slice < 65536)) && fuse_ && count == rows_ && dtype_ == DT_F64) {
        const int rc = run_once(row0, count, d_in, in_stride, n_in, d_out, out_stride, out_cap, false, s, n_out, err, false, true);
        if (rc != -1) return rc;  // -1: K5 does not take it, nothing was touched
    }
    {   // ErrBufferTooSmall is decided for the whole call, before any state changes (constant.go:107-109)
        StreamState st = streams_[(size_t)row0];
        Plan P;
        plan(st, n_in, false, P);
        if (n_out) *n_out = P.n_out;
        if (P.n_out > out_cap) {
            err = "output buffer too small";
            return 2;
        }
    }
    int64_t produced = 0;
    for (int64_t off = 0; off < n_in; off += slice) {
        const int64_t len = std::min(slice, n_in - off);
        int64_t k = 0;
        const int rc = run_once(row0, count, (const char*)d_in + (size_t)off * esz_, in_stride, len,
                                (char*)d_out + (size_t)produced * esz_, out_stride, out_cap - produced, false, s, &k, err);
        if (rc) return rc;
        produced += k;
    }
    if (n_out) *n_out = produced;
    return 0;
}

int Engine::run_once(int row0, int count, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out, int64_t out_stride,
                     int64_t out_cap, bool flush, cudaStream_t s, int64_t* n_out, std::string& err, int io32, bool chain_required) {
    if (count <= 0) return 0;
    if (device_ < 0) {
        err = "geometry-only handle (device = -1) cannot process samples; there is no CPU fallback";
        return 4;
    }
    if (n_in > 0x7ffffff0LL) {
        err = "chunk too long (row length must fit in 31 bits)";
        return 1;
    }
    cudaSetDevice(device_);
    StreamState st = streams_[(size_t)row0];
    const StreamState before = st;
    Plan P;
    plan(st, n_in, flush, P);
    if (n_out) *n_out = P.n_out;
    if (P.n_out > out_cap) {
        err = "output buffer too small";
        return 2;
    }
    if (P.n_out > 0x7ffffff0LL) {
        err = "output too long";
        return 1;
    }
    // x2 -> polyphase pairs of one engine that the persistent chain kernel (K5) takes: decided on the geometry alone, before the
    // inter-stage buffers are sized — the pair's intermediate buffer is not needed then
    auto up2_poly_pair = [&](size_t oi) -> bool {
        if (!fuse_ || oi + 1 >= P.ops.size()) return false;
        const Op &op = P.ops[oi], &nx = P.ops[oi + 1];
        if (op.stage < 0 || nx.stage != op.stage + 1) return false;
        const StageDesign& su = chain_.stages[(size_t)op.stage];
        return su.kind == STAGE_UP && su.factor == 2 && op.n_out > 0 && chain_.stages[(size_t)nx.stage].kind == STAGE_POLY &&
               nx.src_buf == op.dst_buf && op.dst_buf >= 0 && nx.n_in == op.n_out &&
               chain_.stages[(size_t)nx.stage].engine_index == su.engine_index;
    };
    std::vector<char> use_chain(P.ops.size(), 0);
    bool any_chain = false;
    if (dtype_ == DT_F64 && !io32 && !flush && (chain_required || chain_kernel_mode() == 1)) {
        for (size_t oi = 0; oi < P.ops.size(); ++oi) {
            if (!up2_poly_pair(oi)) continue;
            const Op &op = P.ops[oi], &nx = P.ops[oi + 1];
            const StageDesign &su = chain_.stages[(size_t)op.stage], &spd = chain_.stages[(size_t)nx.stage];
            FusedCall f{};
            f.hu = (int32_t)op.hist_len; f.n_in = (int32_t)op.n_in; f.new_hu = (int32_t)op.new_hist_len;
            f.t1 = su.taps; f.np = (int32_t)(op.n_out / 2);
            f.hp = (int32_t)nx.hist_len; f.new_hp = (int32_t)nx.new_hist_len;
            f.t2 = spd.taps; f.L = spd.factor; f.at0 = nx.first; f.step = spd.step;
            f.n_out = (int32_t)nx.n_out; f.interp = nx.interp ? 1 : 0; f.n_streams = count;
            if (launch_chain_up2_poly(f, s, nullptr, true)) {
                use_chain[oi] = 1;
                any_chain = true;
                bool shared = false;  // (the buffer between the two stages has no other user inside a Process call)
                for (size_t k = 0; k < P.ops.size(); ++k)
                    if (k != oi && k != oi + 1 && (P.ops[k].dst_buf == op.dst_buf || P.ops[k].src_buf == op.dst_buf)) shared = true;
                if (!shared) P.buf_need[(size_t)op.dst_buf] = 0;
            }
        }
    }
    if (chain_required && !any_chain) return -1;  // the caller falls back to time slices; nothing has been touched
    order_before(s);
    int rc = ensure_internal(P, s, err);
    if (rc) return rc;
    for (const Op& op : P.ops)
        if (op.stage >= 0) {
            rc = ensure_hist(op.stage, op.new_hist_len, s, err);
            if (rc) return rc;
        }
    if (!P.cubic_idx.empty()) {
        cudaMemcpyAsync(d_cubic_idx_, P.cubic_idx.data(), P.cubic_idx.size() * 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_cubic_phase_, P.cubic_phase.data(), P.cubic_phase.size() * 8, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);  // tables live in pageable vectors
    }

    auto src_ptr = [&](const Op& op, int64_t& stride) -> const void* {
        if (op.src_buf == BUF_EXT_IN) {
            stride = in_stride;
            return d_in;
        }
        if (op.src_buf == BUF_ZERO) {
            stride = 0;
            return zeros_;
        }
        stride = ibuf_cap_[(size_t)op.src_buf];
        return (const char*)ibuf_[(size_t)op.src_buf] + ((size_t)row0 * (size_t)stride + (size_t)op.src_off) * esz_;
    };
    auto dst_ptr = [&](const Op& op, int64_t& stride) -> void* {
        if (op.dst_buf == BUF_OUT) {
            stride = out_stride;
            return (char*)d_out + (size_t)op.dst_off * esz_;
        }
        stride = ibuf_cap_[(size_t)op.dst_buf];
        return (char*)ibuf_[(size_t)op.dst_buf] + ((size_t)row0 * (size_t)stride + (size_t)op.dst_off) * esz_;
    };

    for (size_t oi = 0; oi < P.ops.size(); ++oi) {
        const Op& op = P.ops[oi];
        int64_t sstride = 0, dstride = 0;
        const void* sp = src_ptr(op, sstride);
        void* dp = dst_ptr(op, dstride);
        // K4: an x2 stage whose whole output is consumed by the polyphase stage of the same engine runs as
        // ONE fused launch; the intermediate-rate samples never reach HBM.
        if (fuse_ && io32 != 2 && op.stage >= 0 && oi + 1 < P.ops.size()) {
            const Op& nx = P.ops[oi + 1];
            const StageDesign& su = chain_.stages[(size_t)op.stage];
            if (nx.stage == op.stage + 1 && su.kind == STAGE_UP && su.factor == 2 && op.n_out > 0 &&
                chain_.stages[(size_t)nx.stage].kind == STAGE_POLY && nx.src_buf == op.dst_buf && op.dst_buf >= 0 &&
                nx.n_in == op.n_out && chain_.stages[(size_t)nx.stage].engine_index == su.engine_index) {
                const StageDesign& spd = chain_.stages[(size_t)nx.stage];
                const StageDev& du = dev_[(size_t)op.stage];
                StageDev& dpv = dev_[(size_t)nx.stage];
                int64_t ostride = 0;
                void* optr = dst_ptr(nx, ostride);
                FusedCall f{};
                f.hist_u = (const char*)du.hist[op.parity_in] + (size_t)row0 * (size_t)du.hist_cap * esz_;
                f.hist_u_stride = du.hist_cap; f.hu = (int32_t)op.hist_len;
                f.in = sp; f.in_stride = sstride; f.n_in = (int32_t)op.n_in;
                f.hist_u_out = (char*)du.hist[op.parity_in ^ 1] + (size_t)row0 * (size_t)du.hist_cap * esz_;
                f.hist_u_out_stride = du.hist_cap; f.drop_u = (int32_t)op.drop; f.new_hu = (int32_t)op.new_hist_len;
                f.bank_u = du.bank[0]; f.t1 = su.taps; f.np = (int32_t)(op.n_out / 2);
                f.bank_u_host = du.bank_h64.empty() ? nullptr : du.bank_h64.data();
                f.hist_p = (const char*)dpv.hist[nx.parity_in] + (size_t)row0 * (size_t)dpv.hist_cap * esz_;
                f.hist_p_stride = dpv.hist_cap; f.hp = (int32_t)nx.hist_len;
                f.hist_p_out = (char*)dpv.hist[nx.parity_in ^ 1] + (size_t)row0 * (size_t)dpv.hist_cap * esz_;
                f.hist_p_out_stride = dpv.hist_cap; f.drop_p = (int32_t)nx.drop; f.new_hp = (int32_t)nx.new_hist_len;
                f.bank_a = dpv.bank[0]; f.bank_b = dpv.bank[1]; f.bank_c = dpv.bank[2]; f.bank_d = dpv.bank[3];
                f.bank_il = dpv.bank_il;
                f.t2 = spd.taps; f.L = spd.factor; f.at0 = nx.first; f.step = spd.step;
                f.n_out = (int32_t)nx.n_out; f.interp = nx.interp ? 1 : 0;
                f.out = optr; f.out_stride = ostride; f.n_streams = count;
                f.in_f32 = f.out_f32 = io32 == 1 ? 1 : 0;
                if (use_chain[oi]) {
                    const size_t ws_before = dpv.chain_ws.bytes;
                    const bool ok = launch_chain_up2_poly(f, s, &dpv.chain_ws);
                    device_bytes_ += (int64_t)dpv.chain_ws.bytes - (int64_t)ws_before;
                    if (!ok) {
                        err = "chain kernel: workspace allocation failed";
                        return 4;
                    }
                    note_kernel("chain_up2_poly_f64_mma");
                    ++launches_;
                    ++oi;
                    continue;
                }
                if (const char* kn = launch_fused_up2_poly(f, dtype_, s, &dpv.rat_cache)) {
                    note_kernel(kn);
                    ++launches_;
                    ++oi;  // the polyphase op is done too
                    continue;
                }
                if (io32) {
                    err = "internal: float32 I/O was folded into a call the fused kernel did not take";
                    return 5;
                }
            }
        }
        if (io32 == 1) {
            err = "internal: float32 I/O folding needs the fused x2 -> polyphase launch";
            return 5;
        }
        if (op.stage < 0) {
            launch_cast(sp, sstride, dtype_, dp, dstride, dtype_, (int32_t)op.n_in, count, s);
            ++launches_;
            continue;
        }
        const StageDesign& sd = chain_.stages[(size_t)op.stage];
        StageDev& dv = dev_[(size_t)op.stage];
        const void* hin = (const char*)dv.hist[op.parity_in] + (size_t)row0 * (size_t)dv.hist_cap * esz_;
        void* hout = (char*)dv.hist[op.parity_in ^ 1] + (size_t)row0 * (size_t)dv.hist_cap * esz_;
        switch (sd.kind) {
            case STAGE_UP:
            case STAGE_DECIM: {
                FirCall c{};
                c.hist = hin; c.hist_stride = dv.hist_cap; c.hist_len = (int32_t)op.hist_len;
                c.in = sp; c.in_stride = sstride; c.n_in = (int32_t)op.n_in;
                c.out = dp; c.out_stride = dstride;
                c.hist_out = hout; c.hist_out_stride = dv.hist_cap;
                c.drop = (int32_t)op.drop; c.new_hist_len = (int32_t)op.new_hist_len;
                c.bank = dv.bank[0];
                c.bank_host_f32 = dv.bank_h32.empty() ? nullptr : dv.bank_h32.data();
                c.bank_host_f64 = dv.bank_h64.empty() ? nullptr : dv.bank_h64.data();
                c.taps = sd.taps;
                if (sd.kind == STAGE_UP) {
                    c.stride = 1; c.nf = sd.factor; c.first = 0; c.n_pos = (int32_t)(op.n_out / sd.factor);
                } else {
                    c.stride = sd.factor; c.nf = 1; c.first = (int32_t)op.first; c.n_pos = (int32_t)op.n_out;
                }
                c.n_streams = count;
                c.in_f32 = io32 == 2 && op.src_buf == BUF_EXT_IN ? 1 : 0;
                c.out_f32 = io32 == 2 && op.dst_buf == BUF_OUT ? 1 : 0;
                const char* kn = launch_fir(c, dtype_, s);
                if (!kn) {
                    err = "internal: float32 I/O was folded into a call the tensor-core FIR kernels did not take";
                    return 5;
                }
                note_kernel(kn);
                break;
            }
            case STAGE_POLY: {
                PolyCall c{};
                c.hist = hin; c.hist_stride = dv.hist_cap; c.hist_len = (int32_t)op.hist_len;
                c.in = sp; c.in_stride = sstride; c.n_in = (int32_t)op.n_in;
                c.out = dp; c.out_stride = dstride;
                c.hist_out = hout; c.hist_out_stride = dv.hist_cap;
                c.drop = (int32_t)op.drop; c.new_hist_len = (int32_t)op.new_hist_len;
                c.bank_a = dv.bank[0]; c.bank_b = dv.bank[1]; c.bank_c = dv.bank[2]; c.bank_d = dv.bank[3];
                c.bank_il = dv.bank_il;
                c.taps = sd.taps; c.L = sd.factor; c.at0 = op.first; c.step = sd.step;
                c.n_out = (int32_t)op.n_out; c.interp = op.interp ? 1 : 0;
                c.n_streams = count;
                c.out_f32 = io32 == 2 && op.dst_buf == BUF_OUT ? 1 : 0;
                const char* kn = launch_poly(c, dtype_, s, &dv.rat_cache);
                if (!kn) {
                    err = "internal: float32 output was folded into a call the pipelined polyphase kernel did not take";
                    return 5;
                }
                note_kernel(kn);
                break;
            }
            case STAGE_CUBIC: {
                CubicCall c{};
                c.hist = hin; c.hist_stride = dv.hist_cap;
                c.in = sp; c.in_stride = sstride; c.n_in = (int32_t)op.n_in;
                c.out = dp; c.out_stride = dstride;
                c.hist_out = hout; c.hist_out_stride = dv.hist_cap;
                c.idx = d_cubic_idx_ + op.table_off; c.phase = d_cubic_phase_ + op.table_off;
                c.n_out = (int32_t)op.n_out; c.n_streams = count;
                note_kernel(launch_cubic(c, dtype_, s));
                break;
            }
        }
        ++launches_;
    }
    order_after(s);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        err = std::string("kernel launch failed: ") + cudaGetErrorString(le);
        return 4;
    }
    // commit: identical stage state for the whole run, per-row statistics advance by the same delta
    const size_t E = chain_.engines.size();
    for (int r = 0; r < count; ++r) {
        StreamState& dst = streams_[(size_t)(row0 + r)];
        dst.st = st.st;
        for (size_t e = 0; e < E; ++e) {
            dst.samples_in[e] += st.samples_in[e] - before.samples_in[e];
            dst.samples_out[e] += st.samples_out[e] - before.samples_out[e];
        }
    }
    return 0;
}

}  // namespace gar
