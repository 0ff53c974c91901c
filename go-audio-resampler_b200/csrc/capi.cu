// capi.cu — the C ABI declared in include/gar.h on top of gar::Engine.
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gar.h"
#include "engine.hpp"
#include "hostnuma.hpp"

using namespace gar;

// ---- multi-device handles ------------------------------------------------------------------------
// The reference fans the channels of one call out over goroutines (constant.go:223-241). The GPU analogue: one handle
// shards its rows (whole streams, or channels when there is a single stream) over K devices. Every shard is a complete
// single-device handle driven by its own worker thread, which is bound to the cores of the device's NUMA node so that
// staging, launches and the first touch of sharded pinned buffers all happen next to the device. Rows never migrate
// between devices (carry state is per row), and there is no cross-device exchange: the host joins the workers.
struct ShardWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false;
    int node = -1;

    void start(int device) {
        th = std::thread([this, device] {
            if (device >= 0) {
                node = bind_thread_to_device(device);
                cudaSetDevice(device);
            }
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return has_job || quit; });
                if (quit) return;
                lk.unlock();
                job();
                lk.lock();
                has_job = false;
                cv.notify_all();
            }
        });
    }
    void post(std::function<void()> f) {
        std::unique_lock<std::mutex> lk(m);
        job = std::move(f);
        has_job = true;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return !has_job; });
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [this] { return !has_job; });
            quit = true;
            cv.notify_all();
        }
        th.join();
    }
};

struct Shard {
    gar_handle* h = nullptr;  // complete single-device handle for rows [row0, row0 + rows)
    int row0 = 0, rows = 0, device = 0;
    int rc = 0;               // status of the last job
    std::string err;          // message of a failed job that has no handle to carry it (shard creation)
    int64_t n_out = 0;
    ShardWorker w;
};

struct gar_handle {
    gar_config cfg{};
    Engine eng;
    double ratio = 1.0;
    int rows = 1;
    int compute_dtype = DT_F64;
    // float32 ENGINE handle whose batches run on the float64 tensor-core kernels: samples and coefficients are float32 values
    // (I/O, banks rounded to float32), arithmetic and carried state are float64 — at least as accurate as float32 arithmetic
    // and 1.6x (rational) to 7x (irrational ratios) faster than the float32 one-thread-per-output kernels for >= 8-32 rows.
    bool wide_f32 = false;
    std::string err;
    std::string gpu_name;
    // batch pipelining
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    void* slot_in[2] = {nullptr, nullptr};
    void* slot_out[2] = {nullptr, nullptr};
    size_t slot_in_cap = 0, slot_out_cap = 0;
    void* slot_cin = nullptr;   // cast scratch (compute dtype) for I/O dtype != compute dtype
    void* slot_cout = nullptr;
    size_t slot_cin_cap = 0, slot_cout_cap = 0;
    void* pin = nullptr;  // pinned, device-mapped staging of streaming-size per-channel calls (zero-copy over PCIe)
    size_t pin_cap = 0;
    // multi-device handle: `eng` is geometry-only (chain, banks, static info); all streaming state lives in the shards
    std::vector<std::unique_ptr<Shard>> shards;
    bool multi() const { return !shards.empty(); }
};

namespace {
// run f(shard) on every shard's worker thread and join; returns the first non-OK status (error text copied up)
template <typename F>
int for_shards(gar_handle* h, F f) {
    for (auto& sp : h->shards) {
        Shard* sh = sp.get();
        sh->w.post([sh, f] { sh->rc = f(*sh); });
    }
    int rc = 0;
    for (auto& sp : h->shards) {
        sp->w.wait();
        if (sp->rc && !rc) {
            rc = sp->rc;
            h->err = "device " + std::to_string(sp->device) + ": " + (sp->h ? sp->h->err : sp->err);
        }
    }
    return rc;
}
// shard owning `row`; *local = row index inside it
Shard* shard_of(const gar_handle* h, int row, int* local) {
    for (auto& sp : h->shards)
        if (row >= sp->row0 && row < sp->row0 + sp->rows) {
            if (local) *local = row - sp->row0;
            return sp.get();
        }
    return nullptr;
}
// pinned buffers made by gar_host_alloc_rows (mmap + first touch + cudaHostRegister): base -> bytes
std::mutex g_reg_mutex;
std::map<void*, size_t> g_registered;
}  // namespace

static thread_local std::string g_create_err;

static int fail(gar_handle* h, int status, const std::string& msg) {
    if (h) h->err = msg;
    return status;
}

static inline size_t dsize(int dt) { return dt == GAR_F32 ? 4 : 8; }

extern "C" {

const char* gar_version(void) { return "gar-b200 0.1.0 (sm_100a)"; }

const char* gar_status_string(int32_t s) {
    switch (s) {
        case GAR_OK: return "ok";
        case GAR_INVALID_CONFIG: return "invalid resampler configuration";
        case GAR_BUFFER_TOO_SMALL: return "output buffer too small";
        case GAR_NOT_SUPPORTED: return "operation not supported";
        case GAR_CUDA_ERROR: return "CUDA error";
        default: return "internal error";
    }
}

const char* gar_last_error(const gar_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int32_t gar_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int32_t gar_create(const gar_config* cfg, gar_handle** out) {
    if (out) *out = nullptr;
    if (!cfg || !out) {
        g_create_err = "config is nil";
        return GAR_INVALID_CONFIG;
    }
    auto bad = [&](const char* m) {
        g_create_err = m;
        return (int32_t)GAR_INVALID_CONFIG;
    };
    // Config.Validate (resample.go:168-214) / NewResampler (resampler.go:51-70)
    if (!(cfg->input_rate > 0) || !(cfg->output_rate > 0)) return bad("sample rates must be positive");
    const int path = cfg->path;
    if (path != GAR_PATH_PIPELINE && path != GAR_PATH_ENGINE) return bad("unknown path");
    int channels = cfg->channels;
    if (path == GAR_PATH_ENGINE && channels == 0) channels = 1;
    if (channels < 1) return bad("channels must be at least 1");
    if (channels > 256) return bad("too many channels (max 256)");
    const double ratio = cfg->output_rate / cfg->input_rate;
    if (ratio < 1.0 / 256.0 || ratio > 256.0) return bad("resampling ratio out of range");
    if (cfg->preset == GAR_QUALITY_CUSTOM && path == GAR_PATH_PIPELINE) {  // QualitySpec.Validate
        if (cfg->custom_precision < 8 || cfg->custom_precision > 33) return bad("precision must be 8-33 bits");
        if (cfg->custom_phase_response < 0 || cfg->custom_phase_response > 100) return bad("phase response must be 0-100");
        if (!(cfg->custom_passband_end > 0) || !(cfg->custom_passband_end < 1)) return bad("passband end must be in (0, 1)");
        if (!(cfg->custom_stopband_begin > cfg->custom_passband_end) || cfg->custom_stopband_begin > 1)
            return bad("stopband begin must be in (passband_end, 1]");
    }
    if (cfg->dtype != GAR_F64 && cfg->dtype != GAR_F32) return bad("unknown dtype");
    if (cfg->n_streams < 0) return bad("n_streams must be >= 0");

    Chain chain;
    std::string err;
    int compute = DT_F64;
    if (path == GAR_PATH_PIPELINE) {
        // the pipeline always computes in float64 (constant.go:161-199); dtype only selects the I/O type
        const int precision = cfg->preset == GAR_QUALITY_CUSTOM ? cfg->custom_precision : preset_precision(cfg->preset);
        if (!design_pipeline(cfg->input_rate, cfg->output_rate, precision, chain, err)) {
            g_create_err = err;
            return GAR_INVALID_CONFIG;
        }
    } else {
        const int q = cfg->engine_quality >= 0 ? cfg->engine_quality : preset_to_engine_quality(cfg->preset);
        if (q > EQ_32BIT) return bad("unknown engine quality");
        chain.ratio = ratio;
        if (!design_engine(cfg->input_rate, cfg->output_rate, q, chain, err)) {
            g_create_err = err;
            return GAR_INVALID_CONFIG;
        }
        compute = cfg->dtype == GAR_F32 ? DT_F32 : DT_F64;
    }
    chain.ratio = ratio;
    // wide mode: a float32 engine at a non-integer ratio (x2 stage + polyphase stage) computes in float64 where those kernels are
    // the faster ones: rational ratios from 32 rows (K1m + K3p on the FP64 tensor cores; below that the float32 fused kernel is
    // on a par), ratios with a fractional phase step at ANY row count — the float32 thread-per-output kernel with interpolated
    // coefficients runs a 10 s stream in 199 us, the phase-sorted float64 kernel K4s in 72 us (tools/bench_few_rows.py)
    bool wide = false;
    if (path == GAR_PATH_ENGINE && compute == DT_F32) {
        const int rows_total = channels * (cfg->n_streams > 1 ? cfg->n_streams : 1);
        for (const StageDesign& sd : chain.stages)
            if (sd.kind == STAGE_POLY && rows_total >= (sd.interp ? 1 : 32)) wide = true;
        if (const char* e = gar::tune_env("GAR_NO_WIDE_F32")) wide = wide && !(e[0] && e[0] != '0');
        if (wide) compute = DT_F64;
    }

    std::unique_ptr<gar_handle> h(new gar_handle());
    h->cfg = *cfg;
    h->cfg.channels = channels;
    h->ratio = ratio;
    h->rows = channels * (cfg->n_streams > 1 ? cfg->n_streams : 1);
    h->compute_dtype = compute;
    h->wide_f32 = wide;
    int rc = h->eng.init(chain, h->rows, compute, cfg->device, err, wide);
    if (rc) {
        g_create_err = err;
        return rc;
    }
    cudaDeviceProp prop{};
    if (cfg->device >= 0 && cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->gpu_name = prop.name;
    *out = h.release();
    return GAR_OK;
}

int32_t gar_create_multi(const gar_config* cfg, const int32_t* devices, int32_t n_devices, gar_handle** out) {
    if (out) *out = nullptr;
    if (!cfg || !out || !devices || n_devices < 1 || n_devices > 64) {
        g_create_err = "config or device list is nil / empty";
        return GAR_INVALID_CONFIG;
    }
    // all ordinals -1: geometry-only shards (host logic of the sharding testable without a GPU, like device = -1 handles)
    bool geometry_only = true;
    for (int i = 0; i < n_devices; ++i) geometry_only = geometry_only && devices[i] == -1;
    const int ndev = geometry_only ? 0 : gar_device_count();
    if (!geometry_only && ndev <= 0) {
        g_create_err = "no CUDA device available (this engine has no CPU fallback)";
        return GAR_CUDA_ERROR;
    }
    for (int i = 0; i < n_devices && !geometry_only; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) {
            g_create_err = "CUDA device ordinal out of range";
            return GAR_INVALID_CONFIG;
        }
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) {
                g_create_err = "device listed twice";
                return GAR_INVALID_CONFIG;
            }
    }
    // the top handle validates the config, designs the filters once and answers every state-independent query
    gar_config top = *cfg;
    top.device = -1;
    gar_handle* h = nullptr;
    int32_t rc = gar_create(&top, &h);
    if (rc) return rc;
    h->cfg.device = devices[0];
    // units that may be separated: whole streams (each with all its channels), or channels of the single stream
    const bool by_stream = h->cfg.n_streams > 1;
    const int units = by_stream ? h->cfg.n_streams : h->cfg.channels;
    const int unit_rows = by_stream ? h->cfg.channels : 1;
    const int K = std::min<int>(n_devices, units);
    int unit0 = 0;
    for (int d = 0; d < K; ++d) {
        const int cnt = units / K + (d < units % K ? 1 : 0);
        std::unique_ptr<Shard> sh(new Shard());
        sh->device = devices[d];
        sh->row0 = unit0 * unit_rows;
        sh->rows = cnt * unit_rows;
        sh->w.start(sh->device);
        h->shards.push_back(std::move(sh));
        unit0 += cnt;
    }
    const gar_config base = h->cfg;
    rc = for_shards(h, [base, by_stream, unit_rows](Shard& sh) -> int {
        gar_config c = base;
        c.device = sh.device;
        if (by_stream) c.n_streams = sh.rows / unit_rows;
        else c.channels = sh.rows;
        const int32_t r = gar_create(&c, &sh.h);
        if (r) sh.err = g_create_err;  // gar_create's message is thread-local on the worker: keep it for for_shards
        return r;
    });
    if (rc) {
        g_create_err = h->err;
        gar_destroy(h);
        return rc;
    }
    h->gpu_name = h->shards[0]->h->gpu_name + " x" + std::to_string(K);
    *out = h;
    return GAR_OK;
}

int32_t gar_num_devices(const gar_handle* h) { return !h ? 0 : (h->multi() ? (int32_t)h->shards.size() : 1); }

int32_t gar_shard_info(const gar_handle* h, int32_t shard, int32_t* device, int32_t* row0, int32_t* rows, int32_t* numa_node) {
    if (!h) return GAR_INVALID_CONFIG;
    if (!h->multi()) {
        if (shard != 0) return GAR_INVALID_CONFIG;
        if (device) *device = h->eng.device();
        if (row0) *row0 = 0;
        if (rows) *rows = h->rows;
        if (numa_node) *numa_node = h->eng.device() >= 0 ? device_numa_node(h->eng.device()) : -1;
        return GAR_OK;
    }
    if (shard < 0 || shard >= (int)h->shards.size()) return GAR_INVALID_CONFIG;
    const Shard& sh = *h->shards[(size_t)shard];
    if (device) *device = sh.device;
    if (row0) *row0 = sh.row0;
    if (rows) *rows = sh.rows;
    if (numa_node) *numa_node = sh.device >= 0 ? device_numa_node(sh.device) : -1;
    return GAR_OK;
}

void gar_destroy(gar_handle* h) {
    if (!h) return;
    if (h->multi()) {
        for (auto& sp : h->shards) {
            Shard* sh = sp.get();
            sh->w.post([sh] {
                gar_destroy(sh->h);
                sh->h = nullptr;
            });
        }
        for (auto& sp : h->shards) sp->w.stop();
        h->shards.clear();
    }
    if (h->eng.device() < 0) {
        delete h;
        return;
    }
    cudaSetDevice(h->eng.device());
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
        if (h->slot_in[i]) cudaFree(h->slot_in[i]);
        if (h->slot_out[i]) cudaFree(h->slot_out[i]);
    }
    if (h->pin) cudaFreeHost(h->pin);
    if (h->slot_cin) cudaFree(h->slot_cin);
    if (h->slot_cout) cudaFree(h->slot_cout);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
}

// ---- geometry / info -------------------------------------------------------------------------
int64_t gar_estimate_output(const gar_handle* h, int64_t n_in) {
    return (int64_t)((double)n_in * h->ratio) + 64;  // constants.go:58
}

int64_t gar_next_output_count(const gar_handle* h, int32_t stream, int64_t n_in) {
    if (!h || stream < 0 || stream >= h->rows || n_in < 0) return -1;
    if (h->multi()) {
        int local = 0;
        const Shard* sh = shard_of(h, stream, &local);
        return gar_next_output_count(sh->h, local, n_in);
    }
    StreamState st = h->eng.state(stream);
    Plan p;
    h->eng.plan(st, n_in, false, p);
    return p.n_out;
}

int64_t gar_next_flush_count(const gar_handle* h, int32_t stream) {
    if (!h || stream < 0 || stream >= h->rows) return -1;
    if (h->multi()) {
        int local = 0;
        const Shard* sh = shard_of(h, stream, &local);
        return gar_next_flush_count(sh->h, local);
    }
    StreamState st = h->eng.state(stream);
    Plan p;
    h->eng.plan(st, 0, true, p);
    return p.n_out;
}

double gar_get_ratio(const gar_handle* h) { return h->ratio; }

static int engine_latency(const Chain& c, const EngineDesign& e) {  // stage_adapter.go:43-57
    if (e.has_cubic) return 2;                                      // cubic.go:98 (cubicLatencySamples)
    int l = 0;
    if (e.has_pre) {
        const StageDesign& s = c.stages[(size_t)e.first_stage];
        if (s.factor > 1) l += (s.taps * s.factor) / 2;
    }
    if (e.has_poly) l += c.stages[(size_t)e.first_stage + 1].taps / 2;
    return l;
}

int32_t gar_get_latency(const gar_handle* h) {  // constant.go:407-423
    const Chain& c = h->eng.chain();
    int tot = 0;
    for (const EngineDesign& e : c.engines) tot += (int)((double)engine_latency(c, e) * e.ratio);
    return tot;
}

int32_t gar_get_info(const gar_handle* h, gar_info* out) {  // constant.go:452-485
    if (!h || !out) return GAR_INVALID_CONFIG;
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->algorithm, sizeof(out->algorithm), "%s", "multi-stage");
    out->latency = gar_get_latency(h);
    out->memory_usage = h->eng.device_bytes();
    for (const auto& sp : h->shards) out->memory_usage += sp->h->eng.device_bytes();
    const Chain& c = h->eng.chain();
    if (!c.engines.empty()) {
        const EngineDesign& e = c.engines[0];
        if (e.has_cubic) {
            out->filter_length = 4;  // cubic.go:110
        } else {
            if (e.has_pre) {
                const StageDesign& s = c.stages[(size_t)e.first_stage];
                if (s.factor > 1) out->filter_length += s.taps * s.factor;  // stage_adapter.go:98-110
            }
            if (e.has_poly) {
                const StageDesign& s = c.stages[(size_t)e.first_stage + 1];
                out->filter_length += s.taps * s.factor;
                out->phases = s.factor;
            }
            out->simd_enabled = 1;
        }
    }
    std::snprintf(out->simd_type, sizeof(out->simd_type), "CUDA sm_100a (%s)", h->gpu_name.c_str());
    return GAR_OK;
}

int32_t gar_get_stats(const gar_handle* h, int32_t stream, int32_t e, int64_t* in, int64_t* outp) {
    if (!h || stream < 0 || stream >= h->rows) return GAR_INVALID_CONFIG;
    if (h->multi()) {
        int local = 0;
        const Shard* sh = shard_of(h, stream, &local);
        return gar_get_stats(sh->h, local, e, in, outp);
    }
    const StreamState& st = h->eng.state(stream);
    if (e < 0 || e >= (int)st.samples_in.size()) return GAR_INVALID_CONFIG;
    if (in) *in = st.samples_in[(size_t)e];
    if (outp) *outp = st.samples_out[(size_t)e];
    return GAR_OK;
}

int32_t gar_num_stages(const gar_handle* h) { return (int32_t)h->eng.chain().stages.size(); }
int32_t gar_num_engines(const gar_handle* h) { return (int32_t)h->eng.chain().engines.size(); }

int32_t gar_describe_stage(const gar_handle* h, int32_t stream, int32_t stage, gar_stage_desc* out) {
    if (!h || !out || stream < 0 || stream >= h->rows) return GAR_INVALID_CONFIG;
    if (h->multi()) {
        int local = 0;
        const Shard* sh = shard_of(h, stream, &local);
        return gar_describe_stage(sh->h, local, stage, out);
    }
    const Chain& c = h->eng.chain();
    if (stage < 0 || stage >= (int)c.stages.size()) return GAR_INVALID_CONFIG;
    const StageDesign& s = c.stages[(size_t)stage];
    const StageState& st = h->eng.state(stream).st[(size_t)stage];
    out->kind = s.kind;
    out->engine_index = s.engine_index;
    out->factor = s.factor;
    out->taps = s.taps;
    out->proto_taps = s.proto_taps;
    out->engine_quality = s.quality;
    out->step = s.step;
    out->at = st.at;
    out->hist_len = s.kind == STAGE_CUBIC ? 0 : st.hist_len;
    out->decim_phase = st.decim_phase;
    out->ratio = s.ratio;
    return GAR_OK;
}

int32_t gar_plan_stage_type(const gar_handle* h, int32_t e) {
    const Chain& c = h->eng.chain();
    if (e < 0 || e >= (int)c.engines.size()) return -1;
    return c.engines[(size_t)e].plan_type;
}

int64_t gar_get_bank(const gar_handle* h, int32_t stage, int32_t which, double* out, int64_t cap) {
    const Chain& c = h->eng.chain();
    if (stage < 0 || stage >= (int)c.stages.size() || which < 0 || which > 3) return -1;
    const std::vector<double>& b = c.stages[(size_t)stage].bank[which];
    if ((int64_t)b.size() > cap) return -(int64_t)b.size();
    const bool f32 = h->compute_dtype == DT_F32 || h->wide_f32;
    for (size_t i = 0; i < b.size(); ++i) out[i] = f32 ? (double)(float)b[i] : b[i];
    return (int64_t)b.size();
}

int32_t gar_upload_bank(gar_handle* h, int32_t stage, int32_t which, const double* coef, int64_t n) {
    if (!h || !coef) return GAR_INVALID_CONFIG;
    const int rc = h->eng.set_bank(stage, which, coef, n, h->err);
    if (rc || !h->multi()) return rc;
    return for_shards(h, [=](Shard& sh) -> int { return gar_upload_bank(sh.h, stage, which, coef, n); });
}

int32_t gar_set_fusion(gar_handle* h, int32_t enabled) {
    if (!h) return GAR_INVALID_CONFIG;
    h->eng.set_fuse(enabled != 0);
    for (auto& sp : h->shards) sp->h->eng.set_fuse(enabled != 0);
    return GAR_OK;
}

int32_t gar_set_slice_budget(gar_handle* h, int64_t bytes) {
    if (!h) return GAR_INVALID_CONFIG;
    h->eng.set_slice_budget(bytes < 0 ? 0 : bytes);
    for (auto& sp : h->shards) sp->h->eng.set_slice_budget(bytes < 0 ? 0 : bytes);
    return GAR_OK;
}

void gar_set_tiled_polyphase(int32_t enabled) { gar::set_tiled_polyphase(enabled != 0); }
void gar_set_tensor_fir(int32_t enabled) { gar::set_tensor_fir(enabled != 0); }
void gar_set_chain_kernel(int32_t mode) { gar::set_chain_kernel(mode); }
int32_t gar_debug_chain_tile_hi(int32_t chunk_len, int32_t kp, int32_t n_chunks, int32_t n_tiles, int32_t chunk, int32_t hist_len,
                                int64_t L, int64_t at0, int64_t step, int32_t n_out) {
    return gar::chain_debug_tile_hi(chunk_len, kp, n_chunks, n_tiles, chunk, hist_len, L, at0, step, n_out);
}

int64_t gar_kernel_launches(const gar_handle* h, int32_t reset) {
    (void)h;  // process-wide counter: every <<<>>> of this library
    return (int64_t)launch_count(reset != 0);
}

int32_t gar_kernels_used(const gar_handle* h, char* buf, int32_t cap) {
    if (!h) return 0;
    std::string all;
    auto add = [&](const Engine& e) {
        for (const char* k : e.kernels_used()) {
            if (("," + all + ",").find(std::string(",") + k + ",") != std::string::npos) continue;
            if (!all.empty()) all += ",";
            all += k;
        }
    };
    add(h->eng);
    for (const auto& sp : h->shards) add(sp->h->eng);
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(all.size(), (size_t)cap - 1);
        std::memcpy(buf, all.data(), n);
        buf[n] = 0;
    }
    return (int32_t)all.size();
}

const char* gar_stage_kernel_name(const gar_handle* h, int32_t stage) {
    if (!h || stage < 0 || stage >= (int)h->eng.chain().stages.size()) return "";
    if (h->multi()) return h->shards[0]->h->eng.stage_dev(stage).kernel;
    return h->eng.stage_dev(stage).kernel;
}

// ---- processing --------------------------------------------------------------------------------

// rows [row0,row0+count) with per-row host pointers; io dtype may differ from the compute dtype
static int process_rows_host(gar_handle* h, int row0, int count, int io_dtype, const void* const* in,
                             const int64_t* n_in, void* const* out, int64_t out_cap, int64_t* n_out, bool flush,
                             bool check_estimate) {
    Engine& E = h->eng;
    if (row0 < 0 || count < 1 || row0 + count > h->rows) return fail(h, GAR_INVALID_CONFIG, "channel out of range");
    if (h->multi()) {
        // every row range is served by the shard(s) that own it; ErrBufferTooSmall must come before ANY state change
        // (constant.go:107-109), so the capacity check runs over all rows first
        for (int r = 0; r < count; ++r) {
            const int64_t n = flush ? 0 : n_in[r];
            if (n < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
            if (!flush && check_estimate && out_cap < gar_estimate_output(h, n))
                return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer smaller than EstimateOutput(len(input))");
            const int64_t c = flush ? gar_next_flush_count(h, row0 + r) : gar_next_output_count(h, row0 + r, n);
            if (c > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
        }
        return for_shards(h, [=](Shard& sh) -> int {
            const int a = std::max(row0, sh.row0), b = std::min(row0 + count, sh.row0 + sh.rows);
            if (a >= b) return 0;
            const int o = a - row0;
            return process_rows_host(sh.h, a - sh.row0, b - a, io_dtype, in ? in + o : nullptr, n_in ? n_in + o : nullptr, out + o,
                                     out_cap, n_out + o, flush, check_estimate);
        });
    }
    if (E.device() < 0) return fail(h, GAR_CUDA_ERROR, "geometry-only handle (device = -1) cannot process samples");
    int64_t max_in = 0;
    for (int r = 0; r < count; ++r) {
        const int64_t n = flush ? 0 : n_in[r];
        if (n < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
        if (!flush && check_estimate && out_cap < gar_estimate_output(h, n))
            return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer smaller than EstimateOutput(len(input))");
        max_in = n > max_in ? n : max_in;
    }
    // exact output counts first (pure integer), so nothing is advanced when a buffer is too small
    std::vector<int64_t> cnt((size_t)count, 0);
    int64_t max_out = 0;
    for (int r = 0; r < count; ++r) {
        cnt[(size_t)r] = flush ? gar_next_flush_count(h, row0 + r) : gar_next_output_count(h, row0 + r, n_in[r]);
        if (cnt[(size_t)r] > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
        max_out = cnt[(size_t)r] > max_out ? cnt[(size_t)r] : max_out;
    }
    // GAR_DEBUG_TUNING=1 GAR_TIMING=1: host-side breakdown of the per-channel call (staging / enqueue / wait / copy-out), printed
    // every 200 calls
    static const bool timing = [] { const char* e = gar::tune_env("GAR_TIMING"); return e && e[0] == '1'; }();
    static double t_acc[4] = {0, 0, 0, 0};
    static long t_calls = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_0 = now();
    cudaSetDevice(E.device());
    cudaStream_t s = E.stream();
    E.begin_on(s);  // the scratch slots may still be read by work a batch call enqueued on a caller's stream
    const size_t iosz = dsize(io_dtype), csz = dsize(h->compute_dtype);
    const int64_t in_stride = (max_in + 3) & ~int64_t(3), out_stride = (max_out + 3) & ~int64_t(3);
    bool cast = io_dtype != h->compute_dtype;
    // lock-step groups: consecutive rows with identical state and identical chunk length share launches
    struct Group { int r, run; };
    std::vector<Group> groups;
    for (int r = 0; r < count;) {
        int run = E.lockstep_run(row0 + r, row0 + count);
        if (!flush) {
            int k = 1;
            while (k < run && n_in[r + k] == n_in[r]) ++k;
            run = k;
        }
        groups.push_back({r, run});
        r += run;
    }
    // float32 I/O on a float64 engine (ProcessFloat32Into, constant.go:161-199): when every group is one streaming-size fused
    // launch, the kernel converts on load / store and the two cast launches disappear
    bool fold = cast && io_dtype == GAR_F32 && !flush;
    for (const Group& g : groups) fold = fold && E.io32_foldable(row0 + g.r, n_in[g.r], flush);
    if (fold) cast = false;
    // Streaming-size calls skip the copy engines: the chunk is placed in pinned, device-mapped staging that the kernels read
    // and write directly over PCIe (two cudaMemcpyAsync calls and their DMA start-up cost more than the 16 KB they move).
    const size_t in_bytes = (size_t)count * (size_t)in_stride * iosz, out_bytes = (size_t)count * (size_t)out_stride * iosz;
    static const size_t zc_max = [] { const char* e = gar::tune_env("GAR_ZC_MAX_KB"); return (size_t)(e ? std::atoi(e) : 2048) << 10; }();  // (8k -> 192k chunks: 786 KB out, 182 -> 106 us)
    const bool zero_copy = in_bytes + out_bytes <= zc_max;
    if (zero_copy && in_bytes + out_bytes > h->pin_cap) {
        if (h->pin) {
            cudaStreamSynchronize(s);
            cudaFreeHost(h->pin);
            h->pin = nullptr;
            h->pin_cap = 0;
        }
        const size_t want = std::max<size_t>(2 * (in_bytes + out_bytes), 128u << 10);
        if (cudaHostAlloc(&h->pin, want, cudaHostAllocMapped) != cudaSuccess) {
            h->pin = nullptr;
            cudaGetLastError();
        } else {
            h->pin_cap = want;
        }
    }
    const bool zc = zero_copy && h->pin != nullptr;
    char* d_in_io = nullptr;
    char* d_in_c = nullptr;
    if (max_in > 0) {
        d_in_io = zc ? (char*)h->pin : (char*)E.scratch(0, in_bytes, h->err);
        if (!d_in_io) return GAR_CUDA_ERROR;
        d_in_c = d_in_io;
        if (cast) {
            d_in_c = (char*)E.scratch(1, (size_t)count * (size_t)in_stride * csz, h->err);
            if (!d_in_c) return GAR_CUDA_ERROR;
        }
    }
    char* d_out_c = nullptr;
    char* d_out_io = nullptr;
    if (max_out > 0) {
        char* io_buf = zc ? (char*)h->pin + ((in_bytes + 255) & ~(size_t)255) : nullptr;
        if (cast) {
            d_out_c = (char*)E.scratch(2, (size_t)count * (size_t)out_stride * csz, h->err);
            d_out_io = zc ? io_buf : (char*)E.scratch(3, out_bytes, h->err);
        } else {
            d_out_c = d_out_io = zc ? io_buf : (char*)E.scratch(2, out_bytes, h->err);
        }
        if (!d_out_c || !d_out_io) return GAR_CUDA_ERROR;
    }
    if (!flush)
        for (int r = 0; r < count; ++r)
            if (n_in[r] > 0) {
                if (zc) std::memcpy(d_in_io + (size_t)r * (size_t)in_stride * iosz, in[r], (size_t)n_in[r] * iosz);
                else cudaMemcpyAsync(d_in_io + (size_t)r * (size_t)in_stride * iosz, in[r], (size_t)n_in[r] * iosz,
                                     cudaMemcpyHostToDevice, s);
            }
    if (cast && max_in > 0) {
        launch_cast(d_in_io, in_stride, io_dtype, d_in_c, in_stride, h->compute_dtype, (int32_t)max_in, count, s);
    }
    auto t_1 = now();
    const size_t esz_run = fold ? iosz : csz;  // element size of the buffers Engine::run sees
    for (const Group& g : groups) {
        int64_t got = 0;
        int rc = E.run(row0 + g.r, g.run, d_in_c ? d_in_c + (size_t)g.r * (size_t)in_stride * esz_run : nullptr, in_stride,
                       flush ? 0 : n_in[g.r], d_out_c ? d_out_c + (size_t)g.r * (size_t)out_stride * esz_run : nullptr,
                       out_stride, out_stride, flush, s, &got, h->err, fold);
        if (rc) return rc;
        for (int k = 0; k < g.run; ++k) n_out[g.r + k] = got;
    }
    if (cast && max_out > 0)
        launch_cast(d_out_c, out_stride, h->compute_dtype, d_out_io, out_stride, io_dtype, (int32_t)max_out, count, s);
    if (!zc)
        for (int q = 0; q < count; ++q)
            if (n_out[q] > 0)
                cudaMemcpyAsync(out[q], d_out_io + (size_t)q * (size_t)out_stride * iosz, (size_t)n_out[q] * iosz,
                                cudaMemcpyDeviceToHost, s);
    auto t_2 = now();
    // (a cudaStreamQuery spin instead of the blocking call measured no better: 26.6 against 25.0 us of waiting per 4096-frame
    // chunk, of which the kernel is 11 us — the rest is launch-to-start and completion-to-host latency of the platform)
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, GAR_CUDA_ERROR, std::string("stream sync: ") + cudaGetErrorString(e));
    auto t_3 = now();
    if (zc)
        for (int q = 0; q < count; ++q)
            if (n_out[q] > 0) std::memcpy(out[q], d_out_io + (size_t)q * (size_t)out_stride * iosz, (size_t)n_out[q] * iosz);
    if (timing) {
        auto t_4 = now();
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        t_acc[0] += us(t_0, t_1); t_acc[1] += us(t_1, t_2); t_acc[2] += us(t_2, t_3); t_acc[3] += us(t_3, t_4);
        if (++t_calls % 200 == 0) {
            std::fprintf(stderr, "[gar timing] per call over %ld calls: staging %.2f us, enqueue %.2f us, wait %.2f us, copy-out %.2f us\n",
                         t_calls, t_acc[0] / t_calls, t_acc[1] / t_calls, t_acc[2] / t_calls, t_acc[3] / t_calls);
        }
    }
    return GAR_OK;
}

int32_t gar_process_f64(gar_handle* h, int32_t ch, const double* in, int64_t n_in, double* out, int64_t out_cap,
                        int64_t* n_out) {
    if (!h || !n_out) return GAR_INVALID_CONFIG;
    *n_out = 0;
    if (h->compute_dtype != DT_F64 || h->wide_f32) return fail(h, GAR_NOT_SUPPORTED, "float32 engine: use gar_process_f32");
    if (n_in < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
    if (out_cap < gar_estimate_output(h, n_in)) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    if (n_in == 0) return GAR_OK;
    const void* ip = in;
    void* op = out;
    return process_rows_host(h, ch, 1, GAR_F64, &ip, &n_in, &op, out_cap, n_out, false, true);
}

int32_t gar_process_f32(gar_handle* h, int32_t ch, const float* in, int64_t n_in, float* out, int64_t out_cap,
                        int64_t* n_out) {
    if (!h || !n_out) return GAR_INVALID_CONFIG;
    *n_out = 0;
    if (h->cfg.path == GAR_PATH_ENGINE && h->compute_dtype != DT_F32 && !h->wide_f32)
        return fail(h, GAR_NOT_SUPPORTED, "float64 engine: use gar_process_f64");
    if (n_in < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
    if (out_cap < gar_estimate_output(h, n_in)) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    if (n_in == 0) return GAR_OK;
    const void* ip = in;
    void* op = out;
    return process_rows_host(h, ch, 1, GAR_F32, &ip, &n_in, &op, out_cap, n_out, false, true);
}

int32_t gar_process_multi_f64(gar_handle* h, const double* const* in, const int64_t* n_in, double* const* out,
                              int64_t out_cap, int64_t* n_out) {
    if (!h || !in || !n_in || !out || !n_out) return GAR_INVALID_CONFIG;
    if (h->compute_dtype != DT_F64 || h->wide_f32) return fail(h, GAR_NOT_SUPPORTED, "float32 engine has no ProcessMulti");
    const int C = h->cfg.channels;
    for (int c = 0; c < C; ++c) n_out[c] = 0;
    return process_rows_host(h, 0, C, GAR_F64, (const void* const*)in, n_in, (void* const*)out, out_cap, n_out, false,
                             false);
}

int32_t gar_flush_f64(gar_handle* h, int32_t ch, double* out, int64_t out_cap, int64_t* n_out) {
    if (!h || !n_out) return GAR_INVALID_CONFIG;
    *n_out = 0;
    if (h->compute_dtype != DT_F64 || h->wide_f32) return fail(h, GAR_NOT_SUPPORTED, "float32 engine: use gar_flush_f32");
    void* op = out;
    return process_rows_host(h, ch, 1, GAR_F64, nullptr, nullptr, &op, out_cap, n_out, true, false);
}

int32_t gar_flush_f32(gar_handle* h, int32_t ch, float* out, int64_t out_cap, int64_t* n_out) {
    if (!h || !n_out) return GAR_INVALID_CONFIG;
    *n_out = 0;
    if (h->cfg.path == GAR_PATH_ENGINE && h->compute_dtype != DT_F32 && !h->wide_f32)
        return fail(h, GAR_NOT_SUPPORTED, "float64 engine: use gar_flush_f64");
    void* op = out;
    return process_rows_host(h, ch, 1, GAR_F32, nullptr, nullptr, &op, out_cap, n_out, true, false);
}

int32_t gar_flush_multi_f64(gar_handle* h, double* const* out, int64_t out_cap, int64_t* n_out) {
    if (!h || !out || !n_out) return GAR_INVALID_CONFIG;
    if (h->compute_dtype != DT_F64 || h->wide_f32) return fail(h, GAR_NOT_SUPPORTED, "float32 engine has no FlushMulti");
    const int C = h->cfg.channels;
    for (int c = 0; c < C; ++c) n_out[c] = 0;
    return process_rows_host(h, 0, C, GAR_F64, nullptr, nullptr, (void* const*)out, out_cap, n_out, true, false);
}

int32_t gar_advance_geometry(gar_handle* h, int32_t stream, int64_t n_in, int32_t flush, int64_t* n_out) {
    if (!h || stream < 0 || stream >= h->rows || n_in < 0) return GAR_INVALID_CONFIG;
    if (h->multi()) {
        int local = 0;
        Shard* sh = shard_of(h, stream, &local);
        return gar_advance_geometry(sh->h, local, n_in, flush, n_out);
    }
    const int64_t n = h->eng.advance(stream, flush ? 0 : n_in, flush != 0);
    if (n_out) *n_out = n;
    return GAR_OK;
}

int32_t gar_reset(gar_handle* h) {
    if (!h) return GAR_INVALID_CONFIG;
    h->eng.reset_state();
    if (h->multi()) return for_shards(h, [](Shard& sh) -> int { return gar_reset(sh.h); });
    return GAR_OK;
}

// ---- batched streams --------------------------------------------------------------------------

static int batch_dev(gar_handle* h, int io_dtype, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out,
                     int64_t out_stride, int64_t out_cap, int64_t* n_out, bool flush, cudaStream_t s, int row0,
                     int count) {
    Engine& E = h->eng;
    const size_t iosz = dsize(io_dtype), csz = dsize(h->compute_dtype);
    const bool cast = io_dtype != h->compute_dtype;
    cudaSetDevice(E.device());
    if (cast) E.begin_on(s);  // cast scratch is shared handle state (Engine::run orders everything else itself)
    int r = 0;
    int64_t got_all = -1;
    while (r < count) {
        const int run = E.lockstep_run(row0 + r, row0 + count);
        const char* ip = d_in ? (const char*)d_in + (size_t)r * (size_t)in_stride * iosz : nullptr;
        char* op = d_out ? (char*)d_out + (size_t)r * (size_t)out_stride * iosz : nullptr;
        int64_t got = 0;
        int rc;
        if (!cast) {
            rc = E.run(row0 + r, run, ip, in_stride, flush ? 0 : n_in, op, out_stride, out_cap, flush, s, &got, h->err);
        } else if (io_dtype == GAR_F32 && !flush && E.io32_foldable(row0 + r, n_in, flush)) {
            // streaming-size fused launch: float32 samples converted on load / store, no cast launches
            rc = E.run(row0 + r, run, ip, in_stride, n_in, op, out_stride, out_cap, flush, s, &got, h->err, 1);
        } else if (io_dtype == GAR_F32 && !flush && E.pair32_foldable(row0 + r, run, n_in, in_stride)) {
            // large batched x2 -> polyphase call on the tensor cores: K1m widens the float32 windows in shared memory, K3p narrows
            // on the store — no cast launches, no float64 copies of the input and the output in HBM
            const int64_t want = gar_next_output_count(h, row0 + r, n_in);
            if (want > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
            rc = E.run(row0 + r, run, ip, in_stride, n_in, op, out_stride, out_cap, flush, s, &got, h->err, 2);
        } else {
            // I/O dtype differs from the compute dtype (path A with float32 I/O): cast through scratch
            const int64_t want = flush ? gar_next_flush_count(h, row0 + r) : gar_next_output_count(h, row0 + r, n_in);
            if (want > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
            const int64_t cis = (n_in + 3) & ~int64_t(3), cos = (want + 3) & ~int64_t(3);
            char* cin = nullptr;
            if (!flush && n_in > 0) {
                cin = (char*)E.scratch(1, (size_t)run * (size_t)cis * csz, h->err);
                if (!cin) return GAR_CUDA_ERROR;
                launch_cast(ip, in_stride, io_dtype, cin, cis, h->compute_dtype, (int32_t)n_in, run, s);
            }
            char* cout = nullptr;
            if (want > 0) {
                cout = (char*)E.scratch(2, (size_t)run * (size_t)cos * csz, h->err);
                if (!cout) return GAR_CUDA_ERROR;
            }
            rc = E.run(row0 + r, run, cin, cis, flush ? 0 : n_in, cout, cos, cos, flush, s, &got, h->err);
            if (!rc && got > 0) launch_cast(cout, cos, h->compute_dtype, op, out_stride, io_dtype, (int32_t)got, run, s);
        }
        if (rc) return rc;
        if (got_all < 0) got_all = got;
        else if (got != got_all)
            return fail(h, GAR_NOT_SUPPORTED, "batch rows are not in lock step (mixed per-channel calls before a batch call)");
        r += run;
    }
    if (cast) {
        E.end_on(s);
        if (cudaGetLastError() != cudaSuccess) return fail(h, GAR_CUDA_ERROR, "cast kernel launch failed");
    }
    if (n_out) *n_out = got_all < 0 ? 0 : got_all;
    return GAR_OK;
}

int32_t gar_process_batch_dev(gar_handle* h, int32_t io_dtype, const void* d_in, int64_t in_stride, int64_t n_in,
                              void* d_out, int64_t out_stride, int64_t out_cap, int64_t* n_out, void* cuda_stream) {
    if (!h) return GAR_INVALID_CONFIG;
    if (h->multi()) return fail(h, GAR_NOT_SUPPORTED, "device-pointer calls need a single-device handle (one per device)");
    if (n_in < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
    if (h->cfg.path == GAR_PATH_ENGINE && io_dtype != h->cfg.dtype)
        return fail(h, GAR_NOT_SUPPORTED, "engine handles take their own dtype");
    if (n_in == 0) {
        if (n_out) *n_out = 0;
        return GAR_OK;
    }
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->eng.stream();
    return batch_dev(h, io_dtype, d_in, in_stride, n_in, d_out, out_stride, out_cap, n_out, false, s, 0, h->rows);
}

int32_t gar_flush_batch_dev(gar_handle* h, int32_t io_dtype, void* d_out, int64_t out_stride, int64_t out_cap,
                            int64_t* n_out, void* cuda_stream) {
    if (!h) return GAR_INVALID_CONFIG;
    if (h->multi()) return fail(h, GAR_NOT_SUPPORTED, "device-pointer calls need a single-device handle (one per device)");
    if (h->cfg.path == GAR_PATH_ENGINE && io_dtype != h->cfg.dtype)
        return fail(h, GAR_NOT_SUPPORTED, "engine handles take their own dtype");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->eng.stream();
    return batch_dev(h, io_dtype, nullptr, 0, 0, d_out, out_stride, out_cap, n_out, true, s, 0, h->rows);
}

// Host-buffer batch: rows are cut into slices; slice k's H2D copy, kernels and D2H copy run on three
// streams chained by events, with two staging slots, so PCIe traffic in both directions overlaps the SMs.
static int batch_host(gar_handle* h, int io_dtype, const void* in, int64_t in_stride, int64_t n_in, void* out,
                      int64_t out_stride, int64_t out_cap, int64_t* n_out, bool flush) {
    Engine& E = h->eng;
    if (h->multi()) {
        // shard the rows over the devices: every worker runs the single-device pipeline (sliced H2D / kernels / D2H overlap)
        // on its own row block of the caller's buffers; all shards must agree on the count (lock step)
        const size_t isz = dsize(io_dtype);
        const int64_t want = flush ? gar_next_flush_count(h, 0) : gar_next_output_count(h, 0, n_in);
        for (auto& sp : h->shards) {
            const int64_t w = flush ? gar_next_flush_count(sp->h, 0) : gar_next_output_count(sp->h, 0, n_in);
            if (w != want || sp->h->eng.lockstep_run(0, sp->rows) != sp->rows)
                return fail(h, GAR_NOT_SUPPORTED, "batch rows are not in lock step (mixed per-channel calls before a batch call)");
        }
        if (want > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
        if (n_out) *n_out = want;
        if (!flush && n_in == 0) return GAR_OK;
        return for_shards(h, [=](Shard& sh) -> int {
            const char* ip = in ? (const char*)in + (size_t)sh.row0 * (size_t)in_stride * isz : nullptr;
            char* op = out ? (char*)out + (size_t)sh.row0 * (size_t)out_stride * isz : nullptr;
            return batch_host(sh.h, io_dtype, ip, in_stride, n_in, op, out_stride, out_cap, &sh.n_out, flush);
        });
    }
    if (E.device() < 0) return fail(h, GAR_CUDA_ERROR, "geometry-only handle (device = -1) cannot process samples");
    cudaSetDevice(E.device());
    const size_t iosz = dsize(io_dtype);
    const int rows = h->rows;
    // exact per-row output count; all rows must agree (lock step)
    const int64_t want = flush ? gar_next_flush_count(h, 0) : gar_next_output_count(h, 0, n_in);
    if (E.lockstep_run(0, rows) != rows)
        return fail(h, GAR_NOT_SUPPORTED, "batch rows are not in lock step (mixed per-channel calls before a batch call)");
    if (want > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    if (n_out) *n_out = want;
    if (!flush && n_in == 0) return GAR_OK;
    if (!h->s_in) {
        cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
        }
    }
    const int64_t is = (n_in + 3) & ~int64_t(3), os = (want + 3) & ~int64_t(3);
    // slice size: ~64 MiB of input per slice, at least 1 row, at most all rows
    const int64_t row_bytes = (int64_t)((flush ? 0 : n_in) + want) * (int64_t)iosz;
    int slice = (int)std::max<int64_t>(1, std::min<int64_t>(rows, (64ll << 20) / std::max<int64_t>(row_bytes, 1)));
    // a handful of long rows stays together (their time segments fill the tensor-core kernels' columns) up to 1 GiB
    if (rows <= 2 || (rows <= 8 && rows * row_bytes <= (1ll << 30))) slice = rows;
    const size_t need_in = (size_t)slice * (size_t)is * iosz, need_out = (size_t)slice * (size_t)os * iosz;
    auto grow = [&](void* (&slot)[2], size_t& cap, size_t need, const char* what) -> bool {
        if (need <= cap) return true;
        cudaDeviceSynchronize();
        cap = 0;  // nothing is published until both slots exist
        for (int i = 0; i < 2; ++i) {
            if (slot[i]) cudaFree(slot[i]);
            slot[i] = nullptr;
        }
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(&slot[i], need) != cudaSuccess) {
                slot[i] = nullptr;
                fail(h, GAR_CUDA_ERROR, what);
                return false;
            }
        cap = need;
        return true;
    };
    if (!grow(h->slot_in, h->slot_in_cap, need_in, "cudaMalloc(staging in)")) return GAR_CUDA_ERROR;
    if (!grow(h->slot_out, h->slot_out_cap, need_out, "cudaMalloc(staging out)")) return GAR_CUDA_ERROR;
    cudaStream_t sc = E.stream();
    int k = 0;
    for (int r0 = 0; r0 < rows; r0 += slice, ++k) {
        const int cnt = std::min(slice, rows - r0);
        const int slot = k & 1;
        if (!flush) {
            if (k >= 2) cudaStreamWaitEvent(h->s_in, h->ev_comp[slot], 0);  // slot's previous kernels done reading
            cudaMemcpy2DAsync(h->slot_in[slot], (size_t)is * iosz, (const char*)in + (size_t)r0 * (size_t)in_stride * iosz,
                              (size_t)in_stride * iosz, (size_t)n_in * iosz, (size_t)cnt, cudaMemcpyHostToDevice, h->s_in);
            cudaEventRecord(h->ev_in[slot], h->s_in);
            cudaStreamWaitEvent(sc, h->ev_in[slot], 0);
        }
        if (k >= 2) cudaStreamWaitEvent(sc, h->ev_out[slot], 0);  // slot's previous D2H done
        int64_t got = 0;
        int rc = batch_dev(h, io_dtype, h->slot_in[slot], is, n_in, h->slot_out[slot], os, os, &got, flush, sc, r0, cnt);
        if (rc) {
            cudaDeviceSynchronize();
            return rc;
        }
        cudaEventRecord(h->ev_comp[slot], sc);
        if (got > 0) {
            cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0);
            cudaMemcpy2DAsync((char*)out + (size_t)r0 * (size_t)out_stride * iosz, (size_t)out_stride * iosz,
                              h->slot_out[slot], (size_t)os * iosz, (size_t)got * iosz, (size_t)cnt,
                              cudaMemcpyDeviceToHost, h->s_out);
        }
        cudaEventRecord(h->ev_out[slot], h->s_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->s_in);
    cudaError_t e2 = cudaStreamSynchronize(sc);
    cudaError_t e3 = cudaStreamSynchronize(h->s_out);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        return fail(h, GAR_CUDA_ERROR, std::string("batch sync: ") + cudaGetErrorString(e));
    }
    return GAR_OK;
}

int32_t gar_process_batch(gar_handle* h, int32_t io_dtype, const void* in, int64_t in_stride, int64_t n_in, void* out,
                          int64_t out_stride, int64_t out_cap, int64_t* n_out) {
    if (!h) return GAR_INVALID_CONFIG;
    if (n_in < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
    if (h->cfg.path == GAR_PATH_ENGINE && io_dtype != h->cfg.dtype)
        return fail(h, GAR_NOT_SUPPORTED, "engine handles take their own dtype");
    if (out_cap < gar_estimate_output(h, n_in)) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    return batch_host(h, io_dtype, in, in_stride, n_in, out, out_stride, out_cap, n_out, false);
}

int32_t gar_flush_batch(gar_handle* h, int32_t io_dtype, void* out, int64_t out_stride, int64_t out_cap,
                        int64_t* n_out) {
    if (!h) return GAR_INVALID_CONFIG;
    if (h->cfg.path == GAR_PATH_ENGINE && io_dtype != h->cfg.dtype)
        return fail(h, GAR_NOT_SUPPORTED, "engine handles take their own dtype");
    return batch_host(h, io_dtype, nullptr, 0, 0, out, out_stride, out_cap, n_out, true);
}


// ---- interleaved / integer-PCM boundary (N1) -------------------------------------------------------
static size_t fmt_size(int fmt) {
    switch (fmt) {
        case GAR_FMT_F64: case GAR_FMT_I64: return 8;
        case GAR_FMT_F32: case GAR_FMT_I32: return 4;
        default: return 2;
    }
}
static double pcm_max(int bit_depth) {  // cmd/resample-wav/main.go:54-56,429-440
    switch (bit_depth) {
        case 24: return 8388607.0;
        case 32: return 2147483647.0;
        default: return 32767.0;
    }
}

static int interleaved_call(gar_handle* h, int fmt, int bit_depth, const void* in, int64_t n_frames, void* out,
                            int64_t out_cap, int64_t* n_frames_out, bool flush) {
    Engine& E = h->eng;
    if (fmt < GAR_FMT_F64 || fmt > GAR_FMT_I64) return fail(h, GAR_INVALID_CONFIG, "unknown sample format");
    if (h->multi()) return fail(h, GAR_NOT_SUPPORTED, "interleaved calls need a single-device handle");
    if (E.device() < 0) return fail(h, GAR_CUDA_ERROR, "geometry-only handle (device = -1) cannot process samples");
    const int C = h->rows;
    if (E.lockstep_run(0, C) != C) return fail(h, GAR_NOT_SUPPORTED, "channels are not in lock step");
    const int64_t want = flush ? gar_next_flush_count(h, 0) : gar_next_output_count(h, 0, n_frames);
    if (n_frames_out) *n_frames_out = 0;
    if (!flush && out_cap < gar_estimate_output(h, n_frames)) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    if (want > out_cap) return fail(h, GAR_BUFFER_TOO_SMALL, "output buffer too small");
    if (!flush && n_frames == 0) return GAR_OK;
    cudaSetDevice(E.device());
    cudaStream_t s = E.stream();
    E.begin_on(s);
    const size_t isz = fmt_size(fmt), csz = dsize(h->compute_dtype);
    const bool is_int = fmt >= GAR_FMT_I16;
    const double maxv = is_int ? pcm_max(bit_depth) : 0.0;
    const int64_t is = (n_frames + 3) & ~int64_t(3), os = (want + 3) & ~int64_t(3);
    char *d_il_in = nullptr, *d_pl_in = nullptr, *d_pl_out = nullptr, *d_il_out = nullptr;
    if (!flush) {
        d_il_in = (char*)E.scratch(0, (size_t)C * (size_t)n_frames * isz, h->err);
        d_pl_in = (char*)E.scratch(1, (size_t)C * (size_t)is * csz, h->err);
        if (!d_il_in || !d_pl_in) return GAR_CUDA_ERROR;
        cudaMemcpyAsync(d_il_in, in, (size_t)C * (size_t)n_frames * isz, cudaMemcpyHostToDevice, s);
        launch_deinterleave(d_il_in, fmt, C, n_frames, d_pl_in, is, h->compute_dtype, is_int ? 1.0 / maxv : 0.0, h->wide_f32 ? 1 : 0, s);
    }
    if (want > 0) {
        d_pl_out = (char*)E.scratch(2, (size_t)C * (size_t)os * csz, h->err);
        d_il_out = (char*)E.scratch(3, (size_t)C * (size_t)want * isz, h->err);
        if (!d_pl_out || !d_il_out) return GAR_CUDA_ERROR;
    }
    int64_t got = 0;
    int rc = E.run(0, C, d_pl_in, is, flush ? 0 : n_frames, d_pl_out, os, os, flush, s, &got, h->err);
    if (rc) return rc;
    if (got > 0) {
        launch_interleave(d_pl_out, os, h->compute_dtype, C, got, d_il_out, fmt, maxv, h->wide_f32 ? 1 : 0, s);
        cudaMemcpyAsync(out, d_il_out, (size_t)C * (size_t)got * isz, cudaMemcpyDeviceToHost, s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(h, GAR_CUDA_ERROR, std::string("stream sync: ") + cudaGetErrorString(e));
    if (n_frames_out) *n_frames_out = got;
    return GAR_OK;
}

int32_t gar_process_interleaved(gar_handle* h, int32_t fmt, int32_t bit_depth, const void* in, int64_t n_frames,
                                void* out, int64_t out_cap_frames, int64_t* n_frames_out) {
    if (!h) return GAR_INVALID_CONFIG;
    if (n_frames < 0) return fail(h, GAR_INVALID_CONFIG, "negative input length");
    return interleaved_call(h, fmt, bit_depth, in, n_frames, out, out_cap_frames, n_frames_out, false);
}

int32_t gar_flush_interleaved(gar_handle* h, int32_t fmt, int32_t bit_depth, void* out, int64_t out_cap_frames,
                              int64_t* n_frames_out) {
    if (!h) return GAR_INVALID_CONFIG;
    return interleaved_call(h, fmt, bit_depth, nullptr, 0, out, out_cap_frames, n_frames_out, true);
}

// ---- utilities ---------------------------------------------------------------------------------
void* gar_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void* gar_host_alloc_wc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void gar_host_free(void* p) {
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_reg_mutex);
        auto it = g_registered.find(p);
        if (it != g_registered.end()) {
            bytes = it->second;
            g_registered.erase(it);
        }
    }
    if (bytes) {
        cudaHostUnregister(p);
        munmap(p, bytes);
        return;
    }
    cudaFreeHost(p);
}

void* gar_host_alloc_rows(gar_handle* h, size_t row_bytes) {
    if (!h || row_bytes == 0 || h->rows <= 0) return nullptr;
    const size_t page = 4096;
    const size_t total = (((size_t)h->rows * row_bytes) + page - 1) / page * page;
    void* base = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    // first touch decides the NUMA node of a page: every shard's rows are touched by a thread bound to its device's node
    if (h->multi()) {
        for_shards(h, [=](Shard& sh) -> int {
            std::memset((char*)base + (size_t)sh.row0 * row_bytes, 0, (size_t)sh.rows * row_bytes);
            return 0;
        });
    } else {
        const int dev = h->eng.device();
        std::thread t([=] {
            if (dev >= 0) bind_thread_to_device(dev);
            std::memset(base, 0, total);
        });
        t.join();
    }
    if (h->eng.device() >= 0 || h->multi()) cudaSetDevice(h->multi() ? h->shards[0]->device : h->eng.device());
    if (cudaHostRegister(base, total, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(base, total);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    g_registered[base] = total;
    return base;
}

int32_t gar_memcpy_async(void* dst, const void* src, size_t bytes, int32_t kind, void* cuda_stream) {
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDefault);
    return cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)cuda_stream) == cudaSuccess ? GAR_OK : GAR_CUDA_ERROR;
}

int32_t gar_bind_thread_to_device(int32_t device) { return bind_thread_to_device(device); }
int32_t gar_device_numa_node(int32_t device) { return device_numa_node(device); }

int32_t gar_measure_fma_peak(int32_t device, int32_t dtype, double* tflops) {
    if (!tflops) return GAR_INVALID_CONFIG;
    if (cudaSetDevice(device) != cudaSuccess) return GAR_CUDA_ERROR;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        double flops = 0;
        // dtype 2 = packed fp32x2 (FFMA2) probe, 3 = FP64 tensor cores (DMMA.8x8x4)
        const float ms = run_fma_probe(dtype == 2 || dtype == 3 ? dtype : (dtype == GAR_F32 ? DT_F32 : DT_F64),
                                       dtype == 3 ? 512 : (dtype == GAR_F64 ? 2048 : 4096), &flops, 0);
        if (ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    if (cudaGetLastError() != cudaSuccess || best <= 0) return GAR_CUDA_ERROR;
    *tflops = best;
    return GAR_OK;
}

}  // extern "C"
