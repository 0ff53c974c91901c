// gar.hpp — header-only C++ mirror of the reference's Go API over the C ABI in gar.h.
//
// The reference is compiled Go; no Go toolchain exists in the build image, so this is the compiled-language
// host side that is exercised here (the cgo equivalent is go-audio-resampler_b200/go/b200/b200.go).
// Names, argument meaning and error behaviour follow resample.go / constant.go / convenience.go.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gar.h"

namespace resampler {

struct ErrInvalidConfig : std::invalid_argument { using std::invalid_argument::invalid_argument; };   // resample.go:158
struct ErrBufferTooSmall : std::length_error { using std::length_error::length_error; };               // resample.go:161
struct ErrNotSupported : std::logic_error { using std::logic_error::logic_error; };                    // resample.go:164
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

enum QualityPreset { QualityQuick = 0, QualityLow, QualityMedium, QualityHigh, QualityVeryHigh, QualityCustom };

struct QualitySpec {  // resample.go:77-102
    QualityPreset Preset = QualityMedium;
    int Precision = 0;
    double PhaseResponse = 0, PassbandEnd = 0, StopbandBegin = 0;
    uint32_t Flags = 0;
};

struct Config {  // resample.go:46-73
    double InputRate = 0, OutputRate = 0;
    int Channels = 0;
    QualitySpec Quality;
    int MaxInputSize = 0;
    bool EnableSIMD = false, EnableParallel = false;
    int Device = 0;  // extension: CUDA ordinal (-1: geometry-only handle)
};

namespace detail {
inline void check(int32_t st, const gar_handle* h) {
    if (st == GAR_OK) return;
    const std::string msg = gar_last_error(h);
    switch (st) {
        case GAR_INVALID_CONFIG: throw ErrInvalidConfig(msg);
        case GAR_BUFFER_TOO_SMALL: throw ErrBufferTooSmall(msg);
        case GAR_NOT_SUPPORTED: throw ErrNotSupported(msg);
        case GAR_CUDA_ERROR: throw CudaError(msg);
        default: throw std::runtime_error(msg);
    }
}
struct Deleter { void operator()(gar_handle* h) const { gar_destroy(h); } };
using Handle = std::unique_ptr<gar_handle, Deleter>;
inline Handle create(const gar_config& c) {
    gar_handle* h = nullptr;
    check(gar_create(&c, &h), nullptr);
    return Handle(h);
}
}  // namespace detail

class Base {
  public:
    int EstimateOutput(int n) const { return (int)gar_estimate_output(h_.get(), n); }   // constant.go:117-119
    double GetRatio() const { return gar_get_ratio(h_.get()); }
    int GetLatency() const { return gar_get_latency(h_.get()); }
    void Reset() { gar_reset(h_.get()); }
    gar_handle* raw() const { return h_.get(); }
  protected:
    explicit Base(detail::Handle h) : h_(std::move(h)) {}
    template <class T, class F>
    std::vector<T> owned(const std::vector<T>& in, F call) {
        if (in.empty()) return {};
        int64_t cap = gar_next_output_count(h_.get(), 0, (int64_t)in.size());
        if (cap < EstimateOutput((int)in.size())) cap = EstimateOutput((int)in.size());
        std::vector<T> out((size_t)cap);
        int64_t n = 0;
        detail::check(call(h_.get(), 0, in.data(), (int64_t)in.size(), out.data(), cap, &n), h_.get());
        out.resize((size_t)n);
        return out;
    }
    detail::Handle h_;
};

// constantRateResampler behind New(Config) (constant.go:16-485)
class Resampler : public Base {
  public:
    explicit Resampler(const Config& c) : Base(detail::create(to_c(c))), channels_(c.Channels) {}
    std::vector<double> Process(const std::vector<double>& in) { return owned<double>(in, gar_process_f64); }
    std::vector<float> ProcessFloat32(const std::vector<float>& in) { return owned<float>(in, gar_process_f32); }
    int ProcessInto(const double* in, int n, double* out, int cap) {   // constant.go:103-112
        int64_t got = 0;
        detail::check(gar_process_f64(h_.get(), 0, in, n, out, cap, &got), h_.get());
        return (int)got;
    }
    int ProcessFloat32Into(const float* in, int n, float* out, int cap) {   // constant.go:161-199
        int64_t got = 0;
        detail::check(gar_process_f32(h_.get(), 0, in, n, out, cap, &got), h_.get());
        return (int)got;
    }
    std::vector<std::vector<double>> ProcessMulti(const std::vector<std::vector<double>>& in) {   // constant.go:204-252
        if ((int)in.size() != channels_) throw std::invalid_argument("wrong channel count");
        std::vector<const double*> ip(in.size());
        std::vector<int64_t> nin(in.size()), nout(in.size());
        int64_t cap = 1;
        for (size_t c = 0; c < in.size(); ++c) {
            ip[c] = in[c].data();
            nin[c] = (int64_t)in[c].size();
            cap = std::max(cap, gar_next_output_count(h_.get(), (int32_t)c, nin[c]));
        }
        std::vector<std::vector<double>> out(in.size(), std::vector<double>((size_t)cap));
        std::vector<double*> op(in.size());
        for (size_t c = 0; c < in.size(); ++c) op[c] = out[c].data();
        detail::check(gar_process_multi_f64(h_.get(), ip.data(), nin.data(), op.data(), cap, nout.data()), h_.get());
        for (size_t c = 0; c < in.size(); ++c) out[c].resize((size_t)nout[c]);
        return out;
    }
    std::vector<double> Flush() {   // channel 0 only, constant.go:349-354
        std::vector<double> out((size_t)std::max<int64_t>(1, gar_next_flush_count(h_.get(), 0)));
        int64_t n = 0;
        detail::check(gar_flush_f64(h_.get(), 0, out.data(), (int64_t)out.size(), &n), h_.get());
        out.resize((size_t)n);
        return out;
    }
  private:
    static gar_config to_c(const Config& c) {
        gar_config g{};
        g.input_rate = c.InputRate; g.output_rate = c.OutputRate; g.channels = c.Channels;
        g.path = GAR_PATH_PIPELINE; g.preset = c.Quality.Preset; g.custom_precision = c.Quality.Precision;
        g.custom_phase_response = c.Quality.PhaseResponse; g.custom_passband_end = c.Quality.PassbandEnd;
        g.custom_stopband_begin = c.Quality.StopbandBegin; g.dtype = GAR_F64; g.engine_quality = -1;
        g.device = c.Device; g.max_input_size = c.MaxInputSize;
        g.flags = c.Quality.Flags | (c.EnableParallel ? 1u << 16 : 0u);
        return g;
    }
    int channels_;
};

inline std::unique_ptr<Resampler> New(const Config& c) { return std::make_unique<Resampler>(c); }   // resample.go:272

// SimpleResampler[Float32] (convenience.go:118-186, 315-395)
template <class T>
class SimpleResamplerT : public Base {
  public:
    SimpleResamplerT(double in, double out, QualityPreset q, int device = 0)
        : Base(detail::create(cfg(in, out, q, device))) {}
    std::vector<T> Process(const std::vector<T>& in) {
        if constexpr (sizeof(T) == 4) return this->template owned<float>(in, gar_process_f32);
        else return this->template owned<double>(in, gar_process_f64);
    }
    std::vector<T> Flush() {
        std::vector<T> out((size_t)std::max<int64_t>(1, gar_next_flush_count(h_.get(), 0)));
        int64_t n = 0;
        if constexpr (sizeof(T) == 4) detail::check(gar_flush_f32(h_.get(), 0, out.data(), (int64_t)out.size(), &n), h_.get());
        else detail::check(gar_flush_f64(h_.get(), 0, out.data(), (int64_t)out.size(), &n), h_.get());
        out.resize((size_t)n);
        return out;
    }
  private:
    static gar_config cfg(double in, double out, QualityPreset q, int device) {
        gar_config g{};
        g.input_rate = in; g.output_rate = out; g.channels = 1; g.path = GAR_PATH_ENGINE; g.preset = q;
        g.dtype = sizeof(T) == 4 ? GAR_F32 : GAR_F64; g.engine_quality = -1; g.device = device;
        return g;
    }
};
using SimpleResampler = SimpleResamplerT<double>;
using SimpleResamplerFloat32 = SimpleResamplerT<float>;

inline std::vector<double> ResampleMono(const std::vector<double>& in, double ir, double orate, QualityPreset q) {   // convenience.go:204-229
    SimpleResampler r(ir, orate, q);
    auto a = r.Process(in);
    auto b = r.Flush();
    a.insert(a.end(), b.begin(), b.end());
    return a;
}
inline std::vector<float> ResampleMonoFloat32(const std::vector<float>& in, double ir, double orate, QualityPreset q) {   // convenience.go:407-429
    SimpleResamplerFloat32 r(ir, orate, q);
    auto a = r.Process(in);
    auto b = r.Flush();
    a.insert(a.end(), b.begin(), b.end());
    return a;
}

}  // namespace resampler
