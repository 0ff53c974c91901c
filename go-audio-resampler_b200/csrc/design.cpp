// design.cpp — see design.hpp. Host-only; runs once per configuration.
#include "design.hpp"

#include <cmath>

namespace gar {

namespace {

// f64.Sum of tphakala/simd (un-vendored): 4-lane (AVX2) partial sums assumed.
// Only a relative 1e-16 effect on every coefficient; kept identical to the
// test oracle so banks can be compared bit for bit.
double lane_sum(const std::vector<double>& v) {
    double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    size_t i = 0, n = v.size();
    for (; i + 4 <= n; i += 4) {
        l0 += v[i];
        l1 += v[i + 1];
        l2 += v[i + 2];
        l3 += v[i + 3];
    }
    double s = ((l0 + l1) + l2) + l3;
    for (; i < n; ++i) s += v[i];
    return s;
}

void scale_in_place(std::vector<double>& v, double s) {
    for (double& x : v) x *= s;
}

// Abramowitz & Stegun 9.8.1 / 9.8.2 coefficient tables (mathutil/constants.go:21-41).
const double kI0Small[6] = {3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.360768e-1, 0.45813e-2};
const double kI0Large[9] = {0.39894228,  0.1328592e-1,  0.225319e-2,  -0.157565e-2, 0.916281e-2,
                            -0.2057706e-1, 0.2635537e-1, -0.1647633e-1, 0.392377e-2};

}  // namespace

double bessel_i0(double x) {  // mathutil/bessel.go:22-49
    const double ax = std::fabs(x);
    if (ax < 3.75) {
        double t = x / 3.75;
        t *= t;
        double p = kI0Small[5];
        for (int i = 4; i >= 0; --i) p = kI0Small[i] + t * p;  // Horner, same nesting as the Go expression
        return 1.0 + t * p;
    }
    const double t = 3.75 / ax;
    double p = kI0Large[8];
    for (int i = 7; i >= 0; --i) p = kI0Large[i] + t * p;
    return std::exp(ax) * p / std::sqrt(ax);
}

double kaiser_beta(double att) {  // mathutil/bessel.go:126-134
    if (att > 50.0) return 0.1102 * (att - 8.7);
    if (att >= 21.0) {
        const double d = att - 21.0;
        return 0.5842 * std::pow(d, 0.4) + 0.07886 * d;
    }
    return 0.0;
}

int estimate_filter_length(double att, double tbw) {  // mathutil/bessel.go:245-268
    if (tbw <= 0) tbw = 0.01;
    // Go folds 2.285*2.0*math.Pi as an exact untyped constant -> 14.357078426905355.
    const double kaiser_den = 14.357078426905355;
    int taps = (int)std::ceil((att - 8.0) / (kaiser_den * tbw));
    if ((taps & 1) == 0) ++taps;
    if (taps < 3) taps = 3;
    if (taps > 8191) taps = 8191;
    return taps;
}

// filter/kaiser.go:47-91 (window) + :159-203 (windowed sinc, DC normalisation)
bool design_lowpass(int taps, double fc, double att, double gain, std::vector<double>& h) {
    if (taps < 3 || taps > 8191) return false;       // kaiser.go:111-118
    if (!(fc > 0.0) || !(fc < 0.5)) return false;    // :120-122
    if (att < 0.0 || att > 500.0) return false;      // :124-130
    if (!(gain > 0.0)) return false;                 // :132-134
    const double beta = std::fabs(kaiser_beta(att));
    const double alpha = (double)(taps - 1) / 2.0;
    const double i0_beta = bessel_i0(beta);
    const double two_pi = 2.0 * M_PI;
    h.assign((size_t)taps, 0.0);
    for (int n = 0; n < taps; ++n) {
        // window
        const double xw = ((double)n - alpha) / alpha;
        const double arg = beta * std::sqrt(1.0 - xw * xw);
        const double i0_arg = bessel_i0(arg);
        double w;
        if (std::isinf(i0_arg) && i0_arg > 0 && std::isinf(i0_beta) && i0_beta > 0)
            w = std::exp(arg - beta);
        else
            w = i0_arg / i0_beta;
        // sinc
        const double x = (double)n - alpha;
        double s;
        if (std::fabs(x) < 1e-10) {
            s = 2.0 * fc;
        } else {
            const double a = two_pi * fc * x;
            s = std::sin(a) / (M_PI * x);
        }
        h[(size_t)n] = s * w;
    }
    const double sum = lane_sum(h);
    if (std::fabs(sum) > 1e-10) scale_in_place(h, gain / sum);
    return true;
}

double quality_attenuation(int q) {  // filter_params.go:150-175; (bits+1)*6.0206 folded exactly by Go
    switch (q) {
        case EQ_QUICK: return 54.1854;
        case EQ_LOW: case EQ_MEDIUM: case EQ_16BIT: return 102.3502;
        case EQ_HIGH: case EQ_20BIT: return 126.4326;
        case EQ_VERYHIGH: case EQ_28BIT: return 174.5974;
        case EQ_24BIT: return 150.515;
        case EQ_32BIT: return 198.6798;
        default: return 126.4326;
    }
}

double quality_passband_end(int q) {  // filter_params.go:180-195
    switch (q) {
        case EQ_QUICK: case EQ_LOW: case EQ_16BIT: return 0.67625;
        case EQ_MEDIUM: return 0.91;
        case EQ_HIGH: case EQ_20BIT: return 0.912;
        case EQ_VERYHIGH: case EQ_24BIT: case EQ_28BIT: case EQ_32BIT: return 0.913;
        default: return 0.912;
    }
}

bool is_integer_ratio(double r) {  // resampler.go:356-360
    const double k = std::round(r);
    return std::fabs(r - k) < 1e-9 && k >= 1.0;
}

namespace {

// ---- integer up-sampler bank (dft_stage.go:50-101) --------------------------------------
bool design_up(int factor, int quality, StageDesign& st, std::string& err) {
    st.kind = STAGE_UP;
    st.quality = quality;
    st.factor = factor;
    st.ratio = (double)factor;
    const double fc = 0.4778321 / (double)factor;
    const double tbw = 0.05 / (double)factor;
    const double att = quality_attenuation(quality);
    std::vector<double> h;
    if (!design_lowpass(estimate_filter_length(att, tbw), fc, att, 1.0, h)) {
        err = "failed to design DFT filter";
        return false;
    }
    st.proto_taps = (int)h.size();
    st.taps = ((int)h.size() + factor - 1) / factor;
    st.bank[0].assign((size_t)factor * st.taps, 0.0);
    for (int p = 0; p < factor; ++p)
        for (int t = 0; t < st.taps; ++t) {
            const int src = t * factor + p;
            if (src < (int)h.size()) st.bank[0][(size_t)p * st.taps + (st.taps - 1 - t)] = h[(size_t)src] * (double)factor;
        }
    return true;
}

// ---- integer decimator taps (dft_stage.go:401-475) ---------------------------------------
bool design_decim(int factor, int quality, StageDesign& st, std::string& err) {
    st.kind = STAGE_DECIM;
    st.quality = quality;
    st.factor = factor;
    st.ratio = 1.0 / (double)factor;
    const double fp_n = quality_passband_end(quality) / (double)factor;
    const double fs_n = 1.0 / (double)factor;
    const double tr = 0.5 * (fs_n - fp_n);
    const double fc = fs_n - tr;
    const double att = quality_attenuation(quality);
    std::vector<double> h;
    if (!design_lowpass(estimate_filter_length(att, tr * 0.5), fc * 0.5, att, 1.0, h)) {
        err = "failed to design decimation filter";
        return false;
    }
    st.proto_taps = st.taps = (int)h.size();
    st.bank[0].assign(h.size(), 0.0);
    for (size_t i = 0; i < h.size(); ++i) st.bank[0][h.size() - 1 - i] = h[i];
    return true;
}

// ---- soxr-style inverse response (filter_params.go:355-394) -------------------------------
double lsx_inv_f_resp(double drop, double a) {
    if (a < 1.0) a = 1.0;
    else if (a > 300.0) a = 300.0;
    double x = ((2.0517e-07 * a + -1.1303e-04) * a + 0.023154) * a + 0.55924;
    const double lin = std::exp(drop * M_LN10 * 0.05);
    const double s = lin > 0.5 ? 1 - lin : lin;
    double sv = std::sin(x * 0.5);
    if (sv <= 1e-10) sv = 1e-10;
    const double sine_pow = std::log(0.5) / std::log(sv);
    x = std::asin(std::pow(s, 1.0 / sine_pow)) / x;
    return lin > 0.5 ? x : 1 - x;
}

// ---- arbitrary-ratio polyphase bank (filter_params.go:229-329,446-630; polyphase_stage.go:69-154)
bool design_poly(double ratio, double total_io, bool has_pre, int quality, StageDesign& st, std::string& err) {
    if (!(ratio > 0)) {
        err = "ratio must be positive";
        return false;
    }
    st.kind = STAGE_POLY;
    st.quality = quality;
    st.ratio = ratio;
    // findRationalApprox: L in [64,256], default 80
    const double inv = 1.0 / ratio;
    int L = 80;
    double best = std::fabs((double)(int)std::round(inv * 80.0) / 80.0 - inv);
    for (int cand = 64; cand <= 256; ++cand) {
        const int cs = (int)std::round(inv * (double)cand);
        if (cs <= 0) continue;
        const double e = std::fabs((double)cs / (double)cand - inv);
        if (e < best) {
            L = cand;
            best = e;
        }
        if (best < 1e-10) break;
    }
    st.factor = L;
    const double att = quality_attenuation(quality);
    const double pb = quality_passband_end(quality);
    const double phases = (double)L;

    // ComputePolyphaseFilterParams
    const bool up = total_io < 1.0;
    const double mult = up ? 1.0 : total_io;
    double fp1, fs1;
    if (up) {
        fp1 = total_io * pb;
        fs1 = total_io * 1.0;
    } else {
        fp1 = pb * ratio;
        fs1 = ratio;
    }
    double fn, fs_raw, fp_raw = fp1;
    if (!up && has_pre) {
        fn = 2.0 * mult;
        fs_raw = 3.0 + std::fabs(fs1 - 1.0);
    } else {
        fn = 1.0;
        fs_raw = 2.0 - (fp1 + (fs1 - fp1) * 0.7);
    }
    const double ifr = lsx_inv_f_resp(-0.01, att);
    if (ifr < 0.999) {
        const double adj = fs_raw - (fs_raw - fp_raw) / (1.0 - ifr);
        if (adj > 0 && adj < fs_raw) fp_raw = adj;
    }
    const double fp = fp_raw / std::fabs(fn);
    const double fs = fs_raw / std::fabs(fn);
    double tr = 0.5 * (fs - fp);
    tr /= phases;
    const double tr_lim = 0.5 * fs / phases;
    if (tr > tr_lim) tr = tr_lim;
    if (tr < 0.001) tr = 0.001;
    double fc = fs / phases - tr;
    if (fc < 0.001) fc = 0.001;

    const int lib_limit = 8190;
    int cap;
    if (att < 110.0) cap = 32;
    else if (att < 130.0) cap = 64;
    else if (att < 160.0) cap = 100;
    else cap = (lib_limit + 1) / L;
    const int ideal = (int)std::ceil(att / tr + 1);
    int tpp = (ideal + L - 1) / L;
    if (tpp < 8) tpp = 8;
    else if (tpp > cap) tpp = cap;
    int total = L * tpp - 1;
    if (total > lib_limit) {
        tpp = (lib_limit + 1) / L;
        if (tpp < 8) tpp = 8;
        total = L * tpp - 1;
    }
    st.taps = tpp;
    st.proto_taps = total;

    double cutoff = fc / 2.0;
    if (cutoff <= 0) cutoff = 0.001;
    if (cutoff >= 0.5) cutoff = 0.499;
    std::vector<double> proto;
    if (!design_lowpass(total, cutoff, att, 1.0, proto)) {
        err = "failed to design prototype filter";
        return false;
    }
    const double sum = lane_sum(proto);
    if (sum != 0) scale_in_place(proto, (double)L / sum);

    // coefficient (tap, phase) with the reference's same-tap phase wrap
    auto coef = [&](int phase, int tap) -> double {
        int w = phase % L;
        if (w < 0) w += L;
        const long idx = (long)tap * L + w;
        return (idx >= 0 && idx < (long)proto.size() && idx < (long)L * tpp) ? proto[(size_t)idx] : 0.0;
    };
    st.step = (int64_t)std::round((1.0 / ratio) * (double)L * 65536.0);
    st.interp = (st.step & 0xFFFF) != 0;
    for (auto& b : st.bank) b.assign((size_t)L * tpp, 0.0);
    for (int p = 0; p < L; ++p)
        for (int t = 0; t < tpp; ++t) {
            const double f0 = coef(p, t), f1 = coef(p + 1, t), fm1 = coef(p - 1, t), f2 = coef(p + 2, t);
            const double a = f0;
            const double c = 0.5 * (f1 + fm1) - f0;
            const double d = (1.0 / 6.0) * (f2 - f1 + fm1 - f0 - 4.0 * c);
            const double b = f1 - f0 - d - c;
            const size_t o = (size_t)p * tpp + (size_t)(tpp - 1 - t);
            st.bank[0][o] = a;
            st.bank[1][o] = b;
            st.bank[2][o] = c;
            st.bank[3][o] = d;
        }
    return true;
}

}  // namespace

bool design_engine(double in_rate, double out_rate, int quality, Chain& chain, std::string& err) {
    if (!(in_rate > 0) || !(out_rate > 0)) {
        err = "sample rates must be positive";
        return false;
    }
    const double ratio = out_rate / in_rate;
    if (ratio < 1.0 / 256.0 || ratio > 256.0) {
        err = "resampling ratio out of valid range";
        return false;
    }
    EngineDesign e;
    e.in_rate = in_rate;
    e.out_rate = out_rate;
    e.ratio = ratio;
    e.quality = quality;
    e.first_stage = (int)chain.stages.size();
    const int eidx = (int)chain.engines.size();
    auto push = [&](StageDesign& s) {
        s.engine_index = eidx;
        chain.stages.push_back(std::move(s));
        e.n_stages++;
    };
    if (quality == EQ_QUICK) {  // resampler.go:76-81
        StageDesign s;
        s.kind = STAGE_CUBIC;
        s.quality = quality;
        s.ratio = ratio;
        s.taps = 4;
        e.has_cubic = true;
        push(s);
    } else if (ratio >= 1.0) {
        if (is_integer_ratio(ratio)) {  // :88-96
            const int f = (int)std::round(ratio);
            if (f > 1) {
                StageDesign s;
                if (!design_up(f, quality, s, err)) return false;
                e.has_pre = true;
                push(s);
            }  // f == 1: DFTStage(1) is a pass-through, no stage at all
        } else {  // :97-121
            StageDesign s1, s2;
            if (!design_up(2, quality, s1, err)) return false;
            const double inter = in_rate * 2.0;
            if (!design_poly(out_rate / inter, in_rate / out_rate, true, quality, s2, err)) return false;
            e.has_pre = e.has_poly = true;
            push(s1);
            push(s2);
        }
    } else {
        const double io = in_rate / out_rate;
        if (is_integer_ratio(io) && io >= 2.0) {  // :132-141
            StageDesign s;
            if (!design_decim((int)std::round(io), quality, s, err)) return false;
            e.has_decim = true;
            push(s);
        } else {  // :142-175
            StageDesign s1, s2;
            if (!design_up(2, quality, s1, err)) return false;
            const double inter = in_rate * 2.0;
            if (!design_poly(out_rate / inter, io, false, quality, s2, err)) return false;
            e.has_pre = e.has_poly = true;
            push(s1);
            push(s2);
        }
    }
    chain.engines.push_back(e);
    return true;
}

int preset_precision(int preset) {  // resample.go:217-267
    switch (preset) {
        case 0: return 8;
        case 1: case 2: return 16;
        case 3: return 24;
        case 4: return 32;
        default: return 0;
    }
}

int precision_to_engine_quality(int p) {  // stages.go:92-108
    if (p <= 8) return EQ_QUICK;
    if (p <= 16) return EQ_LOW;
    if (p <= 20) return EQ_HIGH;
    if (p <= 24) return EQ_24BIT;
    if (p <= 28) return EQ_VERYHIGH;
    return EQ_32BIT;
}

int preset_to_engine_quality(int preset) {  // convenience.go:189-200
    switch (preset) {
        case 0: case 1: return EQ_LOW;
        case 2: return EQ_MEDIUM;
        case 3: case 4: return EQ_HIGH;
        default: return EQ_MEDIUM;
    }
}

bool design_pipeline(double in_rate, double out_rate, int precision, Chain& chain, std::string& err) {
    const double ratio = out_rate / in_rate;
    chain.ratio = ratio;
    chain.precision = precision;
    if (!(ratio > 0)) {
        err = "invalid ratio";
        return false;
    }
    struct Spec { int type; double ratio; };
    std::vector<Spec> plan;
    if (precision <= 8) {  // pipeline.go:115-121
        plan.push_back({PLAN_CUBIC, ratio});
    } else {
        double rem = ratio;
        if (ratio < 1.0)
            while (rem < 0.5) {  // :127-138
                plan.push_back({PLAN_HALFBAND, 0.5});
                rem *= 2.0;
            }
        if (ratio > 1.0)
            while (rem > 2.0) {  // :141-152
                plan.push_back({PLAN_HALFBAND, 2.0});
                rem /= 2.0;
            }
        if (std::fabs(rem - 1.0) > 0.001) {  // :155-177, shouldUseFFT :320-334
            bool fft = precision >= 28;
            const double common[6] = {44100.0 / 48000.0, 48000.0 / 44100.0, 44100.0 / 88200.0,
                                      88200.0 / 44100.0, 48000.0 / 96000.0, 96000.0 / 48000.0};
            for (double c : common)
                if (std::fabs(rem - c) < 0.0001) fft = true;
            plan.push_back({fft ? PLAN_FFT : PLAN_POLYPHASE, rem});
        }
    }
    for (const Spec& sp : plan) {
        if (sp.type == PLAN_CUBIC) {  // newCubicStage(spec.Ratio): no 48 kHz round trip (stages.go:21-23)
            EngineDesign e;
            e.in_rate = 48000.0;
            e.out_rate = 48000.0 * sp.ratio;
            e.ratio = sp.ratio;
            e.quality = EQ_QUICK;
            e.first_stage = (int)chain.stages.size();
            e.n_stages = 1;
            e.has_cubic = true;
            e.plan_type = PLAN_CUBIC;
            StageDesign s;
            s.kind = STAGE_CUBIC;
            s.quality = EQ_QUICK;
            s.ratio = sp.ratio;
            s.taps = 4;
            s.engine_index = (int)chain.engines.size();
            chain.stages.push_back(std::move(s));
            chain.engines.push_back(e);
            continue;
        }
        // every other stage type is a whole engine.Resampler designed at a nominal 48 kHz (stages.go:54-71)
        const double ir = 48000.0;
        const double orate = ir * sp.ratio;
        if (!design_engine(ir, orate, precision_to_engine_quality(precision), chain, err)) return false;
        chain.engines.back().plan_type = sp.type;
    }
    return true;
}

}  // namespace gar
