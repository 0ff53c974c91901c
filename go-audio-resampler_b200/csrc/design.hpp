// design.hpp — host-side filter design and chain planning for the B200 engine.
//
// By the north star the Kaiser design stays on the host: it runs once per
// configuration, produces every coefficient bank the device kernels need
// (already in the kernels' layout) and the flat list of primitive stages.
// Unlike the reference, which designs an identical filter set per channel
// (constant.go:57-70), one bank set is shared by all channels/streams.
//
// The arithmetic follows the reference's formulas and operation order
// (internal/mathutil/bessel.go, internal/filter/kaiser.go,
// internal/engine/filter_params.go, dft_stage.go:50-146,401-475,
// polyphase_stage.go:69-170) because coefficients feed a 1e-12 parity bar;
// compiled with -ffp-contract=off (Go/amd64 never fuses scalar a*b+c).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gar {

enum EngineQuality : int {  // engine.Quality, filter_params.go:16-42
    EQ_QUICK = 0, EQ_LOW, EQ_MEDIUM, EQ_HIGH, EQ_VERYHIGH, EQ_16BIT, EQ_20BIT, EQ_24BIT, EQ_28BIT, EQ_32BIT
};

enum StageKind : int { STAGE_UP = 0, STAGE_DECIM = 1, STAGE_POLY = 2, STAGE_CUBIC = 3 };
enum PlanType : int { PLAN_CUBIC = 0, PLAN_HALFBAND = 1, PLAN_POLYPHASE = 2, PLAN_FFT = 3 };  // pipeline.go:58-73

// One primitive FIR stage of the flattened chain, with its host-designed bank.
struct StageDesign {
    int kind = STAGE_UP;
    int engine_index = 0;   // index of the engine.Resampler (path-A pipeline stage) it belongs to
    int quality = EQ_HIGH;
    int factor = 1;         // UP: L; DECIM: M; POLY: numPhases
    int taps = 0;           // UP/POLY: taps per phase; DECIM: total taps
    int proto_taps = 0;
    int64_t step = 0;       // POLY
    double ratio = 1.0;     // CUBIC: ratio; others informational
    // Banks, double precision, reversed-tap order exactly as the reference stores them:
    //  UP:    bank[0] = [factor][taps]
    //  DECIM: bank[0] = [taps]
    //  POLY:  bank[0..3] = a,b,c,d, each [factor][taps]
    std::vector<double> bank[4];
    bool interp = false;    // POLY: (step & 0xFFFF) != 0 -> b,c,d contribute
};

// One engine.Resampler worth of stages (resampler.go:26-44).
struct EngineDesign {
    double in_rate = 0, out_rate = 0, ratio = 1;
    int quality = EQ_HIGH;
    int first_stage = 0;   // index into Chain::stages
    int n_stages = 0;      // 0 (ratio==1 passthrough DFTStage(1)), 1 or 2
    bool has_cubic = false, has_pre = false, has_decim = false, has_poly = false;
    int plan_type = -1;    // PlanType when built by the path-A planner
};

struct Chain {
    std::vector<StageDesign> stages;
    std::vector<EngineDesign> engines;
    double ratio = 1.0;    // overall output/input
    int precision = 0;     // path A: QualitySpec.Precision after preset expansion
};

// --- scalar design helpers (exposed for tests through the C ABI only indirectly) ---
double bessel_i0(double x);
double kaiser_beta(double att);
int estimate_filter_length(double att, double tbw);
bool design_lowpass(int taps, double fc, double att, double gain, std::vector<double>& h);
double quality_attenuation(int q);
double quality_passband_end(int q);
bool is_integer_ratio(double r);

// engine.NewResampler (resampler.go:51-179): appends 0..2 stages for one engine.
// Returns false with `err` set on invalid parameters.
bool design_engine(double in_rate, double out_rate, int quality, Chain& chain, std::string& err);

// path A: New(Config) -> BuildPipeline (pipeline.go:104-183) -> createStage (pipeline_builder.go:76-100).
bool design_pipeline(double in_rate, double out_rate, int precision, Chain& chain, std::string& err);

// Preset maps.
int preset_precision(int preset);             // resample.go:217-267
int precision_to_engine_quality(int prec);    // stages.go:92-108
int preset_to_engine_quality(int preset);     // convenience.go:189-200

}  // namespace gar
