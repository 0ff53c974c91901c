/* gar.h — C ABI of the B200-native go-audio-resampler engine ("gar" = go-audio-resampler).
 *
 * The reference (tphakala/go-audio-resampler, pure Go) has no FFI of its own; its
 * boundary is its Go API.  This header is what a cgo shim binds so that the Go
 * signatures stay the entry point (see INTEGRATION.md for the Go side).  Every
 * entry point cites the reference interface it replaces; paths are relative to
 * the reference repository root.
 *
 * Conventions
 *  - every function returns a gar_status (0 = OK) unless stated otherwise;
 *  - sample buffers are planar, one contiguous run per channel/stream;
 *  - the caller owns every host/device buffer for the duration of the call only
 *    (cgo rule: no pointer is retained after return); the handle owns device
 *    memory (filter banks, per-stream carry state, staging) and its CUDA streams;
 *  - a handle is not re-entrant (doc.go:199-206, constant.go:429-432); calls may
 *    arrive on any OS thread (goroutines migrate) — the library sets the device
 *    per call and keeps no thread-local state;
 *  - there is no CPU fallback: without a usable CUDA device gar_create fails
 *    with GAR_CUDA_ERROR.
 */
#ifndef GAR_H_
#define GAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gar_handle gar_handle;

/* Go error sentinels (resample.go:156-165) + transport errors. */
typedef enum gar_status {
    GAR_OK = 0,
    GAR_INVALID_CONFIG = 1,   /* ErrInvalidConfig   */
    GAR_BUFFER_TOO_SMALL = 2, /* ErrBufferTooSmall  */
    GAR_NOT_SUPPORTED = 3,    /* ErrNotSupported    */
    GAR_CUDA_ERROR = 4,
    GAR_INTERNAL = 5
} gar_status;

/* Which Go constructor family the handle mirrors (the two preset maps differ). */
typedef enum gar_path {
    GAR_PATH_PIPELINE = 0, /* New(*Config)                       resample.go:272-292, constant.go:42-86 */
    GAR_PATH_ENGINE = 1    /* NewEngine / NewEngineFloat32       convenience.go:125-132, 329-336        */
} gar_path;

typedef enum gar_dtype { GAR_F64 = 0, GAR_F32 = 1 } gar_dtype;

/* QualityPreset (resample.go:108-131). */
typedef enum gar_preset {
    GAR_QUALITY_QUICK = 0,
    GAR_QUALITY_LOW = 1,
    GAR_QUALITY_MEDIUM = 2,
    GAR_QUALITY_HIGH = 3,
    GAR_QUALITY_VERY_HIGH = 4,
    GAR_QUALITY_CUSTOM = 5
} gar_preset;

/* engine.Quality (internal/engine/filter_params.go:16-42), for engine-level callers
 * such as cmd/resample-wav which constructs engine.Resampler directly. */
typedef enum gar_engine_quality {
    GAR_EQ_FROM_PRESET = -1,
    GAR_EQ_QUICK = 0, GAR_EQ_LOW, GAR_EQ_MEDIUM, GAR_EQ_HIGH, GAR_EQ_VERY_HIGH,
    GAR_EQ_16BIT, GAR_EQ_20BIT, GAR_EQ_24BIT, GAR_EQ_28BIT, GAR_EQ_32BIT
} gar_engine_quality;

/* Mirrors Config + QualitySpec (resample.go:46-102).  Fields the reference accepts but
 * never lets reach the filters (phase response, passband/stopband, flags, EnableSIMD,
 * MaxInputSize — SURVEY.md §5) are validated exactly like Config.Validate and otherwise ignored. */
typedef struct gar_config {
    double input_rate;            /* Config.InputRate  */
    double output_rate;           /* Config.OutputRate */
    int32_t channels;             /* Config.Channels (1..256); path ENGINE: must be 1 */
    int32_t path;                 /* gar_path */
    int32_t preset;               /* gar_preset */
    int32_t custom_precision;     /* QualitySpec.Precision when preset == CUSTOM (8..33) */
    double custom_phase_response; /* validated 0..100 */
    double custom_passband_end;   /* validated (0,1) */
    double custom_stopband_begin; /* validated (passband_end,1] */
    int32_t dtype;                /* gar_dtype; PIPELINE path computes in f64 (constant.go:161-199) */
    int32_t engine_quality;       /* gar_engine_quality; -1 = derive from preset */
    int32_t n_streams;            /* batch extension: number of independent replicas of the whole
                                     channel set that advance in lock step (0 or 1 = none).  Total
                                     device streams = channels * n_streams. */
    int32_t device;               /* CUDA device ordinal; -1 = geometry-only handle (plans, counts, banks and the
                                     integer state machine on the host; every sample-processing call fails
                                     with GAR_CUDA_ERROR — used to test host logic where no GPU exists) */
    int32_t max_input_size;       /* Config.MaxInputSize (pre-sizes staging only) */
    uint32_t flags;               /* QualityFlags | (EnableParallel<<16) | (EnableSIMD<<17): no numeric effect */
} gar_config;

/* One primitive stage of the flattened chain, for bit-exact diffing of integer geometry
 * against the Go engine (dft_stage.go:22-47,370-388; polyphase_stage.go:25-58; cubic.go:15-21). */
typedef enum gar_stage_kind { GAR_STAGE_UP = 0, GAR_STAGE_DECIM = 1, GAR_STAGE_POLY = 2, GAR_STAGE_CUBIC = 3 } gar_stage_kind;
typedef struct gar_stage_desc {
    int32_t kind;          /* gar_stage_kind */
    int32_t engine_index;  /* which engine.Resampler of the path-A pipeline this stage belongs to */
    int32_t factor;        /* UP: L ; DECIM: M ; POLY: numPhases */
    int32_t taps;          /* UP/POLY: tapsPerPhase ; DECIM: numTaps */
    int32_t proto_taps;    /* prototype length */
    int32_t engine_quality;
    int64_t step;          /* POLY: fixed-point step (16 fractional bits) */
    int64_t at;            /* POLY: current accumulator */
    int64_t hist_len;      /* carried history length (samples) */
    int64_t decim_phase;   /* DECIM */
    double ratio;          /* stage ratio */
} gar_stage_desc;

/* Info (resample.go:295-316) */
typedef struct gar_info {
    char algorithm[32];
    int32_t filter_length;
    int32_t phases;
    int32_t latency;
    int64_t memory_usage;
    int32_t simd_enabled;
    char simd_type[64]; /* "CUDA sm_100a (NVIDIA B200)" instead of cpu.Info() (stage_adapter.go:122-124) */
} gar_info;

/* ---- lifecycle -------------------------------------------------------------------------- */

/* New(*Config) (resample.go:272-292) / NewEngine, NewEngineFloat32 (convenience.go:125,329).
 * Validates like Config.Validate (resample.go:168-214) and engine.NewResampler (resampler.go:51-70),
 * designs the Kaiser banks on the host once per config (shared by all channels), uploads them. */
int32_t gar_create(const gar_config* cfg, gar_handle** out);
/* Multi-device handle — the GPU analogue of the channel fan-out inside one call (constant.go:223-241, ProcessMulti's
 * goroutines): ONE handle shards its rows over `n_devices` CUDA devices (whole streams when n_streams > 1, else channels),
 * each shard with its own device state, streams, pinned staging and a worker thread bound to the device's NUMA node; calls
 * fan out to the shards and join on the host. Rows never migrate and nothing is exchanged between devices. cfg->device
 * is ignored. Supported: every host-buffer call (gar_process_* / gar_flush_* / *_multi / *_batch), Reset, queries.
 * Not supported (GAR_NOT_SUPPORTED): the *_dev and *_interleaved calls (their buffers belong to one device). */
int32_t gar_create_multi(const gar_config* cfg, const int32_t* devices, int32_t n_devices, gar_handle** out);
/* Number of shards behind a handle (1 for gar_create handles) and the device / row block of shard k. A device list of
 * all -1 makes geometry-only shards (host logic only, like device = -1). */
int32_t gar_num_devices(const gar_handle* h);
int32_t gar_shard_info(const gar_handle* h, int32_t shard, int32_t* device, int32_t* row0, int32_t* rows, int32_t* numa_node);
void gar_destroy(gar_handle* h);
/* Last error text of this handle (or of the last failed gar_create when h == NULL). */
const char* gar_last_error(const gar_handle* h);
const char* gar_status_string(int32_t status);

/* ---- geometry / info --------------------------------------------------------------------- */

/* EstimateOutput (constant.go:117-119, convenience.go:164-166): int(n*ratio)+64. */
int64_t gar_estimate_output(const gar_handle* h, int64_t n_in);
/* Exact number of samples the next Process(n_in) on `stream` would return (lets the Go shim
 * allocate the owned result slice of Process, constant.go:88-96, without over-allocating). */
int64_t gar_next_output_count(const gar_handle* h, int32_t stream, int64_t n_in);
/* Exact number of samples Flush would return now (constant.go:349-386, resampler.go:275-322). */
int64_t gar_next_flush_count(const gar_handle* h, int32_t stream);
double gar_get_ratio(const gar_handle* h);                 /* GetRatio   constant.go:444-447 */
int32_t gar_get_latency(const gar_handle* h);              /* GetLatency constant.go:407-423 */
int32_t gar_get_info(const gar_handle* h, gar_info* out);  /* GetInfo    constant.go:452-485 */
/* GetStatistics (resampler.go:348-353) of engine `engine_index` of `stream`. */
int32_t gar_get_stats(const gar_handle* h, int32_t stream, int32_t engine_index, int64_t* samples_in, int64_t* samples_out);
int32_t gar_num_stages(const gar_handle* h);
int32_t gar_num_engines(const gar_handle* h);
int32_t gar_describe_stage(const gar_handle* h, int32_t stream, int32_t stage, gar_stage_desc* out);
/* pipeline.StageType sequence of path A (pipeline.go:58-73): 0 cubic, 1 half-band, 2 polyphase, 3 fft. */
int32_t gar_plan_stage_type(const gar_handle* h, int32_t engine_index);
/* Coefficient bank as stored (reversed taps), widened to double.  which: 0 = UP bank [factor][taps] or
 * DECIM taps; POLY: 0..3 = a,b,c,d banks [phases][taps]. Returns count, or -needed. */
int64_t gar_get_bank(const gar_handle* h, int32_t stage, int32_t which, double* out, int64_t cap);
/* Replace a bank with host-designed coefficients (e.g. from the Go internal/filter design,
 * filter/kaiser.go:159-233), same layout as gar_get_bank. */
int32_t gar_upload_bank(gar_handle* h, int32_t stage, int32_t which, const double* coef, int64_t n);

/* ---- processing: host buffers (the Go-facing calls) --------------------------------------- */

/* Process / ProcessInto (constant.go:88-112; convenience.go:134-160) on one channel.
 * GAR_BUFFER_TOO_SMALL if out_cap < gar_estimate_output(n_in), decided before any state
 * change (processinto_test.go:176-224). n_in == 0 => *n_out = 0. */
int32_t gar_process_f64(gar_handle* h, int32_t channel, const double* in, int64_t n_in, double* out, int64_t out_cap, int64_t* n_out);
/* ProcessFloat32Into (constant.go:161-199: f64 pipeline between casts) on PIPELINE handles;
 * SimpleResamplerFloat32.ProcessInto (convenience.go:349-366: f32 end to end) on ENGINE/F32 handles. */
int32_t gar_process_f32(gar_handle* h, int32_t channel, const float* in, int64_t n_in, float* out, int64_t out_cap, int64_t* n_out);
/* ProcessMulti (constant.go:204-252): all channels, one device pass. in/out: arrays of `channels`
 * planar pointers; n_in[c] may differ per channel; out_cap per channel; n_out[c] returned. */
int32_t gar_process_multi_f64(gar_handle* h, const double* const* in, const int64_t* n_in, double* const* out, int64_t out_cap, int64_t* n_out);
/* Flush (constant.go:349-354; resampler.go:275-322) of one channel; FlushMulti (constant.go:390-404). */
int32_t gar_flush_f64(gar_handle* h, int32_t channel, double* out, int64_t out_cap, int64_t* n_out);
int32_t gar_flush_f32(gar_handle* h, int32_t channel, float* out, int64_t out_cap, int64_t* n_out);
int32_t gar_flush_multi_f64(gar_handle* h, double* const* out, int64_t out_cap, int64_t* n_out);
/* Advance only the integer streaming state of `stream` as Process(n_in) (flush == 0) or Flush() would,
 * without moving samples; *n_out = samples that call would have returned. For geometry-only handles. */
int32_t gar_advance_geometry(gar_handle* h, int32_t stream, int64_t n_in, int32_t flush, int64_t* n_out);
/* Reset (constant.go:425-441, resampler.go:325-340). */
int32_t gar_reset(gar_handle* h);

/* ---- processing: batched independent streams (extension) ------------------- */

/* All channels*n_streams rows advance by n_in samples. `in`/`out` are planar 2-D arrays in HOST memory,
 * row r at in + r*in_stride (elements of the handle dtype; PIPELINE handles: dtype of the I/O, f64 or
 * f32 per `io_dtype`). Rows are cut into slices that are copied, resampled and copied back on
 * alternating CUDA streams so PCIe and the SMs overlap. *n_out = samples per row. */
int32_t gar_process_batch(gar_handle* h, int32_t io_dtype, const void* in, int64_t in_stride, int64_t n_in, void* out, int64_t out_stride, int64_t out_cap, int64_t* n_out);
int32_t gar_flush_batch(gar_handle* h, int32_t io_dtype, void* out, int64_t out_stride, int64_t out_cap, int64_t* n_out);
/* Same with DEVICE pointers, enqueued on `cuda_stream` (a cudaStream_t; NULL = the handle's own stream,
 * pass cudaStreamLegacy / cudaStreamPerThread to name the default streams); returns after enqueue.
 * Use when inputs are already resident in HBM. */
int32_t gar_process_batch_dev(gar_handle* h, int32_t io_dtype, const void* d_in, int64_t in_stride, int64_t n_in, void* d_out, int64_t out_stride, int64_t out_cap, int64_t* n_out, void* cuda_stream);
int32_t gar_flush_batch_dev(gar_handle* h, int32_t io_dtype, void* d_out, int64_t out_stride, int64_t out_cap, int64_t* n_out, void* cuda_stream);

/* ---- interleaved / integer-PCM boundary -------------------------------------- */

/* Sample container of an interleaved buffer (frame-major: all channels of frame i are adjacent). */
typedef enum gar_sample_format {
    GAR_FMT_F64 = 0, GAR_FMT_F32 = 1, /* InterleaveToStereo/DeinterleaveFromStereo[Float32] (convenience.go:261-282,463-486) */
    GAR_FMT_I16 = 2,                  /* packed int16 PCM */
    GAR_FMT_I32 = 3,                  /* 16/24/32-bit PCM in int32 containers */
    GAR_FMT_I64 = 4                   /* Go `[]int` (go-audio IntBuffer.Data) on 64-bit platforms */
} gar_sample_format;

/* One block of the resample-wav loop (cmd/resample-wav/helpers.go:77-334) in one call: deinterleave + normalise
 * (`F(float64(v) * (1/maxVal))`, main.go:444-470), resample every channel (rows = channels * n_streams, lock step),
 * clamp to [-1,1] + `int(sample * maxVal)` + interleave (main.go:474-520) — conversions run on the device around
 * the kernels, so the host copies one contiguous block each way. bit_depth 16/24/32 selects maxVal
 * 32767 / 8388607 / 2147483647 (anything else: 32767, like the reference's default); ignored for float formats. */
int32_t gar_process_interleaved(gar_handle* h, int32_t fmt, int32_t bit_depth, const void* in, int64_t n_frames,
                                void* out, int64_t out_cap_frames, int64_t* n_frames_out);
/* flushAndPadChannels (cmd/resample-wav/helpers.go:293-334). */
int32_t gar_flush_interleaved(gar_handle* h, int32_t fmt, int32_t bit_depth, void* out, int64_t out_cap_frames,
                              int64_t* n_frames_out);

/* ---- utilities --------------------------------------------------------------------------- */

/* Pinned host memory for callers that want full PCIe rate (Go slices are pageable). */
void* gar_host_alloc(size_t bytes);
void gar_host_free(void* p);
/* Write-combined pinned memory (cudaHostAllocWriteCombined): for INPUT buffers the host only writes — not snooped during the
 * DMA read, which helps on platforms whose host-to-device path is the ceiling; never read it back on the CPU (slow). */
void* gar_host_alloc_wc(size_t bytes);
/* Pinned planar buffer of rows(h) x row_bytes for the batch calls whose pages are placed shard by shard on the NUMA node of
 * the device that will copy them (first touch by the shard's bound worker thread, then cudaHostRegister). Free with
 * gar_host_free. On single-socket machines it is simply a pinned buffer. */
void* gar_host_alloc_rows(gar_handle* h, size_t row_bytes);
/* Restrict the calling thread to the cores of `device`'s NUMA node (sysfs numa_node of its PCI function). One-process-
 * per-GPU callers (torchrun ranks) call it once before allocating pinned memory. Returns the node, -1 if nothing changed. */
int32_t gar_bind_thread_to_device(int32_t device);
int32_t gar_device_numa_node(int32_t device);
int32_t gar_device_count(void);
/* Enable (default) / disable the fused x2 -> polyphase launch (K4); results agree to rounding, used for A/B tests. */
int32_t gar_set_fusion(gar_handle* h, int32_t enabled);
/* Time slicing of long multi-stage calls (batch / device entry points and every Process call): a call whose inter-stage
 * buffers would exceed `bytes` runs as a sequence of shorter calls with identical results, which bounds the device memory
 * of the inter-stage buffers (default 2 GiB) or, with an L2-sized budget (e.g. 40 MiB), keeps the intermediate-rate
 * streams out of HBM at the price of more launches. 0 disables. */
int32_t gar_set_slice_budget(gar_handle* h, int64_t bytes);
/* Process-wide A/B switch (tests, profiling): 0 routes every polyphase stage through the one-thread-per-output
 * kernels instead of the register-tiled ones (K4r / K3r / K3i). Results are bit-identical in float64. Default 1. */
void gar_set_tiled_polyphase(int32_t enabled);
/* Process-wide A/B switch: 0 routes the float64 integer-factor FIR stages of >= 8-row batches through the vector-FMA
 * kernels instead of the FP64 tensor-core (DMMA) kernels. The two differ in the last bits (taps grouped in fours). Default 1. */
void gar_set_tensor_fir(int32_t enabled);
/* Process-wide policy for the persistent chain kernel (K5): the x2 stage and the polyphase stage of a large float64 batch
 * (>= 32 lock-step rows) as ONE launch per Process call, the intermediate-rate samples in an L2-resident ring instead of a
 * full-size device buffer (resampler.go:182-227). Bit-identical to the two stand-alone tensor-core launches; measured on
 * B200 it moves 2.2 GB instead of 5.5 GB through HBM for 256 rows x 10 s of 44.1k->48k but takes 4.0 instead of 3.1 ms.
 * 0: never; 1: every eligible call; 2 (default): eligible calls whose intermediate buffer would exceed the inter-stage
 * memory budget (gar_set_slice_budget) by so much that time slices would be shorter than 64 K samples — they run as one
 * launch instead of a sequence of short slices. */
void gar_set_chain_kernel(int32_t mode);
/* Test hook (no device needed): K5 assigns the 64-output tiles of the polyphase stage to chunks of `chunk_len` intermediate
 * samples — a tile belongs to the first chunk that contains the END of everything its staging reads. Returns the number of
 * tiles assigned to chunks 0 .. chunk. tests/test_chain_geometry.py checks the invariants the kernel's dependency counters rely on. */
int32_t gar_debug_chain_tile_hi(int32_t chunk_len, int32_t kp, int32_t n_chunks, int32_t n_tiles, int32_t chunk, int32_t hist_len,
                                int64_t L, int64_t at0, int64_t step, int32_t n_out);
/* Number of this library's kernels launched through the handle since creation / last reset of the counter. */
int64_t gar_kernel_launches(const gar_handle* h, int32_t reset);
/* Name of the dominant kernel variant chosen for stage `stage` (for bench/ncu filters). */
const char* gar_stage_kernel_name(const gar_handle* h, int32_t stage);
/* Comma-separated names of the distinct kernel variants the handle has launched so far (fused launches appear
 * under their own names), written to buf (NUL-terminated, truncated to cap). Returns the untruncated length. */
int32_t gar_kernels_used(const gar_handle* h, char* buf, int32_t cap);
/* Plain cudaMemcpyAsync (kind 1 = host to device, 2 = device to host, else default) on `cuda_stream` of the current
 * device: lets ctypes / cgo callers measure the box's copy ceiling next to the batch calls (bench.py `copy_ceiling`). */
int32_t gar_memcpy_async(void* dst, const void* src, size_t bytes, int32_t kind, void* cuda_stream);
/* Dependent-FMA micro-benchmark on `device`: achieved TFLOP/s (roofline denominators). dtype: GAR_F64 / GAR_F32 vector FMA,
 * 2 = packed fma.rn.f32x2 (FFMA2), 3 = FP64 tensor cores (mma.sync.m8n8k4.f64, DMMA.8x8x4). */
int32_t gar_measure_fma_peak(int32_t device, int32_t dtype, double* tflops);
const char* gar_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GAR_H_ */
