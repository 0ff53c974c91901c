// Package b200 binds the B200 resampling engine (libgar_b200.so, include/gar.h) with cgo and exposes
// the reference's own Go API on top of it:
//
//	New(*Config) (Resampler, error)             resample.go:272
//	NewEngine / NewEngineFloat32                convenience.go:125, 329
//	ResampleMono / ResampleMonoFloat32          convenience.go:204, 407
//
// so `resampler.New` can be pointed at the GPU engine without touching callers.
// NOTE: this file cannot be compiled in the build image (no Go toolchain); it is the shim a
// maintainer adds, see INTEGRATION.md. Build (go.mod beside this file; `make -C go-audio-resampler_b200/go check` does
// vet + build + the parity dump on a machine that has Go and CUDA):
//   CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/go-audio-resampler_b200/_build -lgar_b200" go build ./...
package b200

/*
#cgo LDFLAGS: -lgar_b200
#include <stdlib.h>
#include "gar.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// Error sentinels of the reference (resample.go:156-165).
var (
	ErrInvalidConfig  = errors.New("invalid resampler configuration")
	ErrBufferTooSmall = errors.New("output buffer too small")
	ErrNotSupported   = errors.New("operation not supported")
)

// QualityPreset mirrors resample.go:108-131.
type QualityPreset int

const (
	QualityQuick QualityPreset = iota
	QualityLow
	QualityMedium
	QualityHigh
	QualityVeryHigh
	QualityCustom
)

// QualitySpec and Config mirror resample.go:46-102.
type QualitySpec struct {
	Preset        QualityPreset
	Precision     int
	PhaseResponse float64
	PassbandEnd   float64
	StopbandBegin float64
	Flags         uint32
}

type Config struct {
	InputRate, OutputRate float64
	Channels              int
	Quality               QualitySpec
	MaxInputSize          int
	EnableSIMD            bool
	EnableParallel        bool
	Device                int // extension: CUDA ordinal
}

func statusErr(st C.int32_t, h *C.gar_handle) error {
	if st == C.GAR_OK {
		return nil
	}
	msg := C.GoString(C.gar_last_error(h))
	switch st {
	case C.GAR_INVALID_CONFIG:
		return fmt.Errorf("%w: %s", ErrInvalidConfig, msg)
	case C.GAR_BUFFER_TOO_SMALL:
		return ErrBufferTooSmall
	case C.GAR_NOT_SUPPORTED:
		return fmt.Errorf("%w: %s", ErrNotSupported, msg)
	default:
		return fmt.Errorf("b200: %s (status %d)", msg, int(st))
	}
}

type handle struct{ h *C.gar_handle }

func newHandle(cfg *C.gar_config) (*handle, error) {
	var h *C.gar_handle
	if st := C.gar_create(cfg, &h); st != C.GAR_OK {
		return nil, statusErr(st, nil)
	}
	r := &handle{h: h}
	runtime.SetFinalizer(r, func(r *handle) { C.gar_destroy(r.h) })
	return r, nil
}

// Resampler implements the reference's Resampler interface (resample.go:14-43) plus the optional
// ProcessInto / ProcessFloat32Into / EstimateOutput / FlushMulti methods (constant.go:103-199,390-404).
type Resampler struct {
	*handle
	channels int
}

// New mirrors resample.go:272-292. Validation (Config.Validate) happens inside gar_create.
func New(config *Config) (*Resampler, error) {
	if config == nil {
		return nil, fmt.Errorf("%w: config is nil", ErrInvalidConfig)
	}
	flags := C.uint32_t(config.Quality.Flags)
	if config.EnableParallel {
		flags |= 1 << 16
	}
	cfg := C.gar_config{
		input_rate: C.double(config.InputRate), output_rate: C.double(config.OutputRate),
		channels: C.int32_t(config.Channels), path: C.GAR_PATH_PIPELINE,
		preset: C.int32_t(config.Quality.Preset), custom_precision: C.int32_t(config.Quality.Precision),
		custom_phase_response: C.double(config.Quality.PhaseResponse),
		custom_passband_end:   C.double(config.Quality.PassbandEnd),
		custom_stopband_begin: C.double(config.Quality.StopbandBegin),
		dtype:                 C.GAR_F64, engine_quality: -1, device: C.int32_t(config.Device),
		max_input_size: C.int32_t(config.MaxInputSize), flags: flags,
	}
	h, err := newHandle(&cfg)
	if err != nil {
		return nil, err
	}
	return &Resampler{handle: h, channels: config.Channels}, nil
}

// EstimateOutput: constant.go:117-119.
// Every method ends with runtime.KeepAlive(r): r.h is read before the cgo call, so without it the finalizer
// (gar_destroy) could run while the C side still uses the handle.
func (r *handle) EstimateOutput(n int) int {
	v := int(C.gar_estimate_output(r.h, C.int64_t(n)))
	runtime.KeepAlive(r)
	return v
}
func (r *handle) GetRatio() float64 {
	v := float64(C.gar_get_ratio(r.h))
	runtime.KeepAlive(r)
	return v
}
func (r *handle) GetLatency() int {
	v := int(C.gar_get_latency(r.h))
	runtime.KeepAlive(r)
	return v
}
func (r *handle) Reset() {
	C.gar_reset(r.h)
	runtime.KeepAlive(r)
}

// nextOutputCount / nextFlushCount: exact sample counts of the next call (pure integer state machine).
func (r *handle) nextOutputCount(ch, n int) int {
	v := int(C.gar_next_output_count(r.h, C.int32_t(ch), C.int64_t(n)))
	runtime.KeepAlive(r)
	return v
}
func (r *handle) nextFlushCount(ch int) int {
	v := int(C.gar_next_flush_count(r.h, C.int32_t(ch)))
	runtime.KeepAlive(r)
	return v
}

// Process returns an owned, exactly-sized slice (constant.go:88-96). The C side never retains `input`.
func (r *Resampler) Process(input []float64) ([]float64, error) {
	if len(input) == 0 {
		return []float64{}, nil
	}
	capN := max(r.nextOutputCount(0, len(input)), r.EstimateOutput(len(input)))
	out := make([]float64, capN)
	n, err := r.ProcessInto(input, out)
	if err != nil {
		return nil, err
	}
	return out[:n:n], nil
}

// ProcessInto: constant.go:103-112 — ErrBufferTooSmall before any state is advanced.
func (r *Resampler) ProcessInto(input, output []float64) (int, error) {
	var n C.int64_t
	var ip, op *C.double
	if len(input) > 0 {
		ip = (*C.double)(unsafe.Pointer(&input[0]))
	}
	if len(output) > 0 {
		op = (*C.double)(unsafe.Pointer(&output[0]))
	}
	st := C.gar_process_f64(r.h, 0, ip, C.int64_t(len(input)), op, C.int64_t(len(output)), &n)
	err := statusErr(st, r.h)
	runtime.KeepAlive(input)
	runtime.KeepAlive(output)
	runtime.KeepAlive(r)
	return int(n), err
}

// ProcessFloat32Into: constant.go:161-199 (float64 pipeline between the casts, done on the device).
func (r *Resampler) ProcessFloat32Into(input, output []float32) (int, error) {
	var n C.int64_t
	var ip, op *C.float
	if len(input) > 0 {
		ip = (*C.float)(unsafe.Pointer(&input[0]))
	}
	if len(output) > 0 {
		op = (*C.float)(unsafe.Pointer(&output[0]))
	}
	st := C.gar_process_f32(r.h, 0, ip, C.int64_t(len(input)), op, C.int64_t(len(output)), &n)
	err := statusErr(st, r.h)
	runtime.KeepAlive(input)
	runtime.KeepAlive(output)
	runtime.KeepAlive(r)
	return int(n), err
}

// ProcessFloat32: constant.go:128-147.
func (r *Resampler) ProcessFloat32(input []float32) ([]float32, error) {
	out := make([]float32, r.EstimateOutput(len(input)))
	n, err := r.ProcessFloat32Into(input, out)
	if err != nil {
		return nil, err
	}
	return out[:n:n], nil
}

// ProcessMulti: constant.go:204-252 — all channels in one device pass (EnableParallel has no meaning here).
func (r *Resampler) ProcessMulti(input [][]float64) ([][]float64, error) {
	if len(input) != r.channels {
		return nil, fmt.Errorf("expected %d channels, got %d", r.channels, len(input))
	}
	c := r.channels
	// pointer arrays live in C memory so no Go pointer to Go pointers crosses the boundary
	ins := (*[1 << 20]*C.double)(C.malloc(C.size_t(c) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	outs := (*[1 << 20]*C.double)(C.malloc(C.size_t(c) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(ins))
	defer C.free(unsafe.Pointer(outs))
	nin := make([]C.int64_t, c)
	nout := make([]C.int64_t, c)
	capN := 1
	for ch := range input {
		nin[ch] = C.int64_t(len(input[ch]))
		capN = max(capN, r.nextOutputCount(ch, len(input[ch])))
	}
	output := make([][]float64, c)
	var pin runtime.Pinner
	defer pin.Unpin()
	for ch := range input {
		output[ch] = make([]float64, capN)
		pin.Pin(&output[ch][0])
		outs[ch] = (*C.double)(unsafe.Pointer(&output[ch][0]))
		if len(input[ch]) > 0 {
			pin.Pin(&input[ch][0])
			ins[ch] = (*C.double)(unsafe.Pointer(&input[ch][0]))
		}
	}
	st := C.gar_process_multi_f64(r.h, (**C.double)(unsafe.Pointer(ins)), &nin[0],
		(**C.double)(unsafe.Pointer(outs)), C.int64_t(capN), &nout[0])
	err := statusErr(st, r.h)
	runtime.KeepAlive(r)
	if err != nil {
		return nil, err
	}
	for ch := range output {
		output[ch] = output[ch][:nout[ch]:nout[ch]]
	}
	return output, nil
}

// Flush drains channel 0 only (constant.go:349-354); FlushMulti drains every channel (:390-404).
func (r *Resampler) Flush() ([]float64, error) {
	out := make([]float64, max(r.nextFlushCount(0), 1))
	var got C.int64_t
	st := C.gar_flush_f64(r.h, 0, (*C.double)(unsafe.Pointer(&out[0])), C.int64_t(len(out)), &got)
	err := statusErr(st, r.h)
	runtime.KeepAlive(r)
	return out[:got:got], err
}

func (r *Resampler) FlushMulti() ([][]float64, error) {
	out := make([][]float64, r.channels)
	for ch := range out {
		// per-channel calls keep the shim simple; gar_flush_multi_f64 does all channels in one pass
		buf := make([]float64, max(r.nextFlushCount(ch), 1))
		var got C.int64_t
		st := C.gar_flush_f64(r.h, C.int32_t(ch), (*C.double)(unsafe.Pointer(&buf[0])), C.int64_t(len(buf)), &got)
		err := statusErr(st, r.h)
		runtime.KeepAlive(r)
		if err != nil {
			return nil, err
		}
		out[ch] = buf[:got:got]
	}
	return out, nil
}

// SimpleResampler / SimpleResamplerFloat32: convenience.go:118-186, 315-395.
type SimpleResampler struct{ *handle }
type SimpleResamplerFloat32 struct{ *handle }

func newEngine(in, out float64, q QualityPreset, dtype C.int32_t) (*handle, error) {
	cfg := C.gar_config{input_rate: C.double(in), output_rate: C.double(out), channels: 1,
		path: C.GAR_PATH_ENGINE, preset: C.int32_t(q), dtype: dtype, engine_quality: -1}
	return newHandle(&cfg)
}

func NewEngine(in, out float64, q QualityPreset) (*SimpleResampler, error) {
	h, err := newEngine(in, out, q, C.GAR_F64)
	if err != nil {
		return nil, err
	}
	return &SimpleResampler{h}, nil
}

func NewEngineFloat32(in, out float64, q QualityPreset) (*SimpleResamplerFloat32, error) {
	h, err := newEngine(in, out, q, C.GAR_F32)
	if err != nil {
		return nil, err
	}
	return &SimpleResamplerFloat32{h}, nil
}

func (r *SimpleResampler) ProcessInto(input, output []float64) (int, error) {
	return (&Resampler{handle: r.handle, channels: 1}).ProcessInto(input, output)
}

func (r *SimpleResampler) Process(input []float64) ([]float64, error) {
	return (&Resampler{handle: r.handle, channels: 1}).Process(input)
}

func (r *SimpleResampler) Flush() ([]float64, error) {
	return (&Resampler{handle: r.handle, channels: 1}).Flush()
}

func (r *SimpleResamplerFloat32) ProcessInto(input, output []float32) (int, error) {
	return (&Resampler{handle: r.handle, channels: 1}).ProcessFloat32Into(input, output)
}

func (r *SimpleResamplerFloat32) Process(input []float32) ([]float32, error) {
	return (&Resampler{handle: r.handle, channels: 1}).ProcessFloat32(input)
}

func (r *SimpleResamplerFloat32) Flush() ([]float32, error) {
	out := make([]float32, max(r.nextFlushCount(0), 1))
	var got C.int64_t
	st := C.gar_flush_f32(r.h, 0, (*C.float)(unsafe.Pointer(&out[0])), C.int64_t(len(out)), &got)
	err := statusErr(st, r.h)
	runtime.KeepAlive(r)
	return out[:got:got], err
}

// GetStatistics: resampler.go:348-353.
func (r *handle) GetStatistics() map[string]int64 {
	var in, out C.int64_t
	C.gar_get_stats(r.h, 0, 0, &in, &out)
	runtime.KeepAlive(r)
	return map[string]int64{"samplesIn": int64(in), "samplesOut": int64(out)}
}

// ResampleMono: convenience.go:204-229.
func ResampleMono(input []float64, inRate, outRate float64, q QualityPreset) ([]float64, error) {
	r, err := NewEngine(inRate, outRate, q)
	if err != nil {
		return nil, err
	}
	out, err := r.Process(input)
	if err != nil {
		return nil, err
	}
	fl, err := r.Flush()
	if err != nil {
		return nil, err
	}
	return append(out, fl...), nil
}

// ResampleMonoFloat32: convenience.go:407-429.
func ResampleMonoFloat32(input []float32, inRate, outRate float64, q QualityPreset) ([]float32, error) {
	r, err := NewEngineFloat32(inRate, outRate, q)
	if err != nil {
		return nil, err
	}
	out, err := r.Process(input)
	if err != nil {
		return nil, err
	}
	fl, err := r.Flush()
	if err != nil {
		return nil, err
	}
	return append(out, fl...), nil
}

// Batch is the extension used for BASELINE config 4: n independent mono float32 streams in lock step.
type Batch struct {
	*handle
	Streams int
}

func NewBatchFloat32(in, out float64, q QualityPreset, streams, device int) (*Batch, error) {
	cfg := C.gar_config{input_rate: C.double(in), output_rate: C.double(out), channels: 1,
		path: C.GAR_PATH_ENGINE, preset: C.int32_t(q), dtype: C.GAR_F32, engine_quality: -1,
		n_streams: C.int32_t(streams), device: C.int32_t(device)}
	h, err := newHandle(&cfg)
	if err != nil {
		return nil, err
	}
	return &Batch{h, streams}, nil
}

// NewBatchFloat32Multi is NewBatchFloat32 over several GPUs inside ONE handle (gar_create_multi): the streams are sharded
// by contiguous blocks over `devices`, each shard with its own device state and a worker thread bound to the device's NUMA
// node; Process / Flush fan out and join inside the call — the GPU analogue of the per-channel goroutines of
// constant.go:223-241. A single-process Go caller reaches the multi-GPU numbers without running one process per GPU.
func NewBatchFloat32Multi(in, out float64, q QualityPreset, streams int, devices []int) (*Batch, error) {
	if len(devices) == 0 {
		return nil, fmt.Errorf("%w: empty device list", ErrInvalidConfig)
	}
	cfg := C.gar_config{input_rate: C.double(in), output_rate: C.double(out), channels: 1,
		path: C.GAR_PATH_ENGINE, preset: C.int32_t(q), dtype: C.GAR_F32, engine_quality: -1,
		n_streams: C.int32_t(streams)}
	devs := make([]C.int32_t, len(devices))
	for i, d := range devices {
		devs[i] = C.int32_t(d)
	}
	var h *C.gar_handle
	if st := C.gar_create_multi(&cfg, &devs[0], C.int32_t(len(devs)), &h); st != C.GAR_OK {
		return nil, statusErr(st, nil)
	}
	r := &handle{h: h}
	runtime.SetFinalizer(r, func(r *handle) { C.gar_destroy(r.h) })
	return &Batch{r, streams}, nil
}

// HostAllocRows returns a pinned planar [Streams][cols] float32 buffer whose row blocks live on the NUMA node of the
// device that copies them (gar_host_alloc_rows); pass it to Process / Flush for full PCIe rate. Release with HostFree.
func (b *Batch) HostAllocRows(cols int) []float32 {
	p := C.gar_host_alloc_rows(b.h, C.size_t(cols)*4)
	runtime.KeepAlive(b)
	if p == nil {
		return nil
	}
	return unsafe.Slice((*float32)(p), b.Streams*cols)
}

// HostFree releases a buffer made by HostAllocRows.
func HostFree(buf []float32) {
	if len(buf) > 0 {
		C.gar_host_free(unsafe.Pointer(&buf[0]))
	}
}

// Process resamples planar [Streams][nIn] float32 (row stride nIn) into planar [Streams][outStride].
// The C side reads Streams*nIn and writes up to Streams*outStride elements, so both slices are validated here:
// nothing may be written past a Go slice.
func (b *Batch) Process(in []float32, nIn int, out []float32, outStride int) (int, error) {
	if nIn < 0 || outStride < 0 || b.Streams <= 0 {
		return 0, fmt.Errorf("%w: negative length", ErrInvalidConfig)
	}
	if nIn == 0 {
		return 0, nil
	}
	if len(in) < b.Streams*nIn {
		return 0, fmt.Errorf("%w: input holds %d samples, %d streams x %d needed", ErrInvalidConfig, len(in), b.Streams, nIn)
	}
	if outStride < b.EstimateOutput(nIn) || len(out) < b.Streams*outStride {
		return 0, ErrBufferTooSmall
	}
	var n C.int64_t
	st := C.gar_process_batch(b.h, C.GAR_F32, unsafe.Pointer(&in[0]), C.int64_t(nIn), C.int64_t(nIn),
		unsafe.Pointer(&out[0]), C.int64_t(outStride), C.int64_t(outStride), &n)
	err := statusErr(st, b.h)
	runtime.KeepAlive(in)
	runtime.KeepAlive(out)
	runtime.KeepAlive(b)
	return int(n), err
}

// Flush drains every stream of the batch (planar [Streams][outStride]); returns samples per row.
func (b *Batch) Flush(out []float32, outStride int) (int, error) {
	need := b.nextFlushCount(0)
	if need == 0 {
		return 0, nil
	}
	if outStride < need || len(out) < b.Streams*outStride {
		return 0, ErrBufferTooSmall
	}
	var n C.int64_t
	st := C.gar_flush_batch(b.h, C.GAR_F32, unsafe.Pointer(&out[0]), C.int64_t(outStride), C.int64_t(outStride), &n)
	err := statusErr(st, b.h)
	runtime.KeepAlive(out)
	runtime.KeepAlive(b)
	return int(n), err
}

// Info mirrors resample.go:295-316; GetInfo: constant.go:452-485.
type Info struct {
	Algorithm    string
	FilterLength int
	Phases       int
	Latency      int
	MemoryUsage  int64
	SIMDEnabled  bool
	SIMDType     string
}

func (r *handle) GetInfo() Info {
	var ci C.gar_info
	C.gar_get_info(r.h, &ci)
	runtime.KeepAlive(r)
	return Info{
		Algorithm:    C.GoString(&ci.algorithm[0]),
		FilterLength: int(ci.filter_length),
		Phases:       int(ci.phases),
		Latency:      int(ci.latency),
		MemoryUsage:  int64(ci.memory_usage),
		SIMDEnabled:  ci.simd_enabled != 0,
		SIMDType:     C.GoString(&ci.simd_type[0]),
	}
}

// resampleAll: Process + Flush of one channel (convenience.go:206-229).
func resampleAll(r *SimpleResampler, input []float64) ([]float64, error) {
	out, err := r.Process(input)
	if err != nil {
		return nil, err
	}
	fl, err := r.Flush()
	if err != nil {
		return nil, err
	}
	return append(out, fl...), nil
}

// ResampleStereo: one engine for both channels, Reset() in between (convenience.go:233-257).
func ResampleStereo(left, right []float64, inRate, outRate float64, q QualityPreset) (leftOut, rightOut []float64, err error) {
	r, err := NewEngine(inRate, outRate, q)
	if err != nil {
		return nil, nil, err
	}
	if leftOut, err = resampleAll(r, left); err != nil {
		return nil, nil, err
	}
	r.Reset()
	if rightOut, err = resampleAll(r, right); err != nil {
		return nil, nil, err
	}
	return leftOut, rightOut, nil
}

// ProcessInterleavedInt16 is one block of the resample-wav loop (cmd/resample-wav/helpers.go:77-334) for packed
// int16 PCM: de-interleave + normalise, resample every channel, clamp + quantise + interleave on the device.
// `r` must have been created with Channels = the number of interleaved channels.
func (r *Resampler) ProcessInterleavedInt16(frames []int16, out []int16) (int, error) {
	ch := r.channels
	if ch <= 0 || len(frames)%ch != 0 {
		return 0, fmt.Errorf("%w: interleaved length %d is not a multiple of %d channels", ErrInvalidConfig, len(frames), ch)
	}
	if len(frames) == 0 {
		return 0, nil
	}
	if len(out)/ch < r.EstimateOutput(len(frames)/ch) { // also rejects an empty `out` before &out[0]
		return 0, ErrBufferTooSmall
	}
	var got C.int64_t
	st := C.gar_process_interleaved(r.h, C.GAR_FMT_I16, 16, unsafe.Pointer(&frames[0]), C.int64_t(len(frames)/ch),
		unsafe.Pointer(&out[0]), C.int64_t(len(out)/ch), &got)
	err := statusErr(st, r.h)
	runtime.KeepAlive(frames)
	runtime.KeepAlive(out)
	runtime.KeepAlive(r)
	return int(got) * ch, err
}
