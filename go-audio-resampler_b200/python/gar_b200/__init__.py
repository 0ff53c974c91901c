"""gar_b200 — Python mirror of the go-audio-resampler Go API over the B200 engine's C ABI.

Names, argument meaning and error behaviour follow the reference's Go surface
(`resample.go`, `constant.go`, `convenience.go`; paths relative to the reference
repository) so the parity tests read like the reference's own tests:

    New(Config)                  -> Resampler   (Process, ProcessInto, ProcessFloat32, ProcessFloat32Into,
                                                 ProcessMulti, Flush, FlushMulti, EstimateOutput, GetLatency,
                                                 Reset, GetRatio, GetInfo)
    NewEngine / NewEngineFloat32 -> SimpleResampler[Float32]
    ResampleMono[Float32], ResampleStereo[Float32]
    NewBatch                     -> BatchResampler (extension: many independent streams)

Everything here is a thin ctypes shim over `include/gar.h`; all arithmetic runs
in the hand-written sm_100a kernels of libgar_b200.so.  There is no CPU
fallback: importing works without a GPU (so symbols can be checked), but any
constructor raises CudaError when no device is usable, and a missing library
raises ImportError.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent.parent  # go-audio-resampler_b200/
_LIB_PATH = Path(os.environ.get("GAR_B200_LIB", _ROOT / "_build" / "libgar_b200.so"))

# gar_status
OK, INVALID_CONFIG, BUFFER_TOO_SMALL, NOT_SUPPORTED, CUDA_ERROR, INTERNAL = range(6)
PATH_PIPELINE, PATH_ENGINE = 0, 1
F64, F32 = 0, 1
# QualityPreset (resample.go:108-131)
QualityQuick, QualityLow, QualityMedium, QualityHigh, QualityVeryHigh, QualityCustom = range(6)
# engine.Quality (internal/engine/filter_params.go:16-42)
(EngineQualityQuick, EngineQualityLow, EngineQualityMedium, EngineQualityHigh, EngineQualityVeryHigh,
 EngineQuality16Bit, EngineQuality20Bit, EngineQuality24Bit, EngineQuality28Bit, EngineQuality32Bit) = range(10)
STAGE_UP, STAGE_DECIM, STAGE_POLY, STAGE_CUBIC = range(4)


class ErrInvalidConfig(ValueError):
    """resample.go:158 ErrInvalidConfig"""


class ErrBufferTooSmall(BufferError):
    """resample.go:161 ErrBufferTooSmall"""


class ErrNotSupported(NotImplementedError):
    """resample.go:164 ErrNotSupported"""


class CudaError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("input_rate", C.c_double), ("output_rate", C.c_double), ("channels", C.c_int32),
                ("path", C.c_int32), ("preset", C.c_int32), ("custom_precision", C.c_int32),
                ("custom_phase_response", C.c_double), ("custom_passband_end", C.c_double),
                ("custom_stopband_begin", C.c_double), ("dtype", C.c_int32), ("engine_quality", C.c_int32),
                ("n_streams", C.c_int32), ("device", C.c_int32), ("max_input_size", C.c_int32),
                ("flags", C.c_uint32)]


class StageDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("engine_index", C.c_int32), ("factor", C.c_int32), ("taps", C.c_int32),
                ("proto_taps", C.c_int32), ("engine_quality", C.c_int32), ("step", C.c_int64), ("at", C.c_int64),
                ("hist_len", C.c_int64), ("decim_phase", C.c_int64), ("ratio", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _Info(C.Structure):
    _fields_ = [("algorithm", C.c_char * 32), ("filter_length", C.c_int32), ("phases", C.c_int32),
                ("latency", C.c_int32), ("memory_usage", C.c_int64), ("simd_enabled", C.c_int32),
                ("simd_type", C.c_char * 64)]


# every symbol include/gar.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pi64 = C.POINTER(C.c_int64)
SYMBOLS = {
    "gar_create": (_i32, [C.POINTER(_Config), C.POINTER(_vp)]),
    "gar_create_multi": (_i32, [C.POINTER(_Config), C.POINTER(C.c_int32), _i32, C.POINTER(_vp)]),
    "gar_num_devices": (_i32, [_vp]),
    "gar_shard_info": (_i32, [_vp, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gar_destroy": (None, [_vp]),
    "gar_last_error": (C.c_char_p, [_vp]),
    "gar_status_string": (C.c_char_p, [_i32]),
    "gar_estimate_output": (_i64, [_vp, _i64]),
    "gar_next_output_count": (_i64, [_vp, _i32, _i64]),
    "gar_next_flush_count": (_i64, [_vp, _i32]),
    "gar_get_ratio": (_d, [_vp]),
    "gar_get_latency": (_i32, [_vp]),
    "gar_get_info": (_i32, [_vp, C.POINTER(_Info)]),
    "gar_get_stats": (_i32, [_vp, _i32, _i32, _pi64, _pi64]),
    "gar_num_stages": (_i32, [_vp]),
    "gar_num_engines": (_i32, [_vp]),
    "gar_describe_stage": (_i32, [_vp, _i32, _i32, C.POINTER(StageDesc)]),
    "gar_plan_stage_type": (_i32, [_vp, _i32]),
    "gar_get_bank": (_i64, [_vp, _i32, _i32, _vp, _i64]),
    "gar_upload_bank": (_i32, [_vp, _i32, _i32, _vp, _i64]),
    "gar_process_f64": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _pi64]),
    "gar_process_f32": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _pi64]),
    "gar_process_multi_f64": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "gar_flush_f64": (_i32, [_vp, _i32, _vp, _i64, _pi64]),
    "gar_flush_f32": (_i32, [_vp, _i32, _vp, _i64, _pi64]),
    "gar_flush_multi_f64": (_i32, [_vp, _vp, _i64, _vp]),
    "gar_advance_geometry": (_i32, [_vp, _i32, _i64, _i32, _pi64]),
    "gar_reset": (_i32, [_vp]),
    "gar_process_batch": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _pi64]),
    "gar_flush_batch": (_i32, [_vp, _i32, _vp, _i64, _i64, _pi64]),
    "gar_process_batch_dev": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _pi64, _vp]),
    "gar_flush_batch_dev": (_i32, [_vp, _i32, _vp, _i64, _i64, _pi64, _vp]),
    "gar_process_interleaved": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _i64, _pi64]),
    "gar_flush_interleaved": (_i32, [_vp, _i32, _i32, _vp, _i64, _pi64]),
    "gar_host_alloc": (_vp, [C.c_size_t]),
    "gar_host_free": (None, [_vp]),
    "gar_host_alloc_wc": (_vp, [C.c_size_t]),
    "gar_host_alloc_rows": (_vp, [_vp, C.c_size_t]),
    "gar_memcpy_async": (_i32, [_vp, _vp, C.c_size_t, _i32, _vp]),
    "gar_bind_thread_to_device": (_i32, [_i32]),
    "gar_device_numa_node": (_i32, [_i32]),
    "gar_device_count": (_i32, []),
    "gar_set_fusion": (_i32, [_vp, _i32]),
    "gar_kernel_launches": (_i64, [_vp, _i32]),
    "gar_stage_kernel_name": (C.c_char_p, [_vp, _i32]),
    "gar_kernels_used": (_i32, [_vp, C.c_char_p, _i32]),
    "gar_set_tiled_polyphase": (None, [_i32]),
    "gar_set_slice_budget": (_i32, [_vp, _i64]),
    "gar_set_tensor_fir": (None, [_i32]),
    "gar_set_chain_kernel": (None, [_i32]),
    "gar_debug_chain_tile_hi": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i32]),
    "gar_measure_fma_peak": (_i32, [_i32, _i32, C.POINTER(C.c_double)]),
    "gar_version": (C.c_char_p, []),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load libgar_b200.so (built by `make -C go-audio-resampler_b200`). Never falls back to anything else."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(f"{_LIB_PATH} is missing: build it with `make -C {_ROOT}` "
                              "(the B200 engine has no CPU fallback)")
        L = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _raise(status: int, h=None):
    msg = lib().gar_last_error(h)
    msg = msg.decode() if msg else lib().gar_status_string(status).decode()
    if status == INVALID_CONFIG:
        raise ErrInvalidConfig(msg)
    if status == BUFFER_TOO_SMALL:
        raise ErrBufferTooSmall(msg)
    if status == NOT_SUPPORTED:
        raise ErrNotSupported(msg)
    if status == CUDA_ERROR:
        raise CudaError(msg)
    raise RuntimeError(f"gar status {status}: {msg}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class QualitySpec:  # resample.go:77-102
    Preset: int = QualityMedium
    Precision: int = 0
    PhaseResponse: float = 0.0
    PassbandEnd: float = 0.0
    StopbandBegin: float = 0.0
    Flags: int = 0


@dataclass
class Config:  # resample.go:46-73
    InputRate: float = 0.0
    OutputRate: float = 0.0
    Channels: int = 0
    Quality: QualitySpec = field(default_factory=QualitySpec)
    MaxInputSize: int = 0
    EnableSIMD: bool = False
    EnableParallel: bool = False
    Device: int = 0  # extension: CUDA ordinal


class _Handle:
    def __init__(self, cfg: _Config, devices=None):
        h = C.c_void_p()
        if devices is None:
            st = lib().gar_create(C.byref(cfg), C.byref(h))
        else:  # one handle, rows sharded over several devices (gar_create_multi)
            devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            st = lib().gar_create_multi(C.byref(cfg), devs, len(devices), C.byref(h))
        if st != OK:
            _raise(st, None)
        self._h = h
        self.rows = max(1, cfg.channels) * max(1, cfg.n_streams)

    def shards(self):
        """[(device, row0, rows, numa_node)] — one entry per device behind this handle."""
        out = []
        for k in range(lib().gar_num_devices(self._h)):
            d, r0, n, node = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
            lib().gar_shard_info(self._h, k, C.byref(d), C.byref(r0), C.byref(n), C.byref(node))
            out.append((d.value, r0.value, n.value, node.value))
        return out

    def host_alloc_rows(self, n_cols, dtype):
        """Pinned [rows, n_cols] host array whose row blocks sit on the NUMA node of the device that copies them."""
        dt = np.dtype(dtype)
        p = lib().gar_host_alloc_rows(self._h, int(n_cols) * dt.itemsize)
        if not p:
            raise CudaError("gar_host_alloc_rows failed")
        buf = (C.c_char * (self.rows * int(n_cols) * dt.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=self.rows * int(n_cols)).reshape(self.rows, int(n_cols)), p

    def close(self):
        if getattr(self, "_h", None):
            lib().gar_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- shared helpers ---
    def describe(self, stream=0):
        out = []
        for s in range(lib().gar_num_stages(self._h)):
            d = StageDesc()
            lib().gar_describe_stage(self._h, stream, s, C.byref(d))
            out.append(d.as_dict())
        return out

    def bank(self, stage, which=0):
        n = -lib().gar_get_bank(self._h, stage, which, None, 0)
        if n <= 0:
            return np.zeros(0)
        out = np.empty(n, dtype=np.float64)
        got = lib().gar_get_bank(self._h, stage, which, _ptr(out), n)
        assert got == n
        return out

    def upload_bank(self, stage, which, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        st = lib().gar_upload_bank(self._h, stage, which, _ptr(coef), len(coef))
        if st != OK:
            _raise(st, self._h)

    def kernel_names(self):
        return [lib().gar_stage_kernel_name(self._h, s).decode() for s in range(lib().gar_num_stages(self._h))]

    def last_kernels(self):
        """Distinct kernel variants launched through this handle so far."""
        buf = C.create_string_buffer(1024)
        lib().gar_kernels_used(self._h, buf, 1024)
        return [k for k in buf.value.decode().split(",") if k]

    def set_slice_budget(self, nbytes: int):
        """Inter-stage buffer budget of one time slice of a long call (0 disables slicing)."""
        lib().gar_set_slice_budget(self._h, int(nbytes))

    def set_fusion(self, enabled: bool):
        """Enable/disable the fused x2 -> polyphase kernel (K4)."""
        lib().gar_set_fusion(self._h, 1 if enabled else 0)

    def plan_types(self):
        return [lib().gar_plan_stage_type(self._h, e) for e in range(lib().gar_num_engines(self._h))]

    def advance_geometry(self, n_in, flush=False, stream=0):
        """Advance the integer state only (no samples); returns the count Process/Flush would return."""
        n = C.c_int64(0)
        st = lib().gar_advance_geometry(self._h, stream, int(n_in), 1 if flush else 0, C.byref(n))
        if st != OK:
            _raise(st, self._h)
        return n.value

    def EstimateOutput(self, n):  # constant.go:117-119
        return int(lib().gar_estimate_output(self._h, int(n)))

    def GetRatio(self):
        return lib().gar_get_ratio(self._h)

    def GetLatency(self):
        return lib().gar_get_latency(self._h)

    def Reset(self):
        lib().gar_reset(self._h)

    def GetStatistics(self, stream=0, engine=0):  # resampler.go:348-353
        a, b = C.c_int64(0), C.c_int64(0)
        lib().gar_get_stats(self._h, stream, engine, C.byref(a), C.byref(b))
        return {"samplesIn": a.value, "samplesOut": b.value}

    def GetInfo(self):
        info = _Info()
        lib().gar_get_info(self._h, C.byref(info))
        return {"Algorithm": info.algorithm.decode(), "FilterLength": info.filter_length, "Phases": info.phases,
                "Latency": info.latency, "MemoryUsage": info.memory_usage, "SIMDEnabled": bool(info.simd_enabled),
                "SIMDType": info.simd_type.decode()}

    # --- batched rows (all channels x n_streams rows advance in lock step) ---
    def _native_dtype(self):
        return getattr(self, "dtype", np.float64)

    def _io_code(self, io_dtype):
        dt = np.dtype(io_dtype if io_dtype is not None else self._native_dtype())
        return F32 if dt == np.float32 else F64

    def next_output_count(self, n_in):
        return int(lib().gar_next_output_count(self._h, 0, int(n_in)))

    def next_flush_count(self):
        return int(lib().gar_next_flush_count(self._h, 0))

    def ProcessBatch(self, x2d, out2d=None, io_dtype=None):
        """x2d: host [n_streams, n_in] (row stride in elements may exceed n_in). Returns ([n_streams, n_out] view, n_out)."""
        io_dtype = x2d.dtype
        assert x2d.shape[0] == self.rows and x2d.strides[1] == x2d.itemsize
        n_in = x2d.shape[1]
        if out2d is None:
            out2d = np.empty((self.rows, self.EstimateOutput(n_in)), dtype=io_dtype)
        n = C.c_int64(0)
        st = lib().gar_process_batch(self._h, self._io_code(io_dtype), _ptr(x2d), x2d.strides[0] // x2d.itemsize, n_in,
                                     _ptr(out2d), out2d.strides[0] // out2d.itemsize, out2d.shape[1], C.byref(n))
        if st != OK:
            _raise(st, self._h)
        return out2d[:, :n.value], n.value

    def FlushBatch(self, out2d=None, io_dtype=None):
        if out2d is None:
            out2d = np.empty((self.rows, max(self.next_flush_count(), 1)), dtype=io_dtype or self._native_dtype())
        io_dtype = out2d.dtype
        n = C.c_int64(0)
        st = lib().gar_flush_batch(self._h, self._io_code(io_dtype), _ptr(out2d), out2d.strides[0] // out2d.itemsize,
                                   out2d.shape[1], C.byref(n))
        if st != OK:
            _raise(st, self._h)
        return out2d[:, :n.value], n.value

    # device-pointer variants (torch tensors or raw pointers); enqueue only.
    # `stream` is a raw cudaStream_t: pass the stream that produces d_in and consumes d_out (with torch:
    # torch.cuda.current_stream().cuda_stream). stream=0 means the handle's own non-blocking stream, which is NOT ordered
    # against the caller's streams: the caller must then synchronise d_in before and d_out after the call itself. Calls on
    # one handle are ordered against each other whatever streams they use (Engine::order_before/after).
    def process_batch_dev(self, d_in, in_stride, n_in, d_out, out_stride, out_cap, stream=0, io_dtype=None):
        n = C.c_int64(0)
        st = lib().gar_process_batch_dev(self._h, self._io_code(io_dtype), C.c_void_p(d_in), in_stride, n_in, C.c_void_p(d_out),
                                         out_stride, out_cap, C.byref(n), C.c_void_p(stream))
        if st != OK:
            _raise(st, self._h)
        return n.value

    def flush_batch_dev(self, d_out, out_stride, out_cap, stream=0, io_dtype=None):
        n = C.c_int64(0)
        st = lib().gar_flush_batch_dev(self._h, self._io_code(io_dtype), C.c_void_p(d_out), out_stride, out_cap, C.byref(n),
                                       C.c_void_p(stream))
        if st != OK:
            _raise(st, self._h)
        return n.value

    # --- interleaved / integer-PCM boundary (cmd/resample-wav/main.go:358-520) ---
    _FMT = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3,
            np.dtype(np.int64): 4}

    def ProcessInterleaved(self, frames, bit_depth=16):
        """frames: [n_frames, channels] (or flat interleaved) float32/float64/int16/int32/int64; returns same layout."""
        x = np.ascontiguousarray(frames)
        fmt = self._FMT[x.dtype]
        n = x.size // self.rows
        cap = max(self.EstimateOutput(n), int(lib().gar_next_output_count(self._h, 0, n)), 1)
        out = np.empty((cap, self.rows), dtype=x.dtype)
        got = C.c_int64(0)
        st = lib().gar_process_interleaved(self._h, fmt, int(bit_depth), _ptr(x), n, _ptr(out), cap, C.byref(got))
        if st != OK:
            _raise(st, self._h)
        return out[:got.value].copy()

    def FlushInterleaved(self, dtype, bit_depth=16):
        dt = np.dtype(dtype)
        cap = max(int(lib().gar_next_flush_count(self._h, 0)), 1)
        out = np.empty((cap, self.rows), dtype=dt)
        got = C.c_int64(0)
        st = lib().gar_flush_interleaved(self._h, self._FMT[dt], int(bit_depth), _ptr(out), cap, C.byref(got))
        if st != OK:
            _raise(st, self._h)
        return out[:got.value].copy()

    def _process(self, ch, x, dtype, out=None):
        fn = lib().gar_process_f32 if dtype == np.float32 else lib().gar_process_f64
        x = np.ascontiguousarray(x, dtype=dtype)
        owned = out is None
        if owned:  # Process returns an owned, exactly-sized slice (constant.go:88-96)
            cap = max(self.EstimateOutput(len(x)), int(lib().gar_next_output_count(self._h, ch, len(x))))
            out = np.empty(cap, dtype=dtype)
        n = C.c_int64(0)
        st = fn(self._h, ch, _ptr(x), len(x), _ptr(out), len(out), C.byref(n))
        if st != OK:
            _raise(st, self._h)
        return out[:n.value].copy() if owned else n.value

    def _flush(self, ch, dtype):
        fn = lib().gar_flush_f32 if dtype == np.float32 else lib().gar_flush_f64
        cap = max(int(lib().gar_next_flush_count(self._h, ch)), 1)
        out = np.empty(cap, dtype=dtype)
        n = C.c_int64(0)
        st = fn(self._h, ch, _ptr(out), cap, C.byref(n))
        if st != OK:
            _raise(st, self._h)
        return out[:n.value].copy()


class Resampler(_Handle):
    """constantRateResampler behind New(Config) (constant.go:16-485)."""

    def __init__(self, config: Config, n_streams: int = 0, io_dtype=np.float64, devices=None):
        if config is None:
            raise ErrInvalidConfig("config is nil")
        q = config.Quality
        cfg = _Config(float(config.InputRate), float(config.OutputRate), int(config.Channels), PATH_PIPELINE,
                      int(q.Preset), int(q.Precision), float(q.PhaseResponse), float(q.PassbandEnd),
                      float(q.StopbandBegin), F64, -1, int(n_streams), int(config.Device),
                      int(config.MaxInputSize), int(q.Flags) | (int(config.EnableParallel) << 16))
        super().__init__(cfg, devices)
        self.channels = int(config.Channels)

    def Process(self, x):  # constant.go:88-96 (channel 0)
        return self._process(0, x, np.float64)

    def ProcessInto(self, x, out):  # constant.go:103-112
        assert out.dtype == np.float64
        return self._process(0, x, np.float64, out)

    def ProcessFloat32(self, x):  # constant.go:128-147
        return self._process(0, x, np.float32)

    def ProcessFloat32Into(self, x, out):  # constant.go:161-199
        assert out.dtype == np.float32
        return self._process(0, x, np.float32, out)

    def ProcessMulti(self, xs):  # constant.go:204-252
        if len(xs) != self.channels:
            raise ValueError(f"expected {self.channels} channels, got {len(xs)}")
        xs = [np.ascontiguousarray(x, dtype=np.float64) for x in xs]
        nin = np.array([len(x) for x in xs], dtype=np.int64)
        cap = max(int(lib().gar_next_output_count(self._h, c, int(nin[c]))) for c in range(self.channels))
        cap = max(cap, 1)
        outs = [np.empty(cap, dtype=np.float64) for _ in xs]
        ip = (C.c_void_p * self.channels)(*[x.ctypes.data for x in xs])
        op = (C.c_void_p * self.channels)(*[o.ctypes.data for o in outs])
        nout = np.zeros(self.channels, dtype=np.int64)
        st = lib().gar_process_multi_f64(self._h, ip, _ptr(nin), op, cap, _ptr(nout))
        if st != OK:
            _raise(st, self._h)
        return [o[:int(n)].copy() for o, n in zip(outs, nout)]

    def Flush(self):  # constant.go:349-354 (channel 0 only)
        return self._flush(0, np.float64)

    def FlushMulti(self):  # constant.go:390-404
        cap = max(max(int(lib().gar_next_flush_count(self._h, c)) for c in range(self.channels)), 1)
        outs = [np.empty(cap, dtype=np.float64) for _ in range(self.channels)]
        op = (C.c_void_p * self.channels)(*[o.ctypes.data for o in outs])
        nout = np.zeros(self.channels, dtype=np.int64)
        st = lib().gar_flush_multi_f64(self._h, op, cap, _ptr(nout))
        if st != OK:
            _raise(st, self._h)
        return [o[:int(n)].copy() for o, n in zip(outs, nout)]


def New(config: Config) -> Resampler:  # resample.go:272-292
    return Resampler(config)


class SimpleResampler(_Handle):
    """SimpleResampler / SimpleResamplerFloat32 (convenience.go:118-186, 315-395)."""

    def __init__(self, input_rate, output_rate, quality, dtype=np.float64, engine_quality=-1, device=0, n_streams=0,
                 devices=None):
        self.dtype = np.dtype(dtype).type
        cfg = _Config(float(input_rate), float(output_rate), 1, PATH_ENGINE, int(quality), 0, 0.0, 0.0, 0.0,
                      F32 if self.dtype == np.float32 else F64, int(engine_quality), int(n_streams), int(device), 0, 0)
        super().__init__(cfg, devices)

    def Process(self, x):
        return self._process(0, x, self.dtype)

    def ProcessInto(self, x, out):
        assert out.dtype == self.dtype
        return self._process(0, x, self.dtype, out)

    def Flush(self):
        return self._flush(0, self.dtype)


def NewEngine(input_rate, output_rate, quality, **kw):  # convenience.go:125-132
    return SimpleResampler(input_rate, output_rate, quality, np.float64, **kw)


def NewEngineFloat32(input_rate, output_rate, quality, **kw):  # convenience.go:329-336
    return SimpleResampler(input_rate, output_rate, quality, np.float32, **kw)


def _resample_all(r, x):  # convenience.go:217-229
    a = r.Process(x)
    b = r.Flush()
    return np.concatenate([a, b])


def ResampleMono(x, input_rate, output_rate, quality):  # convenience.go:204-211
    return _resample_all(NewEngine(input_rate, output_rate, quality), x)


def ResampleMonoFloat32(x, input_rate, output_rate, quality):  # convenience.go:407-414
    return _resample_all(NewEngineFloat32(input_rate, output_rate, quality), x)


def ResampleStereo(left, right, input_rate, output_rate, quality):  # convenience.go:233-257
    r = NewEngine(input_rate, output_rate, quality)
    lo = _resample_all(r, left)
    r.Reset()
    return lo, _resample_all(r, right)


def ResampleStereoFloat32(left, right, input_rate, output_rate, quality):  # convenience.go:436-457
    r = NewEngineFloat32(input_rate, output_rate, quality)
    lo = _resample_all(r, left)
    r.Reset()
    return lo, _resample_all(r, right)


def InterleaveToStereo(left, right):  # convenience.go:261-269
    n = min(len(left), len(right))
    out = np.empty(2 * n, dtype=np.asarray(left).dtype)
    out[0::2] = left[:n]
    out[1::2] = right[:n]
    return out


def DeinterleaveFromStereo(x):  # convenience.go:273-282
    n = len(x) // 2
    return np.array(x[0:2 * n:2]), np.array(x[1:2 * n:2])


class BatchResampler(SimpleResampler):
    """Extension: `n_streams` independent mono streams advancing in lock step (BASELINE config 4).

    Semantically n_streams separate NewEngine[Float32] instances fed equal-length chunks
    (SURVEY.md CS4); one device pass per chunk.
    """

    def __init__(self, input_rate, output_rate, quality, n_streams, dtype=np.float32, device=0, engine_quality=-1,
                 devices=None):
        super().__init__(input_rate, output_rate, quality, dtype, engine_quality, device, n_streams, devices)
        self.n_streams = int(n_streams)



def NewBatch(input_rate, output_rate, quality, n_streams, dtype=np.float32, **kw):
    return BatchResampler(input_rate, output_rate, quality, n_streams, dtype, **kw)


def host_alloc(shape, dtype):
    """Pinned host ndarray (full PCIe rate for gar_process_batch)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().gar_host_alloc(max(n, 1))
    if not p:
        raise CudaError("cudaMallocHost failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    return arr, p


def host_alloc_wc(shape, dtype):
    """Write-combined pinned host ndarray for input buffers the host only writes."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().gar_host_alloc_wc(max(n, 1))
    if not p:
        raise CudaError("cudaHostAlloc(write-combined) failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape), p


def host_free(p):
    lib().gar_host_free(p)


def device_count():
    return lib().gar_device_count()


def bind_thread_to_device(device: int) -> int:
    """Restrict the calling thread to the cores of the device's NUMA node (-1: nothing changed)."""
    return int(lib().gar_bind_thread_to_device(int(device)))


def device_numa_node(device: int) -> int:
    return int(lib().gar_device_numa_node(int(device)))


def set_tiled_polyphase(enabled: bool):
    """Process-wide A/B switch for the register-tiled polyphase kernels (K4r / K3r / K3i)."""
    lib().gar_set_tiled_polyphase(1 if enabled else 0)


def set_tensor_fir(enabled: bool):
    """Process-wide A/B switch for the FP64 tensor-core (DMMA) FIR kernels."""
    lib().gar_set_tensor_fir(1 if enabled else 0)


def set_chain_kernel(mode):
    """Process-wide policy for the persistent x2 -> polyphase chain kernel (K5) of large float64 batches:
    0 / False never, 1 / True every eligible call, 2 (default) calls that exceed the inter-stage memory budget."""
    lib().gar_set_chain_kernel(int(mode))


def kernel_launches(reset=False):
    return int(lib().gar_kernel_launches(None, 1 if reset else 0))


def measure_fma_peak(dtype=np.float32, device=0, packed=False, tensor=False):
    """Dependent-FMA probe (TFLOP/s). packed=True: fma.rn.f32x2 (FFMA2); tensor=True: FP64 tensor cores (DMMA.8x8x4)."""
    v = C.c_double(0)
    code = 3 if tensor else 2 if packed else (F32 if np.dtype(dtype) == np.float32 else F64)
    st = lib().gar_measure_fma_peak(device, code, C.byref(v))
    if st != OK:
        _raise(st, None)
    return v.value
