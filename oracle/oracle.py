"""ctypes binding of the CPU oracle (oracle/gar_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libgar_oracle.so"

# engine.Quality (internal/engine/filter_params.go:16-42)
Q_QUICK, Q_LOW, Q_MEDIUM, Q_HIGH, Q_VERYHIGH, Q_16BIT, Q_20BIT, Q_24BIT, Q_28BIT, Q_32BIT = range(10)
# QualityPreset (resample.go:108-131)
PRESET_QUICK, PRESET_LOW, PRESET_MEDIUM, PRESET_HIGH, PRESET_VERYHIGH, PRESET_CUSTOM = range(6)
# pipeline.StageType (internal/pipeline/pipeline.go:58-73)
ST_CUBIC, ST_HALFBAND, ST_POLYPHASE, ST_FFT = range(4)

F64, F32 = 0, 1


def build(force: bool = False) -> Path:
    src = _HERE / "gar_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        d, i, sz, vp, l = C.c_double, C.c_int, C.c_size_t, C.c_void_p, C.c_long
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        sig = {
            "orc_bessel_i0": (d, [d]),
            "orc_kaiser_beta": (d, [d]),
            "orc_estimate_filter_length": (i, [d, d]),
            "orc_kaiser_window": (i, [i, d, vp]),
            "orc_design_lowpass": (i, [i, d, d, d, vp]),
            "orc_design_lowpass_auto": (i, [d, d, d, d, vp, i]),
            "orc_quality_attenuation": (d, [i]),
            "orc_quality_passband_end": (d, [i]),
            "orc_is_integer_ratio": (i, [d]),
            "orc_find_rational_approx": (None, [d, ip, ip]),
            "orc_lsx_inv_f_resp": (d, [d, d]),
            "orc_polyphase_params": (None, [i, d, d, i, d, d, vp, vp]),
            "orc_precision_to_engine_quality": (i, [i]),
            "orc_preset_to_engine_quality": (i, [i]),
            "orc_preset_precision": (i, [i]),
            "orc_build_plan": (i, [d, i, vp, vp, i]),
            "orc_dot_f32": (C.c_float, [vp, vp, sz]),
            "orc_dot_f64": (d, [vp, vp, sz]),
            "orc_convolve_valid_f32": (None, [vp, vp, sz, vp, sz]),
            "orc_convolve_valid_f64": (None, [vp, vp, sz, vp, sz]),
            "orc_interleave2_f32": (None, [vp, vp, vp, sz]),
            "orc_interleave2_f64": (None, [vp, vp, vp, sz]),
            "orc_sum_f64": (d, [vp, sz]),
            "orc_sum_f32": (C.c_float, [vp, sz]),
            "orc_scale_f64": (None, [vp, vp, sz, d]),
            "orc_scale_f32": (None, [vp, vp, sz, C.c_float]),
            "orc_cubic_interp_dot_f32": (C.c_float, [vp, vp, vp, vp, vp, C.c_float, sz]),
            "orc_cubic_interp_dot_f64": (d, [vp, vp, vp, vp, vp, d, sz]),
            "orc_engine_new": (vp, [d, d, i, i]),
            "orc_engine_free": (None, [vp]),
            "orc_engine_process": (l, [vp, vp, sz, vp, sz]),
            "orc_engine_flush": (l, [vp, vp, sz]),
            "orc_engine_reset": (None, [vp]),
            "orc_engine_describe": (None, [vp, vp]),
            "orc_engine_ratio": (d, [vp]),
            "orc_engine_stats": (None, [vp, vp, vp]),
            "orc_engine_get_bank": (i, [vp, i, vp, sz]),
            "orc_engine_latency": (i, [vp]),
            "orc_engine_filter_length": (i, [vp]),
            "orc_engine_phases": (i, [vp]),
            "orc_pipeline_new": (vp, [d, d, i, i, i, ip]),
            "orc_pipeline_free": (None, [vp]),
            "orc_pipeline_estimate_output": (i, [vp, l]),
            "orc_pipeline_num_stages": (i, [vp]),
            "orc_pipeline_stage_type": (i, [vp, i]),
            "orc_pipeline_stage_ratio": (d, [vp, i]),
            "orc_pipeline_stage_describe": (None, [vp, i, i, vp]),
            "orc_pipeline_latency": (i, [vp]),
            "orc_pipeline_ratio": (d, [vp]),
            "orc_pipeline_process": (l, [vp, i, vp, sz, vp, sz]),
            "orc_pipeline_process_into": (l, [vp, vp, sz, vp, sz]),
            "orc_pipeline_process_f32_into": (l, [vp, vp, sz, vp, sz]),
            "orc_pipeline_flush": (l, [vp, i, vp, sz]),
            "orc_pipeline_reset": (None, [vp]),
            "orc_batch_resample": (l, [d, d, i, i, vp, sz, sz, vp, sz, vp, i, i]),
            "orc_has_avx2": (i, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


DESC_KEYS = ("cubic", "pre_factor", "pre_tpp", "pre_proto_taps", "pre_hist", "dec_factor", "dec_taps",
             "dec_hist", "dec_phase", "poly_L", "poly_tpp", "poly_step", "poly_at", "poly_hist", "pre_half_band",
             "_r")


class Engine:
    """internal/engine.Resampler[F] (resampler.go:26-353)."""

    def __init__(self, in_rate, out_rate, quality, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self._code = F32 if self.dtype == np.float32 else F64
        self._h = lib().orc_engine_new(float(in_rate), float(out_rate), int(quality), self._code)
        if not self._h:
            raise ValueError("oracle engine construction failed")
        self.ratio = lib().orc_engine_ratio(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_engine_free(self._h)
            self._h = None

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        cap = int(len(x) * self.ratio) + 4096
        out = np.empty(cap, dtype=self.dtype)
        n = lib().orc_engine_process(self._h, _ptr(x), len(x), _ptr(out), cap)
        assert n >= 0, "oracle output capacity too small"
        return out[:n].copy()

    def flush(self):
        cap = 1 << 16
        out = np.empty(cap, dtype=self.dtype)
        n = lib().orc_engine_flush(self._h, _ptr(out), cap)
        assert n >= 0
        return out[:n].copy()

    def reset(self):
        lib().orc_engine_reset(self._h)

    def describe(self):
        v = np.zeros(16, dtype=np.int64)
        lib().orc_engine_describe(self._h, _ptr(v))
        return dict(zip(DESC_KEYS, (int(t) for t in v)))

    def stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        lib().orc_engine_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def bank(self, which):
        cap = 1 << 16
        out = np.empty(cap, dtype=np.float64)
        n = lib().orc_engine_get_bank(self._h, which, _ptr(out), cap)
        assert n >= 0
        return out[:n].copy()

    def latency(self):
        return lib().orc_engine_latency(self._h)

    def filter_length(self):
        return lib().orc_engine_filter_length(self._h)

    def phases(self):
        return lib().orc_engine_phases(self._h)


class Pipeline:
    """constantRateResampler (constant.go:16-485) behind New(Config) (resample.go:272-292)."""

    def __init__(self, in_rate, out_rate, channels=1, preset=PRESET_HIGH, custom_precision=0):
        st = C.c_int(0)
        self._h = lib().orc_pipeline_new(float(in_rate), float(out_rate), int(channels), int(preset),
                                         int(custom_precision), C.byref(st))
        if not self._h:
            raise ValueError(f"invalid resampler configuration (status {st.value})")
        self.channels = channels
        self.ratio = lib().orc_pipeline_ratio(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pipeline_free(self._h)
            self._h = None

    def estimate_output(self, n):
        return lib().orc_pipeline_estimate_output(self._h, int(n))

    def stages(self):
        n = lib().orc_pipeline_num_stages(self._h)
        return [(lib().orc_pipeline_stage_type(self._h, i), lib().orc_pipeline_stage_ratio(self._h, i))
                for i in range(n)]

    def stage_describe(self, ch, i):
        v = np.zeros(16, dtype=np.int64)
        lib().orc_pipeline_stage_describe(self._h, ch, i, _ptr(v))
        return dict(zip(DESC_KEYS, (int(t) for t in v)))

    def latency(self):
        return lib().orc_pipeline_latency(self._h)

    def process(self, x, ch=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        cap = int(len(x) * self.ratio) + 65536
        out = np.empty(cap, dtype=np.float64)
        n = lib().orc_pipeline_process(self._h, ch, _ptr(x), len(x), _ptr(out), cap)
        assert n >= 0
        return out[:n].copy()

    def process_multi(self, xs):
        assert len(xs) == self.channels
        return [self.process(x, ch) for ch, x in enumerate(xs)]

    def process_into(self, x, out):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_pipeline_process_into(self._h, _ptr(x), len(x), _ptr(out), len(out))

    def process_f32_into(self, x, out):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert out.dtype == np.float32
        return lib().orc_pipeline_process_f32_into(self._h, _ptr(x), len(x), _ptr(out), len(out))

    def flush(self, ch=0):
        cap = 1 << 20
        out = np.empty(cap, dtype=np.float64)
        n = lib().orc_pipeline_flush(self._h, ch, _ptr(out), cap)
        assert n >= 0
        return out[:n].copy()

    def flush_multi(self):
        return [self.flush(ch) for ch in range(self.channels)]

    def reset(self):
        lib().orc_pipeline_reset(self._h)


def preset_to_engine_quality(preset):
    return lib().orc_preset_to_engine_quality(int(preset))


def resample_mono(x, in_rate, out_rate, preset, dtype=np.float64):
    """ResampleMono / ResampleMonoFloat32 (convenience.go:204-229,407-429)."""
    e = Engine(in_rate, out_rate, preset_to_engine_quality(preset), dtype)
    a = e.process(x)
    b = e.flush()
    return np.concatenate([a, b])


def build_plan(ratio, precision):
    types = (C.c_int * 32)()
    ratios = (C.c_double * 32)()
    n = lib().orc_build_plan(float(ratio), int(precision), types, ratios, 32)
    if n < 0:
        raise ValueError("invalid ratio")
    return [(types[i], ratios[i]) for i in range(n)]


def batch_resample(x2d, in_rate, out_rate, quality, n_threads=1, flush=True):
    """n_streams independent path-B engines over planar [n_streams][n_in] input."""
    x2d = np.ascontiguousarray(x2d)
    dt = x2d.dtype
    code = F32 if dt == np.float32 else F64
    ns, nin = x2d.shape
    stride = int(nin * (out_rate / in_rate)) + 4096
    out = np.zeros((ns, stride), dtype=dt)
    counts = np.zeros(ns, dtype=np.int64)
    tot = lib().orc_batch_resample(float(in_rate), float(out_rate), int(quality), code, _ptr(x2d), ns, nin,
                                   _ptr(out), stride, _ptr(counts), int(n_threads), 1 if flush else 0)
    if tot < 0:
        raise RuntimeError("oracle batch resample failed")
    return out, counts
