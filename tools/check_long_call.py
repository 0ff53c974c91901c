"""One very long single-stream call (default 250 M samples, 44.1k->48k QualityHigh float64: time-sliced multi-stage call, 31-bit index\nranges) against the oracle. Run on a B200: python tools/check_long_call.py [n_samples]"""
import sys, time, numpy as np
sys.path.insert(0, "go-audio-resampler_b200/python"); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
from helpers import G, O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
rng = np.random.default_rng(5)
x = rng.standard_normal(n)
t0 = time.perf_counter()
e = G.NewEngine(44100, 48000, G.QualityHigh)
y = e.Process(x); f = e.Flush()
t1 = time.perf_counter()
print(f"gpu: {len(y)} + {len(f)} outputs in {t1-t0:.2f} s; kernels {e.kernel_names() if hasattr(e,'kernel_names') else ''}", flush=True)
w = O.resample_mono(x, 44100, 48000, O.PRESET_HIGH)
t2 = time.perf_counter()
print(f"oracle: {len(w)} outputs in {t2-t1:.1f} s")
got = np.concatenate([y, f])
assert len(got) == len(w), (len(got), len(w))
err = 0.0
for lo in range(0, len(w), 1 << 24):
    err = max(err, float(np.max(np.abs(got[lo:lo + (1 << 24)] - w[lo:lo + (1 << 24)]))))
print("max|err|", err)
assert err <= 1e-12
