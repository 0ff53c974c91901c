"""The reference's own audio-quality measurement procedures, restated for numpy (test infrastructure).

thd_db / snr_db: measureTHDInternal / measureSNRInternal (internal/engine/quality_regression_test.go:292-425) — Hann window
over the FIRST 16 384 output samples (start-up transient included), single-bin magnitudes.
precision_thd_db: precisionMeasureTHD (internal/engine/precision_comparison_test.go:553-603) — 8192 samples from the middle.
"""
import numpy as np

N, FFT = 65536, 16384


def thd_sine(rate, n=N):  # quality_regression_test.go:298-303
    return 0.9 * np.sin(2.0 * np.pi * 1000.0 * np.arange(n) / rate)


def _spectrum(y):
    w = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(FFT) / (FFT - 1)))
    seg = np.zeros(FFT)
    m = min(FFT, len(y))
    seg[:m] = y[:m]
    return np.abs(np.fft.fft(seg * w))


def thd_db(y, out_rate, f0=1000.0):  # measureTHDInternal, quality_regression_test.go:292-342
    mag = _spectrum(y)
    fund = mag[int(f0 / out_rate * FFT)]
    hp = 0.0
    for h in range(2, 11):
        if f0 * h >= out_rate / 2:
            break
        b = int(f0 * h / out_rate * FFT)
        if b < FFT // 2:
            hp += mag[b] ** 2
    return 20 * np.log10(np.sqrt(hp) / (fund + 1e-20) + 1e-20)


def snr_db(y, out_rate, f0=1000.0):  # measureSNRInternal, quality_regression_test.go:344-425
    mag = _spectrum(y)
    fb = int(f0 / out_rate * FFT)
    sig = sum(mag[fb + b] ** 2 for b in range(-3, 4) if 0 < fb + b < FFT // 2)
    hbins = [int(f0 * h / out_rate * FFT) for h in range(2, 11) if f0 * h < out_rate / 2]
    noise = 0.0
    for b in range(1, FFT // 2):
        if fb - 3 <= b <= fb + 3 or any(hb - 2 <= b <= hb + 2 for hb in hbins):
            continue
        noise += mag[b] ** 2
    return 10 * np.log10(sig + 1e-20) - 10 * np.log10(noise + 1e-20)


def precision_thd_db(out, f0, rate):  # precision_comparison_test.go:553-603
    out = np.asarray(out, dtype=np.float64)
    n = 8192 if len(out) >= 2 * 8192 else 1024
    s = (len(out) - n) // 2
    w = 0.5 * (1 - np.cos(2 * np.pi * np.arange(n) / (n - 1)))
    sp = np.abs(np.fft.fft(out[s:s + n] * w))
    fb = int(round(f0 / (rate / n)))
    hp = sum(sp[fb * h] ** 2 for h in range(2, 6) if fb * h < n // 2)
    return 10 * np.log10(hp / sp[fb] ** 2)
