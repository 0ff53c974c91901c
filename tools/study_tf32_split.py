#!/usr/bin/env python
"""Error study for a split-TF32 tensor-core form of the float32 decimator (BASELINE config 4, 877-tap /3 FIR) — the
pre-condition DESIGN.md sets before such a kernel may replace the 1e-6-exact FFMA2 kernel (north star: float32 path within
1e-6 max abs error of the reference).

Emulated on the CPU with numpy, bit-faithfully where it matters: TF32 operands (10 explicit mantissa bits; hi = rna(x),
lo = rna(x - hi)), products exact, accumulation in float32 per k-block of 8 taps with the tensor core's truncating
(round-toward-zero) adds emulated pessimistically, partial sums over `chunk` taps folded into float64 (as the shipped kernel
folds its float32 partials). Variants: 1xTF32, 3xTF32 (hi*hi + hi*lo + lo*hi), 4xTF32 (+ lo*lo). Referee: the same float32
operands in exact float64 arithmetic (what the Go reference's float32 path approximates within its own rounding ~1e-7).

    python tools/study_tf32_split.py [--rows 8] [--seconds 2]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import G, O, sig_c4  # noqa: E402


def tf32(x):
    """round-to-nearest-even of float32 to TF32 (10 explicit mantissa bits), returned as float32"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def rtz32(x64):
    """float64 -> float32 with truncation toward zero (pessimistic model of the tensor-core accumulator)"""
    f = x64.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def fir_split(x, c, M, n_out, first, variant, chunk):
    """out[j] = sum_k x[first + j*M + k] * c[k] with split-TF32 products, float32 accumulation per chunk, float64 across."""
    taps = len(c)
    idx = first + M * np.arange(n_out)[:, None]
    ch, cl = tf32(c), tf32(c - tf32(c))
    total = np.zeros(n_out, dtype=np.float64)
    for k0 in range(0, taps, chunk):
        acc = np.zeros(n_out, dtype=np.float32)
        for kb in range(k0, min(taps, k0 + chunk), 8):  # one m16n8k8 k-step: exact products, one truncating add
            k = np.arange(kb, min(taps, kb + 8, k0 + chunk))
            xs = x[idx + k[None, :]]
            xh, xl = tf32(xs), tf32(xs - tf32(xs))
            p = xh.astype(np.float64) * ch[k].astype(np.float64)
            if variant >= 3:
                p = p + xh.astype(np.float64) * cl[k] + xl.astype(np.float64) * ch[k]
            if variant >= 4:
                p = p + xl.astype(np.float64) * cl[k]
            acc = rtz32(acc.astype(np.float64) + p.sum(axis=1))
        total += acc.astype(np.float64)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=2.0)
    a = ap.parse_args()
    n = int(a.seconds * 48000)
    x = sig_c4(a.rows, n)  # float32 rows of the bench workload
    # 877 taps in stored (reversed) order, cast to float32 like the engine does (dft_stage.go:464); geometry-only handle
    c = G.NewBatch(48000, 16000, G.QualityMedium, 1, np.float32, device=-1).bank(0).astype(np.float32)
    taps, M = len(c), 3
    n_out = (n - taps) // M + 1
    print(f"C4 rows: {a.rows} x {n} samples, {taps} taps /{M}; sum|c| = {np.abs(c.astype(np.float64)).sum():.3f}, "
          f"max|x| = {np.abs(x).max():.3f}")
    print(f"{'variant':34s} {'max |err|':>12s} {'rms err':>12s}   (vs exact float64 on the same float32 operands)")
    rows = []
    for r in range(a.rows):
        xr = x[r]
        idx = M * np.arange(n_out)[:, None] + np.arange(taps)[None, :]
        exact = (xr[idx].astype(np.float64) * c.astype(np.float64)[None, :]).sum(axis=1)
        rows.append((xr, exact))
    for name, variant, chunk in (("1xTF32, fp32 accumulate (877 taps)", 1, 1 << 20),
                                 ("3xTF32, fp32 accumulate (877 taps)", 3, 1 << 20),
                                 ("3xTF32, float64 fold every 64 taps", 3, 64),
                                 ("3xTF32, float64 fold every 16 taps", 3, 16),
                                 ("4xTF32, float64 fold every 64 taps", 4, 64),
                                 ("4xTF32, float64 fold every 16 taps", 4, 16)):
        worst, sq, cnt = 0.0, 0.0, 0
        for xr, exact in rows:
            got = fir_split(xr, c, M, n_out, 0, variant, chunk)
            d = got - exact
            worst = max(worst, float(np.abs(d).max()))
            sq += float((d * d).sum())
            cnt += d.size
        print(f"{name:34s} {worst:12.3e} {np.sqrt(sq / cnt):12.3e}")
    # the shipped kernel's own figure on the same rows, for scale
    y, counts = O.batch_resample(x, 48000, 16000, O.Q_MEDIUM, n_threads=4, flush=False)
    worst = max(float(np.abs(y[r, :n_out].astype(np.float64) - rows[r][1]).max()) for r in range(a.rows))
    print(f"{'float32 oracle (AVX2 FMA order)':34s} {worst:12.3e}   (includes the final rounding to float32: 6e-8)")


if __name__ == "__main__":
    main()
