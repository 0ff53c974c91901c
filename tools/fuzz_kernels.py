#!/usr/bin/env python
"""Randomised A/B regression of the fast kernels (K4r / K3r / K3i register-tiled, K1m / K2m / K3m tensor-core) against
the simple one-thread-per-output / vector kernels on the same calls: random rate pairs, presets, row counts, lengths and
chunkings, float64. Counts must be identical, samples within 1e-13 (they have been bit-identical so far).

    python tools/fuzz_kernels.py [--cases 60] [--seed 1]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

RATES = [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 47999, 48000, 88200, 96000, 176400, 192000]
PRESETS = [G.QualityLow, G.QualityMedium, G.QualityHigh, G.QualityVeryHigh]


def run(make, x, cuts, fast):
    G.set_tiled_polyphase(fast)
    G.set_tensor_fir(fast)
    try:
        h = make()
        ys = [h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy() for lo, hi in zip(cuts[:-1], cuts[1:])]
        ys.append(h.FlushBatch()[0].copy())
        return ys, h.last_kernels()
    finally:
        G.set_tiled_polyphase(True)
        G.set_tensor_fir(True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    worst, used = 0.0, set()
    for case in range(a.cases):
        ir, orr = rng.choice(RATES, 2, replace=False)
        ir, orr = int(ir), int(orr)
        if not (1 / 64 <= orr / ir <= 64):
            continue
        preset = PRESETS[int(rng.integers(len(PRESETS)))]
        rows = int(rng.choice([1, 2, 7, 8, 9, 16, 31, 33, 64, 70, 96, 130, 257]))
        n = int(rng.integers(2000, 120000))
        if n * rows * max(1.0, orr / ir) > 3e7:
            n = int(3e7 / (rows * max(1.0, orr / ir)))
        path_a = bool(rng.integers(2))
        if path_a:
            cfg = G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=preset))
            make = lambda: G.Resampler(cfg, n_streams=rows)  # noqa: E731
        else:
            make = lambda: G.NewBatch(ir, orr, preset, rows, np.float64)  # noqa: E731
        x = rng.standard_normal((rows, n))
        k = int(rng.integers(0, 4))
        cuts = sorted(set([0, n] + [int(v) for v in rng.integers(0, n, size=k)]))
        ya, ka = run(make, x, cuts, True)
        yb, kb = run(make, x, cuts, False)
        used.update(ka)
        err = 0.0
        for p, q in zip(ya, yb):
            assert p.shape == q.shape, (case, ir, orr, preset, rows, n, cuts, p.shape, q.shape)
            if p.size:
                err = max(err, float(np.max(np.abs(p - q))))
        worst = max(worst, err)
        tag = "A" if path_a else "B"
        print(f"case {case:3d} {ir:6d}->{orr:6d} q{preset} path{tag} rows {rows:3d} n {n:7d} chunks {len(cuts)-1} "
              f"max|fast-simple| {err:.1e}  fast kernels: {[k_ for k_ in ka if k_ not in kb]}", flush=True)
        assert err <= 1e-13, (case, ir, orr, preset, rows, n, cuts, err)
    print(f"ok: {a.cases} cases, worst {worst:.2e}; kernels exercised: {sorted(used)}")


if __name__ == "__main__":
    main()
