#!/usr/bin/env python
"""Fuzz of the persistent chain kernel K5 against the two stand-alone tensor-core launches: random rate pairs whose engine
is an x2 stage + a polyphase stage, presets, row counts (32 .. 300, ragged groups), call lengths and chunkings, all through
the device entry points of the C ABI. Every sample must be bit-identical (same MMA cores, same order of accumulation).

    python tools/fuzz_chain.py [--cases 40] [--seed 1]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))

import gar_b200 as G  # noqa: E402

PAIRS = [(44100, 48000), (48000, 44100), (44100, 47999), (32000, 44100), (22050, 32000), (48000, 32001), (44100, 64000),
         (96000, 88200), (16000, 22050), (48000, 47000), (37800, 44100), (44100, 50000)]
PRESETS = [G.QualityLow, G.QualityMedium, G.QualityHigh, G.QualityVeryHigh]


def run(ir, orr, preset, x, cuts, mode):
    G.set_chain_kernel(mode)
    try:
        dev = torch.device("cuda", 0)
        rows = x.shape[0]
        h = G.NewBatch(ir, orr, preset, rows, np.float64)
        ys = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            dx = torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).to(dev)
            n = hi - lo
            cap = (h.EstimateOutput(n) + 64 + 3) & ~3
            dy = torch.zeros((rows, cap), dtype=torch.float64, device=dev)
            k = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), cap, cap, 0, np.float64)
            torch.cuda.synchronize()
            ys.append(dy[:, :k].cpu().numpy())
        ys.append(h.FlushBatch()[0].copy())
        return ys, h.last_kernels()
    finally:
        G.set_chain_kernel(2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    taken = 0
    for case in range(a.cases):
        ir, orr = PAIRS[int(rng.integers(len(PAIRS)))]
        preset = PRESETS[int(rng.integers(len(PRESETS)))]
        rows = int(rng.choice([32, 33, 40, 47, 64, 70, 96, 130, 200, 256, 300]))
        n = int(rng.integers(2_200_000 // rows + 2000, 6_000_000 // rows + 4000))
        k = int(rng.integers(0, 3))
        cuts = sorted(set([0, n] + [int(v) for v in rng.integers(0, n, size=k)]))
        x = rng.standard_normal((rows, n))
        ya, ka = run(ir, orr, preset, x, cuts, 1)
        yb, kb = run(ir, orr, preset, x, cuts, 0)
        used = "chain_up2_poly_f64_mma" in ka
        taken += used
        assert "chain_up2_poly_f64_mma" not in kb
        for p, q in zip(ya, yb):
            assert p.shape == q.shape, (case, ir, orr, preset, rows, n, cuts, p.shape, q.shape)
            assert np.array_equal(p, q), (case, ir, orr, preset, rows, n, cuts, float(np.max(np.abs(p - q))))
        print(f"case {case:3d} {ir:6d}->{orr:6d} q{preset} rows {rows:3d} n {n:7d} calls {len(cuts)-1} chain kernel "
              f"{'taken' if used else 'not taken'}: bit-identical", flush=True)
    print(f"ok: {a.cases} cases, chain kernel taken in {taken}, every sample bit-identical to the stand-alone launches")
    assert taken >= a.cases // 3


if __name__ == "__main__":
    main()
