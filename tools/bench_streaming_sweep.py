#!/usr/bin/env python
"""Host-call time per 4096-frame chunk (ProcessFloat32Into on a float64 pipeline, path A; Process on path-B float64 engines)
over many rate pairs: a sweep to spot streaming-size dispatch corners (a pair far above the ~35-45 us of a one-launch chunk)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

PAIRS = [(48000, 44100), (44100, 48000), (48000, 16000), (16000, 48000), (8000, 44100), (96000, 44100), (44100, 22050),
         (22050, 44100), (44100, 47999), (48000, 8000), (8000, 48000), (8000, 192000), (192000, 44100)]
chunk = 4096
rng = np.random.default_rng(0)
for (ir, orr) in PAIRS:
    x = (0.5 * rng.standard_normal(chunk * 120)).astype(np.float32)
    res = []
    for path in ("A", "B"):
        if path == "A":
            r = G.New(G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=G.QualityHigh)))
            out = np.empty(r.EstimateOutput(chunk), dtype=np.float32)
            call = lambda c: r.ProcessFloat32Into(c, out)  # noqa: E731
            xs = x
        else:
            r = G.NewEngine(ir, orr, G.QualityHigh)
            xs = x.astype(np.float64)
            call = lambda c: r.Process(c)  # noqa: E731
        for i in range(0, 20 * chunk, chunk):
            call(xs[i:i + chunk])
        G.kernel_launches(reset=True)
        t0 = time.perf_counter()
        for i in range(20 * chunk, 120 * chunk, chunk):
            call(xs[i:i + chunk])
        dt = (time.perf_counter() - t0) / 100
        res.append(f"path {path}: {dt * 1e6:6.1f} us/chunk, {G.kernel_launches() / 100:.1f} launches {r.last_kernels()}")
    print(f"{ir:6d}->{orr:6d}  " + "  |  ".join(res), flush=True)
