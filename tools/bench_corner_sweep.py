#!/usr/bin/env python
"""Device time of a Process+Flush pass of ONE 10 s stream through path-A pipelines (New(Config)) and path-B engines over many rate
pairs and presets: a sweep to spot dispatch corners (a pair whose output rate in G samples/s is far below its neighbours').

    python tools/bench_corner_sweep.py [rows]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1
BDT = np.float32 if len(sys.argv) > 2 and sys.argv[2] == "f32" else np.float64  # dtype of the path-B engines
PAIRS = [(44100, 48000), (48000, 44100), (48000, 16000), (16000, 48000), (8000, 44100), (96000, 44100), (44100, 22050),
         (22050, 44100), (32000, 48000), (48000, 32000), (11025, 48000), (192000, 44100), (48000, 8000), (8000, 16000)]
PRESETS = [("low", G.QualityLow), ("medium", G.QualityMedium), ("high", G.QualityHigh), ("veryhigh", G.QualityVeryHigh)]
dev = torch.device("cuda", 0)
ts = torch.cuda.Stream(device=dev)
for path in (("B",) if BDT == np.float32 else ("A", "B")):
    for (ir, orr) in PAIRS:
        line = []
        for pname, preset in PRESETS:
            n = 10 * ir
            if path == "A":
                h = G.Resampler(G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=preset)), n_streams=rows)
            else:
                h = G.NewBatch(ir, orr, preset, rows, BDT)
            dt = BDT if path == "B" else np.float64
            x = np.random.default_rng(0).standard_normal((rows, n)).astype(dt)
            dx = torch.from_numpy(x).to(dev)
            ostride = (h.EstimateOutput(n) + 8192 + 3) & ~3
            esz = 4 if dt == np.float32 else 8
            dy = torch.zeros((rows, ostride), dtype=torch.float32 if dt == np.float32 else torch.float64, device=dev)

            def one():
                h.Reset()
                n1 = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), ostride, ostride, ts.cuda_stream, dt)
                n2 = h.flush_batch_dev(dy.data_ptr() + n1 * esz, ostride, ostride - n1, ts.cuda_stream, dt)
                return n1 + n2

            for _ in range(3):
                no = one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(10):
                one()
            e1.record(ts)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 10 * 1e3
            line.append(f"{pname} {us:7.1f} us")
            kern = h.last_kernels()
        print(f"path {path} {ir:6d}->{orr:6d} rows {rows}: " + " | ".join(line) + f"  {kern}", flush=True)
