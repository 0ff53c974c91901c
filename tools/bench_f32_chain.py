#!/usr/bin/env python
"""float32 engine batches at non-integer ratios (NewEngineFloat32, convenience.go:329-366): device time of a Process+Flush
pass of 256 lock-step rows x 10 s, next to the float64 engine on the same signal."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

dev = torch.device("cuda", 0)
ts = torch.cuda.Stream(device=dev)
for (ir, orr, dt, rows, n) in [(44100, 48000, np.float32, 256, 441000), (44100, 48000, np.float64, 256, 441000),
                               (48000, 44100, np.float32, 256, 480000), (44100, 47999, np.float32, 256, 441000),
                               (44100, 48000, np.float32, 16, 441000)]:
    h = G.NewBatch(ir, orr, G.QualityHigh, rows, dt)
    x = np.random.default_rng(0).standard_normal((rows, n)).astype(dt)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    esz = 4 if dt == np.float32 else 8
    dx = torch.from_numpy(x).to(dev)
    ostride = (h.EstimateOutput(n) + 8192 + 3) & ~3
    dy = torch.zeros((rows, ostride), dtype=tdt, device=dev)

    def one():
        h.Reset()
        n1 = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), ostride, ostride, ts.cuda_stream, dt)
        n2 = h.flush_batch_dev(dy.data_ptr() + n1 * esz, ostride, ostride - n1, ts.cuda_stream, dt)
        return n1 + n2

    for _ in range(2):
        no = one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(3):
        one()
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{ir}->{orr} {np.dtype(dt).name} rows {rows}: {ms:.3f} ms, {rows * no / ms / 1e6:.2f} G out-samples/s, kernels {h.last_kernels()}")
