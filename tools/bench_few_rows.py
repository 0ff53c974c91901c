#!/usr/bin/env python
"""Device time of a Process+Flush pass of 10 s streams for 1 .. 24 lock-step rows, float32 and float64 engines, several ratios:
a sweep to spot dispatch corners (a row count or dtype that is much slower per row than its neighbours).

    python tools/bench_few_rows.py [44100:48000 48000:44100 ...]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(44100, 48000), (48000, 44100)]
dev = torch.device("cuda", 0)
ts = torch.cuda.Stream(device=dev)
for (ir, orr) in pairs:
    for rows in (1, 2, 4, 8, 16, 24):
        for dt in (np.float32, np.float64):
            n = 10 * ir
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

            for _ in range(3):
                no = one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(20):
                one()
            e1.record(ts)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print(f"{ir}->{orr} rows {rows:2d} {np.dtype(dt).name}: {us:8.1f} us {us / rows:7.1f} us/row "
                  f"{rows * no / us / 1e3:6.2f} G samples/s  {h.last_kernels()}", flush=True)
