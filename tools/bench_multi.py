#!/usr/bin/env python
"""End-to-end throughput of BASELINE config 4 through ONE multi-device handle in ONE process (gar_create_multi): the
caller's pinned host buffers are sharded by rows over the devices inside gar_process_batch / gar_flush_batch — what a
single-process Go caller gets, next to bench.py's one-process-per-GPU numbers.

    python tools/bench_multi.py --devices 0,1,2,3,4,5,6,7 [--streams 4096] [--seconds 10] [--steps 5]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="")
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--wc", action="store_true", help="write-combined pinned input buffer")
    a = ap.parse_args()
    devs = [int(d) for d in a.devices.split(",")] if a.devices else list(range(G.device_count()))
    n_in = int(a.seconds * 48000)
    h = G.NewBatch(48000, 16000, G.QualityMedium, a.streams, np.float32, devices=devs)
    ostride = (h.EstimateOutput(n_in) + 3) & ~3
    if a.wc:
        xh, px = G.host_alloc_wc((a.streams, n_in), np.float32)
    else:
        xh, px = h.host_alloc_rows(n_in, np.float32)   # pages first-touched by the shards' NUMA-bound workers
    yh, py = h.host_alloc_rows(ostride, np.float32)
    t = np.arange(n_in, dtype=np.float64) / 48000.0
    rng = np.random.default_rng(7)
    base = (np.sin(2 * np.pi * 440.0 * t) + 0.05 * (2 * rng.random(n_in) - 1)).astype(np.float32)
    for r in range(a.streams):
        xh[r] = np.roll(base, 37 * r)

    def step():
        h.Reset()
        _, m1 = h.ProcessBatch(xh, yh)
        _, m2 = h.FlushBatch(yh[:, m1:])
        return m1 + m2

    for _ in range(2):
        n = step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = (time.perf_counter() - t0) / a.steps
    # rows are independent: row r of the multi-device run equals the same signal through a single-device handle
    one = G.NewBatch(48000, 16000, G.QualityMedium, 2, np.float32, device=devs[-1])
    rows = [0, a.streams - 1]
    ref = np.concatenate([one.ProcessBatch(np.ascontiguousarray(xh[rows]))[0], one.FlushBatch()[0]], axis=1)
    err = float(np.max(np.abs(yh[rows, :n].astype(np.float64) - ref.astype(np.float64))))
    print(json.dumps({"metric": "output_msamples_per_s", "e2e_value": round(a.streams * n / dt / 1e6, 1), "unit": "Msamples/s",
                      "ms_per_step": round(dt * 1e3, 2), "devices": devs, "shards": h.shards(), "streams": a.streams,
                      "samples_per_stream": n_in, "h2d_bytes_per_step": a.streams * n_in * 4,
                      "d2h_bytes_per_step": a.streams * n * 4, "api": "one gar_create_multi handle, one process: "
                      "gar_process_batch + gar_flush_batch on sharded pinned buffers (gar_host_alloc_rows)",
                      "write_combined_input": a.wc,
                      "rows_vs_single_device_max_abs_diff": err}))
    G.host_free(px)
    G.host_free(py)


if __name__ == "__main__":
    main()
