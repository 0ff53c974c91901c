#!/usr/bin/env python
"""A/B timing of the persistent chain kernel K5 (one launch per Process, intermediate samples in an L2-resident ring) against
the two stand-alone tensor-core launches (K1m + K3p, full-size intermediate buffer through HBM) on the batched BASELINE chains.

    python tools/bench_chain.py [--reps 10] [--rows 256] [--json out.json]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import gar_b200 as G  # noqa: E402

CHAINS = [("C1 44.1k->48k", 44100, 48000, 441000, 738.0), ("C2 48k->44.1k", 48000, 44100, 480000, 980.8),
          ("C5b 44.1k->47.999k", 44100, 47999, 441000, 692.0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=256)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="")
    ap.add_argument("--check", action="store_true", help="compare the outputs of the two paths bit by bit")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dmma = G.measure_fma_peak(np.float64, tensor=True)
    ts = torch.cuda.Stream(device=dev)
    out = []
    rng = np.random.default_rng(1)
    for name, ir, orr, n_in, flops in CHAINS:
        if a.only and a.only not in name:
            continue
        x = 0.5 * rng.standard_normal((a.rows, n_in))
        dx = torch.from_numpy(x).to(dev)
        rec = {"chain": name, "rows": a.rows, "n_in": n_in}
        ys = {}
        for chain in (True, False):
            G.set_chain_kernel(chain)
            h = G.NewBatch(ir, orr, G.QualityHigh, a.rows, np.float64)
            est = h.EstimateOutput(n_in)
            ostride = (est + 8192 + 3) & ~3
            dy = torch.zeros((a.rows, ostride), dtype=torch.float64, device=dev)

            def one_pass():
                h.Reset()
                n1 = h.process_batch_dev(dx.data_ptr(), n_in, n_in, dy.data_ptr(), ostride, ostride, ts.cuda_stream, np.float64)
                n2 = h.flush_batch_dev(dy.data_ptr() + n1 * 8, ostride, ostride - n1, ts.cuda_stream, np.float64)
                return n1, n2

            for _ in range(2):
                n1, n2 = one_pass()
            torch.cuda.synchronize()
            G.kernel_launches(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(a.reps):
                one_pass()
            e1.record(ts)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            tf = a.rows * (n1 + n2) * flops / (ms * 1e-3) / 1e12
            key = "chain_kernel" if chain else "two_launches"
            rec[key] = {"device_ms": round(ms, 4), "tflops": round(tf, 2), "frac_of_dmma_probe": round(tf / dmma, 4),
                        "launches_per_pass": G.kernel_launches() / a.reps, "kernels": h.last_kernels(),
                        "device_bytes": h.device_bytes() if hasattr(h, "device_bytes") else None}
            if a.check:
                ys[chain] = dy[:, :n1 + n2].cpu().numpy()
            del h, dy
        G.set_chain_kernel(True)
        if a.check:
            rec["bit_identical"] = bool(np.array_equal(ys[True], ys[False]))
            rec["max_abs_diff"] = float(np.max(np.abs(ys[True] - ys[False])))
        rec["speedup"] = round(rec["two_launches"]["device_ms"] / rec["chain_kernel"]["device_ms"], 4)
        out.append(rec)
        print(json.dumps(rec))
    if a.json:
        Path(a.json).write_text(json.dumps({"dmma_probe_tflops": dmma, "chains": out}, indent=1))


if __name__ == "__main__":
    main()
