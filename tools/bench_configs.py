#!/usr/bin/env python
"""Per-config measurement of the five BASELINE.json configs on one B200 (not the driver's bench line).

For every config: device-resident time of one full Process+Flush pass (CUDA events on the launching stream,
Reset between repetitions), output Msamples/s, achieved FMA TFLOP/s against the dependent-FMA probe of the
compute type (algorithmic flops per output from SURVEY.md §8d), the host-facing call path (H2D + kernels +
D2H through the C ABI, wall clock) and the CPU oracle on one host thread for the same input.

    python tools/bench_configs.py [--reps 20] [--json profiles/rNN_configs.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import gar_b200 as G  # noqa: E402
from helpers import sig_c1, sig_c2, sig_c3, sig_c4, sig_c5a  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cfg(ir, orr, ch, preset):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset))


def configs():
    c2l, c2r = sig_c2()
    return [
        dict(name="C1 ResampleMono 44.1k->48k High f64", make=lambda: G.NewEngine(44100, 48000, G.QualityHigh),
             x=sig_c1()[None, :], io=np.float64, flops=738.0, compute="f64",
             cpu=lambda x: O.resample_mono(x[0], 44100, 48000, O.PRESET_HIGH)),
        dict(name="C2 stereo 48k->44.1k High f32 io (f64 inside), one shot", io=np.float32, flops=980.8, compute="f64",
             make=lambda: G.Resampler(cfg(48000, 44100, 1, G.QualityHigh), n_streams=2),
             x=np.stack([c2l, c2r]), cpu=None),
        dict(name="C3 7.1 (8 ch) 96k->48k VeryHigh f64", make=lambda: G.New(cfg(96000, 48000, 8, G.QualityVeryHigh)),
             x=np.stack(sig_c3()), io=np.float64, flops=2446.0, compute="f64", cpu=None),
        dict(name="C4 256 of the 4096 mono streams 48k->16k Medium f32", io=np.float32, flops=1754.0, compute="f32",
             make=lambda: G.NewBatch(48000, 16000, G.QualityMedium, 256, np.float32), x=sig_c4(256, 480000), cpu=None),
        dict(name="K4 roofline: C1 chain batched x256 streams (44.1k->48k High f64, 10 s each)",
             make=lambda: G.NewBatch(44100, 48000, G.QualityHigh, 256, np.float64),
             x=np.tile(sig_c1(441000)[None, :], (256, 1)), io=np.float64, flops=738.0, compute="f64", cpu=None),
        dict(name="K4 roofline: C2 chain batched x256 streams (48k->44.1k 24Bit f64, 10 s each)",
             make=lambda: G.Resampler(cfg(48000, 44100, 1, G.QualityHigh), n_streams=256),
             x=np.tile(sig_c1(480000)[None, :], (256, 1)), io=np.float64, flops=980.8, compute="f64", cpu=None),
        dict(name="K4 roofline: C5b chain batched x256 streams (44.1k->47.999k High f64, cubic coefficient interpolation, 10 s each)",
             make=lambda: G.NewBatch(44100, 47999, G.QualityHigh, 256, np.float64),
             x=np.tile(sig_c1(441000)[None, :], (256, 1)), io=np.float64, flops=692.0, compute="f64", cpu=None),
        dict(name="K4 roofline: C5a chain batched x64 streams (8k->192k High f64, 10 s each)",
             make=lambda: G.Resampler(cfg(8000, 192000, 1, G.QualityHigh), n_streams=64),
             x=np.tile(sig_c5a()[None, :], (64, 1)), io=np.float64, flops=1233.3, compute="f64", cpu=None),
        dict(name="C5a 8k->192k High multistage f64", make=lambda: G.New(cfg(8000, 192000, 1, G.QualityHigh)),
             x=sig_c5a()[None, :], io=np.float64, flops=1233.3, compute="f64", cpu=None),
        dict(name="C5b 44.1k->47.999k High f64 (cubic coefficient interpolation)", io=np.float64, flops=692.0,
             make=lambda: G.NewEngine(44100, 47999, G.QualityHigh), x=sig_c1()[None, :], compute="f64",
             cpu=lambda x: O.resample_mono(x[0], 44100, 47999, O.PRESET_HIGH)),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks = {"f32": G.measure_fma_peak(np.float32), "f64": G.measure_fma_peak(np.float64)}
    print(f"FMA probe: fp32 {peaks['f32']:.1f} TFLOP/s, fp64 {peaks['f64']:.1f} TFLOP/s")
    ts = torch.cuda.Stream(device=dev)
    rows_out = []
    for c in configs():
        if a.only and a.only not in c["name"]:
            continue
        h = c["make"]()
        x = np.ascontiguousarray(c["x"], dtype=c["io"])
        rows, n_in = x.shape
        assert rows == h.rows, (rows, h.rows)
        tdt = torch.float32 if c["io"] == np.float32 else torch.float64
        esz = 4 if c["io"] == np.float32 else 8
        dx = torch.from_numpy(x).to(dev)
        est = h.EstimateOutput(n_in)
        ostride = (est + 8192 + 3) & ~3
        dy = torch.zeros((rows, ostride), dtype=tdt, device=dev)
        torch.cuda.synchronize()

        def one_pass():
            h.Reset()
            n1 = h.process_batch_dev(dx.data_ptr(), n_in, n_in, dy.data_ptr(), ostride, ostride, ts.cuda_stream, c["io"])
            n2 = h.flush_batch_dev(dy.data_ptr() + n1 * esz, ostride, ostride - n1, ts.cuda_stream, c["io"])
            return n1, n2

        for _ in range(3):
            n1, n2 = one_pass()
        torch.cuda.synchronize()
        G.kernel_launches(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(a.reps):
            one_pass()
        e1.record(ts)
        torch.cuda.synchronize()
        launches = G.kernel_launches() / a.reps
        ms = e0.elapsed_time(e1) / a.reps
        n_out = rows * (n1 + n2)
        # host-facing path
        xh = x
        yh = np.empty((rows, ostride), dtype=c["io"])
        for _ in range(2):
            h.Reset()
            _, m1 = h.ProcessBatch(xh, yh)
            h.FlushBatch(yh[:, m1:])
        t0 = time.perf_counter()
        nh = max(3, a.reps // 4)
        for _ in range(nh):
            h.Reset()
            _, m1 = h.ProcessBatch(xh, yh)
            h.FlushBatch(yh[:, m1:])
        host_ms = (time.perf_counter() - t0) / nh * 1e3
        cpu_ms = None
        if c["cpu"] is not None:
            t0 = time.perf_counter()
            c["cpu"](x)
            cpu_ms = (time.perf_counter() - t0) * 1e3
        tf = n_out * c["flops"] / (ms * 1e-3) / 1e12
        rec = dict(config=c["name"], rows=rows, n_in=n_in, n_out_per_row=n1 + n2, device_ms=round(ms, 4),
                   msamples_per_s=round(n_out / ms / 1e3, 1), tflops=round(tf, 3),
                   fma_frac=round(tf / peaks[c["compute"]], 4), compute=c["compute"], kernels=h.kernel_names(),
                   launches_per_pass=launches, host_call_ms=round(host_ms, 3),
                   host_msamples_per_s=round(n_out / host_ms / 1e3, 1), cpu_1thread_ms=cpu_ms and round(cpu_ms, 1))
        rows_out.append(rec)
        print(json.dumps(rec))
        del h, dx, dy
    if a.json:
        Path(a.json).write_text(json.dumps({"fma_peaks_tflops": peaks, "configs": rows_out}, indent=1))


if __name__ == "__main__":
    main()
