#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 resampling engine (contract: see README / DESIGN.md §6).

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): 4096 independent mono
streams, 48 kHz -> 16 kHz, float32 end to end, NewEngineFloat32(48000,16000,QualityMedium) semantics
(877-tap /3 decimating FIR), 480 000 samples (10 s) per stream, one batched Process + Flush per step.
The 4096 streams are sharded across the N ranks (strong scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
sys.path.insert(0, str(ROOT))

IN_RATE, OUT_RATE = 48000, 16000
PRESETS = {"low": 1, "medium": 2, "high": 3}
TAPS = {"low": 245, "medium": 877, "high": 1125}  # dft_stage.go:401-475 via the oracle / product design
METRIC = "output_msamples_per_s"
UNIT = "Msamples/s"


# Exactly ONE line may reach stdout (the driver parses it). Libraries print there too (NCCL's version banner, warnings), so
# file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to a saved copy of the real stdout.
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="total streams over all ranks")
    ap.add_argument("--seconds", type=float, default=10.0, help="audio seconds per stream")
    ap.add_argument("--preset", default="medium", choices=list(PRESETS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the float64 side measurements (BASELINE configs 1 and 3)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node (A/B)")
    ap.add_argument("--no-ceiling", action="store_true", help="skip the in-job copy-ceiling measurement")
    ap.add_argument("--parity-rows", type=int, default=16, help="rows of the timed run checked against the oracle")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.streams} mono streams x {int(a.seconds * IN_RATE)} samples, 48k->16k float32, "
            f"NewEngineFloat32 Quality{a.preset.capitalize()} ({TAPS[a.preset]}-tap /3 FIR), Process+Flush")


def synth_rows_host(first, count, n, dtype=np.float32):
    t = np.arange(n, dtype=np.float64) / IN_RATE
    x = np.empty((count, n), dtype=dtype)
    for i in range(count):
        s = first + i
        rng = np.random.default_rng(s)
        x[i] = (np.sin(2 * np.pi * (200 + 1.7 * s) * t) + 0.05 * (2 * rng.random(n) - 1)).astype(dtype)
    return x


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax = max(smax, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = upper half of the samples (idle samples before/after the region are dropped)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(a, x_host, n_threads, target_s):
    """The oracle (C++ restatement of the Go path, AVX2+FMA) on the host cores, bounded sample."""
    from oracle import oracle as O
    q = O.preset_to_engine_quality(PRESETS[a.preset])
    n = x_host.shape[1]
    probe = min(x_host.shape[0], max(n_threads, 4))
    t0 = time.perf_counter()
    _, counts = O.batch_resample(x_host[:probe], IN_RATE, OUT_RATE, q, n_threads=n_threads)
    dt = time.perf_counter() - t0
    ns = int(min(x_host.shape[0], max(probe, probe * target_s / max(dt, 1e-3))))
    ns = max(n_threads, (ns // n_threads) * n_threads)
    ns = min(ns, x_host.shape[0])
    t0 = time.perf_counter()
    _, counts = O.batch_resample(x_host[:ns], IN_RATE, OUT_RATE, q, n_threads=n_threads)
    dt = time.perf_counter() - t0
    v_all = float(counts.sum()) / dt / 1e6
    # single thread on a smaller sample
    n1 = max(1, min(ns, int(max(1, ns / n_threads / 2))))
    t0 = time.perf_counter()
    _, c1 = O.batch_resample(x_host[:n1], IN_RATE, OUT_RATE, q, n_threads=1)
    d1 = time.perf_counter() - t0
    return {"value": round(v_all, 3), "unit": UNIT, "cores": n_threads, "kind": "port",
            "sample": f"{ns} of the streams x {n} samples (Process+Flush), {dt:.2f} s wall; "
                      "C++ restatement of the Go path (AVX2+FMA), not the Go binary: no Go toolchain in the image",
            "value_1thread": round(float(c1.sum()) / d1 / 1e6, 3), "avx2": bool(O.lib().orc_has_avx2())}


def run_reference(a, rank):
    """--impl reference: the reference's CPU path (oracle port; Go cannot be built here) on all host cores."""
    if rank != 0:
        return
    n = int(a.seconds * IN_RATE)
    cores = os.cpu_count() or 1
    ns = min(a.streams, cores * 4)
    x = synth_rows_host(0, ns, n)
    from oracle import oracle as O
    q = O.preset_to_engine_quality(PRESETS[a.preset])
    times, outs = [], 0
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        _, counts = O.batch_resample(x, IN_RATE, OUT_RATE, q, n_threads=cores)
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            times.append(dt)
            outs = int(counts.sum())
    ms = 1e3 * sum(times) / len(times)
    val = outs / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "step": f"bounded sample: {ns} streams per step on {cores} host threads",
                       "reference_sample_streams": ns, "workload_streams": a.streams,
                       "note": "throughput metric: the bounded sample runs the same per-stream work as the full workload"},
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{ns} streams x {n} samples per step; C++ restatement of the Go path "
                                       "(oracle/, AVX2+FMA) — the Go reference cannot be built in this image"},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def extra_f64(dev, local, tstream):
    """Side measurements, not the headline: the other BASELINE configs on this GPU, so that their numbers are driver-observed.
    Per config: device time of one Process+Flush pass (CUDA events on the launching stream, inputs resident in HBM, Reset
    between passes), the host-facing call (numpy host buffers through the C ABI, wall clock), the CPU oracle on ONE host
    thread on the same input, and the achieved FMA rate against BOTH float64 probes of this job: vector DFMA and the FP64
    tensor cores (DMMA.8x8x4) — kernels that run on the tensor cores are quoted against the DMMA probe."""
    import torch

    import gar_b200 as G
    from oracle import oracle as O

    peak64 = G.measure_fma_peak(np.float64, local)
    peak_dmma = G.measure_fma_peak(np.float64, local, tensor=True)
    out = {"fp64_fma_peak_tflops": round(peak64, 2), "fp64_dmma_peak_tflops": round(peak_dmma, 2)}

    def timed(h, x, flops_per_out, reps=5, io=np.float64, host=True):
        rows, n_in = x.shape
        tdt, esz = (torch.float32, 4) if io == np.float32 else (torch.float64, 8)
        est = h.EstimateOutput(n_in)
        ostride = (est + 8192 + 3) & ~3
        dx = torch.from_numpy(np.ascontiguousarray(x, dtype=io)).to(dev)
        dy = torch.zeros((rows, ostride), dtype=tdt, device=dev)

        def one():
            h.Reset()
            n1 = h.process_batch_dev(dx.data_ptr(), n_in, n_in, dy.data_ptr(), ostride, ostride, tstream.cuda_stream, io)
            n2 = h.flush_batch_dev(dy.data_ptr() + n1 * esz, ostride, ostride - n1, tstream.cuda_stream, io)
            return n1 + n2

        for _ in range(3):
            n = one()
        torch.cuda.synchronize()
        G.kernel_launches(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstream)
        for _ in range(reps):
            one()
        e1.record(tstream)
        torch.cuda.synchronize()
        launches = G.kernel_launches() / reps
        ms = e0.elapsed_time(e1) / reps
        tf = rows * n * flops_per_out / (ms * 1e-3) / 1e12
        kernels = h.last_kernels()
        tensor = any("mma" in k for k in kernels)
        r = {"device_us": round(ms * 1e3, 2), "launches_per_pass": launches, "msamples_per_s": round(rows * n / ms / 1e3, 1),
             "tflops": round(tf, 2), "frac_of_fp64_fma_peak": round(tf / peak64, 4),
             "frac_of_fp64_dmma_peak": round(tf / peak_dmma, 4), "frac": round(tf / (peak_dmma if tensor else peak64), 4),
             "frac_denominator": "dmma probe" if tensor else "vector dfma probe", "kernels": kernels}
        if host:
            xh = np.ascontiguousarray(x, dtype=io)
            yh = np.empty((rows, ostride), dtype=io)
            for _ in range(2):
                h.Reset()
                _, m1 = h.ProcessBatch(xh, yh)
                h.FlushBatch(yh[:, m1:])
            t0 = time.perf_counter()
            for _ in range(5):
                h.Reset()
                _, m1 = h.ProcessBatch(xh, yh)
                h.FlushBatch(yh[:, m1:])
            r["host_call_us"] = round((time.perf_counter() - t0) / 5 * 1e6, 1)
            # the same calls from PINNED caller buffers (gar_host_alloc): pageable numpy arrays above go through the driver's
            # staging copy at ~12 GB/s, pinned ones at the PCIe rate
            xp, pxp = G.host_alloc(xh.shape, io)
            yp, pyp = G.host_alloc(yh.shape, io)
            xp[...] = xh
            for _ in range(2):
                h.Reset()
                _, m1 = h.ProcessBatch(xp, yp)
                h.FlushBatch(yp[:, m1:])
            t0 = time.perf_counter()
            for _ in range(5):
                h.Reset()
                _, m1 = h.ProcessBatch(xp, yp)
                h.FlushBatch(yp[:, m1:])
            r["host_call_pinned_us"] = round((time.perf_counter() - t0) / 5 * 1e6, 1)
            G.host_free(pxp)
            G.host_free(pyp)
        return r

    def cpu_ms(fn):
        t0 = time.perf_counter()
        fn()
        return round((time.perf_counter() - t0) * 1e3, 1)

    def cfg(ir, orr, ch, preset):
        return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset))

    t1 = np.arange(441000) / 44100.0
    x1 = np.sin(2 * np.pi * 1000.0 * t1)
    # C1: ResampleMono 44.1k -> 48k QualityHigh float64, one stream
    out["c1_single_stream"] = dict(
        timed(G.NewEngine(44100, 48000, G.QualityHigh), x1[None, :], 738.0, reps=20),
        cpu_1thread_ms=cpu_ms(lambda: O.resample_mono(x1, 44100, 48000, O.PRESET_HIGH)),
        workload="BASELINE config 1: ResampleMono 44.1k->48k QualityHigh float64, 441000 samples")
    # C5b: 44.1k -> 47.999k (cubic coefficient interpolation), one stream
    out["c5b_single_stream"] = dict(
        timed(G.NewEngine(44100, 47999, G.QualityHigh), x1[None, :], 692.0, reps=20),
        cpu_1thread_ms=cpu_ms(lambda: O.resample_mono(x1, 44100, 47999, O.PRESET_HIGH)),
        workload="BASELINE config 5b: 44.1k->47.999k QualityHigh float64, 441000 samples")
    # C5a: 8k -> 192k multistage, one stream
    x5 = np.sin(2 * np.pi * 1000.0 * np.arange(80000) / 8000.0)

    def c5a_cpu():
        p = O.Pipeline(8000, 192000, 1, O.PRESET_HIGH)
        p.process(x5)
        p.flush()
    out["c5a_single_stream"] = dict(
        timed(G.New(cfg(8000, 192000, 1, G.QualityHigh)), x5[None, :], 1233.3, reps=20), cpu_1thread_ms=cpu_ms(c5a_cpu),
        workload="BASELINE config 5a: 8k->192k QualityHigh float64 (five x2 stages + polyphase), 80000 samples")
    # C2: stereo 48k -> 44.1k High, float32 I/O, 4096-frame chunks + Flush: two mono instances (constant.go), per chunk
    tc = np.arange(480000) / 48000.0
    left = (np.sin(2 * np.pi * 440 * tc) + 0.1 * np.sin(2 * np.pi * 1320 * tc)).astype(np.float32)
    h2 = G.New(cfg(48000, 44100, 1, G.QualityHigh))
    chunks = [left[i:i + 4096] for i in range(0, len(left), 4096)]
    obuf = np.empty(h2.EstimateOutput(4096), dtype=np.float32)
    for c in chunks[:8]:
        h2.ProcessFloat32Into(c, obuf)
    h2.Reset()
    t0 = time.perf_counter()
    for c in chunks:
        h2.ProcessFloat32Into(c, obuf)
    host_chunk_us = (time.perf_counter() - t0) / len(chunks) * 1e6
    h2.Flush()
    dl = torch.from_numpy(left).to(dev)
    dob = torch.zeros(8192, dtype=torch.float32, device=dev)
    h2.Reset()
    for k in range(8):
        h2.process_batch_dev(dl.data_ptr() + k * 4096 * 4, 4096, 4096, dob.data_ptr(), 8192, 8192, tstream.cuda_stream, np.float32)
    torch.cuda.synchronize()
    h2.Reset()
    G.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    nfull = len(left) // 4096
    for k in range(nfull):
        h2.process_batch_dev(dl.data_ptr() + k * 4096 * 4, 4096, 4096, dob.data_ptr(), 8192, 8192, tstream.cuda_stream, np.float32)
    e1.record(tstream)
    torch.cuda.synchronize()

    def c2_cpu():
        p = O.Pipeline(48000, 44100, 1, O.PRESET_HIGH)
        ob = np.empty(p.estimate_output(4096), dtype=np.float32)
        for c in chunks:
            p.process_f32_into(c, ob)
        p.flush()
    c2cpu = cpu_ms(c2_cpu)
    out["c2_streaming_chunk"] = {
        "device_us_per_chunk": round(e0.elapsed_time(e1) / nfull * 1e3, 2), "launches_per_chunk": G.kernel_launches() / nfull,
        "host_call_us_per_chunk": round(host_chunk_us, 1), "cpu_1thread_us_per_chunk": round(c2cpu * 1e3 / len(chunks), 1),
        "kernels": h2.last_kernels(),
        "workload": "BASELINE config 2, one channel (the reference runs two mono instances): 48k->44.1k QualityHigh, float32 "
                    "I/O through ProcessFloat32Into in 4096-frame chunks (117 chunks + tail), float64 inside"}
    del h2, dl, dob
    # C1 chain batched over 256 lock-step streams (tensor-core kernels)
    x256 = np.tile(x1[None, :], (256, 1))
    out["c1_chain_x256_streams_10s_f64"] = dict(
        timed(G.NewBatch(44100, 48000, G.QualityHigh, 256, np.float64), x256, 738.0, host=False),
        workload="BASELINE config 1's chain (44.1k->48k QualityHigh float64) x 256 lock-step streams x 10 s; default path: "
                 "two tensor-core launches (K1m + K3p), full-size intermediate buffer")
    # the same call through the persistent chain kernel K5: one launch per Process, intermediate samples in an L2-resident
    # ring (no 1.8 GB buffer); DRAM traffic of both paths measured with ncu: profiles/r2_chain_traffic.txt
    G.set_chain_kernel(1)
    try:
        k5 = timed(G.NewBatch(44100, 48000, G.QualityHigh, 256, np.float64), x256, 738.0, host=False)
    finally:
        G.set_chain_kernel(2)
    out["c1_chain_x256_streams_10s_f64_k5"] = dict(
        k5, workload="the same call with gar_set_chain_kernel(1): ONE persistent launch per Process (chain_up2_poly_f64_mma)",
        dram_bytes_per_pass_ncu={"k5": 2.17e9, "two_launches": 5.53e9, "algorithmic": 1.886e9,
                                 "source": "profiles/r2_chain_traffic.txt"})
    del x256
    # C3: 8 channels 96k -> 48k VeryHigh (1223-tap /2)
    rng = np.random.default_rng(4242)
    t3 = np.arange(960000) / 96000.0
    x3 = np.stack([0.7 * np.sin(2 * np.pi * 440 * t3 + c) + 0.2 * np.sin(2 * np.pi * 1750 * t3) + 0.1 * (rng.random(t3.size) - 0.5)
                   for c in range(8)])

    def c3_cpu():
        p = O.Pipeline(96000, 48000, 1, O.PRESET_VERYHIGH)
        p.process(x3[0])
        p.flush()
    out["c3_8ch_96k_to_48k_veryhigh_f64"] = dict(
        timed(G.New(cfg(96000, 48000, 8, G.QualityVeryHigh)), x3, 2446.0, reps=10),
        cpu_1thread_ms_per_channel=cpu_ms(c3_cpu),
        workload="BASELINE config 3: 8 channels x 960000 samples, 96k->48k QualityVeryHigh float64 (1223-tap /2)")
    return out


def main():
    _capture_stdout()
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    import gar_b200 as G

    affinity0 = os.sched_getaffinity(0)
    # One process per GPU: bind this rank to the cores of its GPU's NUMA node BEFORE torch spawns its threads and before any
    # pinned host memory is allocated (first touch places the pages). SCALE_r01's e2e numbers were taken unbound.
    numa = {"bound_node": None if a.no_bind else G.bind_thread_to_device(local), "device_node": G.device_numa_node(local),
            "cpus_allowed": len(os.sched_getaffinity(0))}

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from gar_b200.shard import partition
    n_in = int(a.seconds * IN_RATE)
    first, rows = partition(a.streams, world, rank)  # independent streams: shard by rows, no collective
    taps = TAPS[a.preset]
    flops_per_out = 2.0 * taps  # 2 x MACs per output sample

    b = G.NewBatch(IN_RATE, OUT_RATE, PRESETS[a.preset], rows, np.float32, device=local)
    assert b.describe()[0]["taps"] == taps
    kname = b.kernel_names()[0]
    est = b.EstimateOutput(n_in)
    ostride = (est + 3) & ~3

    # synthetic input resident in HBM: per-stream sine (200 + 1.7 s) Hz + 0.05 uniform noise
    x = torch.empty((rows, n_in), dtype=torch.float32, device=dev)
    t = torch.arange(n_in, device=dev, dtype=torch.float64) / IN_RATE
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    blk = 128
    for r0 in range(0, rows, blk):
        r1 = min(rows, r0 + blk)
        f = 200.0 + 1.7 * torch.arange(first + r0, first + r1, device=dev, dtype=torch.float64)
        sig = torch.sin(2 * np.pi * f[:, None] * t[None, :]).to(torch.float32)
        sig += 0.05 * (2 * torch.rand((r1 - r0, n_in), device=dev, generator=g) - 1)
        x[r0:r1] = sig
    del sig, t
    y = torch.zeros((rows, ostride), dtype=torch.float32, device=dev)
    # a dedicated (non-NULL) stream: the kernels are enqueued on it through the C ABI and the CUDA events
    # that time them are recorded on the same stream
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0
    torch.cuda.synchronize()

    def step_device(ev=None):
        b.Reset()
        if ev:
            ev[0].record(tstream)
        n1 = b.process_batch_dev(x.data_ptr(), n_in, n_in, y.data_ptr(), ostride, ostride, stream)
        if ev:
            ev[1].record(tstream)
        n2 = b.flush_batch_dev(y.data_ptr() + n1 * 4, ostride, ostride - n1, stream)
        if ev:
            ev[2].record(tstream)
        return n1, n2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak32 = G.measure_fma_peak(np.float32, local)

    for _ in range(max(a.warmup, 3)):
        n1, n2 = step_device()
    barrier()
    G.kernel_launches(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_begin.record(tstream)
    for k in range(a.steps):
        n1, n2 = step_device(evs[k])
    e_end.record(tstream)
    barrier()
    launches = G.kernel_launches()
    total_ms = e_begin.elapsed_time(e_end)
    time.sleep(0.15)
    clocks = sampler.stop()
    ms_step = total_ms / a.steps
    k_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)  # the batched FIR kernel (Process)
    f_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)

    tmax = torch.tensor([ms_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step_max = float(tmax.item())
    out_per_rank = rows * (n1 + n2)
    value = out_per_rank * world / (ms_step_max * 1e-3) / 1e6

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this very workload
    # (profiles/r2_traffic.json); null when the launch shape differs (other --streams/--seconds/--gpus)
    traffic = None
    try:
        tj = json.loads((ROOT / "profiles" / "r2_traffic.json").read_text())
        if tj.get("algorithmic_bytes_per_launch") == rows * (n_in + n1) * 4 and tj.get("kernel") == kname:
            traffic = {"bytes_per_launch": tj["traffic_bytes_per_launch"],
                       "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"],
                       "ratio": round(tj["traffic_bytes_per_launch"] / tj["algorithmic_bytes_per_launch"], 4),
                       "source": tj["source"]}
    except Exception:
        traffic = None
    achieved_tf = rows * n1 * flops_per_out / (k_ms * 1e-3) / 1e12
    bytes_alg = rows * (n_in + n1) * 4.0
    roofline = {"bound": "fma", "kernel": kname, "achieved": round(achieved_tf, 3), "peak": round(peak32, 3),
                "unit": "TFLOP/s", "frac": round(achieved_tf / peak32, 4),
                "peak_source": "dependent-FMA fp32 probe (gar_measure_fma_peak) run in this job; MEASURED_PEAKS.json "
                               "holds only HBM and bf16-tensor peaks, neither bounds this fp32-FMA kernel",
                "flops_per_output": flops_per_out, "kernel_ms": round(k_ms, 4), "flush_ms": round(f_ms, 4),
                "traffic": traffic,
                "hbm": {"achieved": round(bytes_alg / (k_ms * 1e-3) / 1e9, 1), "peak": 6446.9, "unit": "GB/s",
                        "frac": round(bytes_alg / (k_ms * 1e-3) / 1e9 / 6446.9, 4),
                        "algorithmic_bytes_per_output": 16}}

    # ---- e2e: the host-facing C-ABI call (gar_process_batch + gar_flush_batch) on pinned host buffers ----
    e2e = None
    xh = None
    if not a.no_e2e:
        # pinned caller buffers, first-touched by a thread bound to this GPU's NUMA node (gar_host_alloc_rows)
        xh, px = b.host_alloc_rows(n_in, np.float32)
        yh, py = b.host_alloc_rows(ostride, np.float32)
        for r0 in range(0, rows, 256):
            xh[r0:r0 + 256] = x[r0:r0 + 256].cpu().numpy()

        def step_host():
            b.Reset()
            _, m1 = b.ProcessBatch(xh, yh)
            _, m2 = b.FlushBatch(yh[:, m1:])
            return m1, m2

        for _ in range(2):
            m1, m2 = step_host()
        assert (m1, m2) == (n1, n2)
        ne = max(3, min(a.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            step_host()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ne
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": round(out_per_rank * world / float(tt.item()) / 1e6, 3), "unit": UNIT,
               "h2d_bytes_per_step": rows * n_in * 4 * world, "d2h_bytes_per_step": rows * (n1 + n2) * 4 * world,
               "ms_per_step": round(float(tt.item()) * 1e3, 3),
               "api": "gar_process_batch + gar_flush_batch (pinned host buffers, sliced H2D/compute/D2H overlap)",
               "timer": "host wall clock around the synchronous calls, max over ranks"}
        # spot-check the host path against the device path
        chk = torch.from_numpy(yh[:2, :n1 + n2].copy()).to(dev)
        assert torch.equal(chk, y[:2, :n1 + n2]), "host and device paths disagree"

        # ---- in-job copy ceiling: the same bytes as one e2e step, plain cudaMemcpyAsync, one call per ~64 MiB slice, H2D and
        # D2H concurrently on two streams, all ranks at once (barrier on both sides, max over ranks) ----
        if not a.no_ceiling:
            L = G.lib()
            s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            slice_rows = max(1, (64 << 20) // (n_in * 4))
            out_cols = n1 + n2

            def copy_pass():
                for r0 in range(0, rows, slice_rows):
                    cnt = min(slice_rows, rows - r0)
                    L.gar_memcpy_async(x.data_ptr() + r0 * n_in * 4, xh.ctypes.data + r0 * n_in * 4, cnt * n_in * 4, 1,
                                       s_h2d.cuda_stream)
                    L.gar_memcpy_async(yh.ctypes.data + r0 * ostride * 4, y.data_ptr() + r0 * ostride * 4,
                                       cnt * ostride * 4, 2, s_d2h.cuda_stream)
            copy_pass()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                copy_pass()
            torch.cuda.synchronize()
            ct = torch.tensor([(time.perf_counter() - t0) / 3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ct, op=dist.ReduceOp.MAX)
            cs = float(ct.item())
            e2e["copy_ceiling"] = {
                "ms_per_step": round(cs * 1e3, 3),
                "h2d_gbs_per_gpu": round(rows * n_in * 4 / cs / 1e9, 2),
                "d2h_gbs_per_gpu": round(rows * ostride * 4 / cs / 1e9, 2),
                "value_at_ceiling": round(out_per_rank * world / cs / 1e6, 1),
                "e2e_over_ceiling": round(cs / float(tt.item()), 4),
                "how": "same H2D + D2H bytes as one e2e step per rank, plain cudaMemcpyAsync per 64 MiB slice on two "
                       "streams, all ranks concurrently, wall clock, max over ranks (out_cols=%d padded to %d)" % (out_cols, ostride)}
        e2e["numa"] = numa

    # ---- parity spot check of THIS run's output (outside every timed region): random rows of the device result against the
    # CPU oracle on the very same input rows ----
    parity_spot = None
    os.sched_setaffinity(0, affinity0)  # the CPU legs use every host core again
    if rank == 0 and a.parity_rows > 0:
        from oracle import oracle as O
        step_device()
        torch.cuda.synchronize()
        pick = sorted(np.random.default_rng(2026).choice(rows, size=min(a.parity_rows, rows), replace=False).tolist())
        xr = x[pick].cpu().numpy()
        got = y[pick, :n1 + n2].cpu().numpy().astype(np.float64)
        want, cnts = O.batch_resample(xr, IN_RATE, OUT_RATE, O.preset_to_engine_quality(PRESETS[a.preset]),
                                      n_threads=min(len(pick), os.cpu_count() or 1))
        err = float(np.max(np.abs(got - want[:, :n1 + n2].astype(np.float64))))
        parity_spot = {"rows": len(pick), "row_ids": pick, "samples_per_row": int(n1 + n2), "counts_equal": bool(np.all(cnts == n1 + n2)),
                       "max_abs_err": err, "tolerance": 1e-6, "ok": bool(err <= 1e-6 and np.all(cnts == n1 + n2)),
                       "referee": "oracle/ (C++ restatement of the Go path) on the same input rows, full length"}
        assert parity_spot["ok"], parity_spot

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        xc = xh if xh is not None else x[:min(rows, 512)].cpu().numpy()
        cpu = cpu_baseline(a, xc, os.cpu_count() or 1, a.cpu_seconds)

    extra = None
    if rank == 0 and world == 1 and not a.no_extra:
        try:
            del x, y
            torch.cuda.empty_cache()
            extra = extra_f64(dev, local, tstream)
        except Exception as e:  # side measurement only: never fail the headline line
            extra = {"error": repr(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step_max, 4), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "streams_per_gpu": rows,
                           "l2": "inputs larger than L2 (%.1f GB per GPU per step)" % (rows * n_in * 4 / 1e9),
                           "timer": "CUDA events on the launching stream, max over ranks"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "parity_spot": parity_spot, "extra_f64": extra}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
