#!/usr/bin/env python
"""resample_wav.py — the reference's `resample-wav` caller (cmd/resample-wav) on the B200 engine.

    python tools/resample_wav.py -rate 48000 [-quality high] [-fast] in.wav out.wav
    python tools/resample_wav.py -rate 48000 [-quality high] [-fast] -outdir DIR a.wav b.wav c.wav ...     (batch)

WAV I/O stays on the host (stdlib `wave`, 16/24/32-bit PCM); each block of interleaved PCM goes through ONE call
(`gar_process_interleaved`): deinterleave + normalise, resample all channels, clamp + scale + interleave run on
the device. Flags follow cmd/resample-wav/main.go:84-99 (-rate, -quality quick|low|medium|high|veryhigh, -fast =
float32); blocks are the reference's 65 536 frames (main.go:38). Like the reference it drives one engine.Resampler per
channel with an engine-level quality (helpers.go:77-96) — here the channels are rows of one lock-step device pass, the
GPU analogue of the per-channel goroutines (helpers.go:242-279).

Batch mode goes one step further: files with the same rate / channel count / sample width are ROWS OF THE SAME PASS
(files x channels rows, at most 256 per group). Files shorter than the longest of their group are continued with zero
frames and their output is cut at the exact frame count the file yields on its own (the integer state machine of a
geometry-only handle gives it): a FIR chain followed by zeros produces exactly the samples Flush produces, so every
file's output is identical to a single-file run (tests/test_resample_wav.py checks that bit for bit).
"""
import argparse
import sys
import time
import wave
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

QUAL = {"quick": G.EngineQualityQuick, "low": G.EngineQualityLow, "medium": G.EngineQualityMedium,
        "high": G.EngineQualityHigh, "veryhigh": G.EngineQualityVeryHigh}
BLOCK = 65536  # cmd/resample-wav/main.go:38


def decode(raw, width, channels):
    if width == 2:
        return np.frombuffer(raw, dtype="<i2").astype(np.int32).reshape(-1, channels)
    if width == 4:
        return np.frombuffer(raw, dtype="<i4").reshape(-1, channels)
    b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)  # 24-bit little endian
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32).reshape(-1, channels)


def encode(x, width):
    if width == 2:
        return x.astype("<i2").tobytes()
    if width == 4:
        return x.astype("<i4").tobytes()
    v = x.astype(np.int32).reshape(-1) & 0xFFFFFF
    out = np.empty((v.size, 3), dtype=np.uint8)
    out[:, 0], out[:, 1], out[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
    return out.tobytes()


def _open_out(path, ch, width, rate):
    wo = wave.open(str(path), "wb")
    wo.setnchannels(ch)
    wo.setsampwidth(width)
    wo.setframerate(rate)
    return wo


def resample_file(src, dst, rate, quality="high", fast=False, block=BLOCK):
    """One file: the reference's block loop (helpers.go:77-334). Returns (frames_in, frames_out, seconds)."""
    with wave.open(str(src), "rb") as wi:
        ch, width, in_rate, frames = wi.getnchannels(), wi.getsampwidth(), wi.getframerate(), wi.getnframes()
        bits = width * 8
        h = G.SimpleResampler(in_rate, rate, G.QualityHigh, np.float32 if fast else np.float64,
                              engine_quality=QUAL[quality], n_streams=ch)
        t0 = time.perf_counter()
        n_out = 0
        with _open_out(dst, ch, width, rate) as wo:
            while True:
                blk = decode(wi.readframes(block), width, ch)
                if len(blk) == 0:
                    break
                y = h.ProcessInterleaved(blk, bits)
                wo.writeframes(encode(y, width))
                n_out += len(y)
            y = h.FlushInterleaved(np.int32, bits)  # flushAndPadChannels (helpers.go:293-334)
            wo.writeframes(encode(y, width))
            n_out += len(y)
        return frames, n_out, time.perf_counter() - t0


def expected_frames(in_rate, rate, quality, fast, frames, block=BLOCK):
    """Frames a file of `frames` frames yields on its own (Process per block + Flush), from the integer state machine."""
    g = G.SimpleResampler(in_rate, rate, G.QualityHigh, np.float32 if fast else np.float64,
                          engine_quality=QUAL[quality], device=-1)
    total, left = 0, frames
    while left > 0:
        n = min(block, left)
        total += g.advance_geometry(n)
        left -= n
    return total + g.advance_geometry(0, flush=True)


def resample_batch(srcs, outdir, rate, quality="high", fast=False, block=BLOCK):
    """Many files: files of equal (rate, channels, width) are rows of the same lock-step pass. Returns per-file results."""
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    groups = defaultdict(list)
    for s in srcs:
        with wave.open(str(s), "rb") as w:
            groups[(w.getframerate(), w.getnchannels(), w.getsampwidth())].append(Path(s))
    results = {}
    for (in_rate, ch, width), files in groups.items():
        per = max(1, 256 // ch)  # rows = files x channels, at most 256 per handle (constants.go:48-51 channel bound)
        for g0 in range(0, len(files), per):
            part = files[g0:g0 + per]
            results.update(_batch_group(part, outdir, in_rate, ch, width, rate, quality, fast, block))
    return results


def _batch_group(files, outdir, in_rate, ch, width, rate, quality, fast, block):
    bits, nf = width * 8, len(files)
    ins = [wave.open(str(f), "rb") for f in files]
    frames = [w.getnframes() for w in ins]
    want = [expected_frames(in_rate, rate, quality, fast, n, block) for n in frames]
    outs = [_open_out(outdir / f.name, ch, width, rate) for f in files]
    written = [0] * nf
    h = G.SimpleResampler(in_rate, rate, G.QualityHigh, np.float32 if fast else np.float64,
                          engine_quality=QUAL[quality], n_streams=ch * nf)
    t0 = time.perf_counter()

    def emit(y):  # y: [frames, nf*ch] -> per file, cut at the file's own frame count
        for i in range(nf):
            k = min(len(y), want[i] - written[i])
            if k > 0:
                outs[i].writeframes(encode(np.ascontiguousarray(y[:k, i * ch:(i + 1) * ch]), width))
                written[i] += k

    longest = max(frames)
    for off in range(0, longest, block):
        n = min(block, longest - off)
        blk = np.zeros((n, nf * ch), dtype=np.int32)
        for i, w in enumerate(ins):
            d = decode(w.readframes(n), width, ch)  # shorter files simply run out: zero frames from there on
            blk[:len(d), i * ch:(i + 1) * ch] = d
        emit(h.ProcessInterleaved(blk, bits))
    emit(h.FlushInterleaved(np.int32, bits))
    # files much shorter than the longest of the group could in principle need samples beyond the group's flush
    pad = np.zeros((block, nf * ch), dtype=np.int32)
    while any(written[i] < want[i] for i in range(nf)):
        emit(h.ProcessInterleaved(pad, bits))
    dt = time.perf_counter() - t0
    for w in ins + outs:
        w.close()
    return {str(f): (frames[i], written[i], dt) for i, f in enumerate(files)}


def main(argv=None):
    ap = argparse.ArgumentParser(prefix_chars="-")
    ap.add_argument("-rate", type=int, required=True)
    ap.add_argument("-quality", default="high", choices=list(QUAL))
    ap.add_argument("-fast", action="store_true", help="float32 processing (cmd/resample-wav/main.go:96)")
    ap.add_argument("-block", type=int, default=BLOCK)
    ap.add_argument("-outdir", default="", help="batch mode: every input file is written to OUTDIR under its own name")
    ap.add_argument("files", nargs="+")
    a = ap.parse_args(argv)
    if a.outdir:
        res = resample_batch(a.files, a.outdir, a.rate, a.quality, a.fast, a.block)
        tot_in = sum(v[0] for v in res.values())
        dt = max(v[2] for v in res.values())
        print(f"{len(res)} files, {tot_in} frames in, batched as rows of lock-step passes in {dt*1e3:.1f} ms")
        return 0
    if len(a.files) != 2:
        ap.error("expected IN.wav OUT.wav (or -outdir DIR with any number of inputs)")
    frames, n_out, dt = resample_file(a.files[0], a.files[1], a.rate, a.quality, a.fast, a.block)
    with wave.open(a.files[0], "rb") as w:
        ch, bits, rate = w.getnchannels(), w.getsampwidth() * 8, w.getframerate()
    print(f"{a.files[0]}: {ch} ch, {bits}-bit, {rate} Hz, {frames} frames -> {a.rate} Hz, {n_out} frames "
          f"in {dt*1e3:.1f} ms ({frames/rate/dt:.0f}x realtime)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
