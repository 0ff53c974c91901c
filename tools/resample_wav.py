#!/usr/bin/env python
"""resample_wav.py — the reference's `resample-wav` caller (cmd/resample-wav) on the B200 engine.

    python tools/resample_wav.py -rate 48000 [-quality high] [-fast] in.wav out.wav

WAV I/O stays on the host (stdlib `wave`, 16/24/32-bit PCM); each block of interleaved PCM goes through ONE call
(`gar_process_interleaved`): deinterleave + normalise, resample all channels, clamp + scale + interleave run on
the device. Flags follow cmd/resample-wav/main.go:84-99 (-rate, -quality quick|low|medium|high|veryhigh, -fast =
float32). Like the reference it drives engine.Resampler per channel with engine-level quality.
"""
import argparse
import sys
import time
import wave
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
import gar_b200 as G  # noqa: E402

QUAL = {"quick": G.EngineQualityQuick, "low": G.EngineQualityLow, "medium": G.EngineQualityMedium,
        "high": G.EngineQualityHigh, "veryhigh": G.EngineQualityVeryHigh}


def read_block(w, n, width, channels):
    raw = w.readframes(n)
    if width == 2:
        return np.frombuffer(raw, dtype="<i2").astype(np.int32).reshape(-1, channels)
    if width == 4:
        return np.frombuffer(raw, dtype="<i4").reshape(-1, channels)
    b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)  # 24-bit little endian
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32).reshape(-1, channels)


def write_block(w, x, width):
    if width == 2:
        w.writeframes(x.astype("<i2").tobytes())
    elif width == 4:
        w.writeframes(x.astype("<i4").tobytes())
    else:
        v = x.astype(np.int32).reshape(-1) & 0xFFFFFF
        out = np.empty((v.size, 3), dtype=np.uint8)
        out[:, 0], out[:, 1], out[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
        w.writeframes(out.tobytes())


def main():
    ap = argparse.ArgumentParser(prefix_chars="-")
    ap.add_argument("-rate", type=int, required=True)
    ap.add_argument("-quality", default="high", choices=list(QUAL))
    ap.add_argument("-fast", action="store_true", help="float32 processing (cmd/resample-wav/main.go:96)")
    ap.add_argument("-block", type=int, default=65536)
    ap.add_argument("input")
    ap.add_argument("output")
    a = ap.parse_args()
    with wave.open(a.input, "rb") as wi:
        ch, width, rate, frames = wi.getnchannels(), wi.getsampwidth(), wi.getframerate(), wi.getnframes()
        bits = width * 8
        h = G.SimpleResampler(rate, a.rate, G.QualityHigh, np.float32 if a.fast else np.float64,
                              engine_quality=QUAL[a.quality], n_streams=ch)
        t0 = time.perf_counter()
        n_out = 0
        with wave.open(a.output, "wb") as wo:
            wo.setnchannels(ch)
            wo.setsampwidth(width)
            wo.setframerate(a.rate)
            while True:
                blk = read_block(wi, a.block, width, ch)
                if len(blk) == 0:
                    break
                y = h.ProcessInterleaved(blk, bits)
                write_block(wo, y, width)
                n_out += len(y)
            y = h.FlushInterleaved(np.int32, bits)
            write_block(wo, y, width)
            n_out += len(y)
        dt = time.perf_counter() - t0
    dur = frames / rate
    print(f"{a.input}: {ch} ch, {bits}-bit, {rate} Hz, {frames} frames -> {a.rate} Hz, {n_out} frames "
          f"in {dt*1e3:.1f} ms ({dur/dt:.0f}x realtime)")


if __name__ == "__main__":
    main()
