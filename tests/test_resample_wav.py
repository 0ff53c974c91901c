"""SURVEY.md §8f N3 — the `resample-wav` caller on the B200 engine (tools/resample_wav.py), end to end on generated 16- and
24-bit WAV files against a numpy restatement of cmd/resample-wav/helpers.go:77-334 + main.go:425-520 that drives one oracle
engine per channel (block loop of 65 536 frames, deinterleave + normalise, Process, clamp + truncate + interleave, flush and
pad), and the multi-file batch mode against single-file runs."""
import sys
import wave

import numpy as np
import pytest

from helpers import O, ROOT

sys.path.insert(0, str(ROOT / "tools"))
import resample_wav as RW  # noqa: E402

MAXV = {16: 32767.0, 24: 8388607.0, 32: 2147483647.0}
QUAL = {"medium": O.Q_MEDIUM, "high": O.Q_HIGH, "veryhigh": O.Q_VERYHIGH}


def make_wav(path, frames, channels, width, rate, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(frames) / rate
    cols = [0.5 * np.sin(2 * np.pi * (220 + 97 * c + seed) * t) + 0.2 * (rng.random(frames) - 0.5) for c in range(channels)]
    pcm = np.trunc(np.stack(cols, axis=1) * MAXV[width * 8]).astype(np.int32)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(rate)
        w.writeframes(RW.encode(pcm, width))
    return pcm


def read_wav(path):
    with wave.open(str(path), "rb") as w:
        return RW.decode(w.readframes(w.getnframes()), w.getsampwidth(), w.getnchannels()), w.getframerate()


def reference_restatement(pcm, channels, bits, in_rate, out_rate, quality, F, block=RW.BLOCK):
    """helpers.go:77-334: one engine per channel; per block deinterleaveInto -> Process -> interleaveInto; then
    flushAndPadChannels."""
    inv, mv = 1.0 / MAXV[bits], MAXV[bits]
    engines = [O.Engine(in_rate, out_rate, QUAL[quality], F) for _ in range(channels)]
    out = []

    def interleave(chans):
        n = max(len(c) for c in chans)
        cols = []
        for c in chans:
            p = np.zeros(n, dtype=np.float64)  # pad shorter channels (helpers.go:318-325)
            p[:len(c)] = c
            cols.append(np.trunc(np.clip(p, -1.0, 1.0) * mv).astype(np.int64))
        return np.stack(cols, axis=1) if n else np.zeros((0, len(chans)), np.int64)

    for i in range(0, len(pcm), block):
        blk = pcm[i:i + block].astype(np.float64) * inv
        out.append(interleave([e.process(np.ascontiguousarray(blk[:, c]).astype(F)) for c, e in enumerate(engines)]))
    out.append(interleave([e.flush() for e in engines]))
    return np.concatenate(out, axis=0)


@pytest.mark.gpu
@pytest.mark.parametrize("channels,width,in_rate,out_rate,quality,fast,frames", [
    (2, 2, 44100, 48000, "high", False, 150000),     # 16-bit stereo CD -> DAT, three blocks
    (1, 3, 48000, 16000, "medium", True, 100000),    # 24-bit mono speech decimation, float32 (-fast)
    (2, 3, 96000, 44100, "veryhigh", False, 70000),  # 24-bit stereo
    (6, 2, 48000, 32000, "high", False, 66000)])     # 5.1, 16-bit
def test_resample_wav_end_to_end_matches_the_reference_restatement(tmp_path, channels, width, in_rate, out_rate, quality,
                                                                    fast, frames):
    pcm = make_wav(tmp_path / "in.wav", frames, channels, width, in_rate, 5)
    argv = ["-rate", str(out_rate), "-quality", quality] + (["-fast"] if fast else []) + [str(tmp_path / "in.wav"), str(tmp_path / "out.wav")]
    assert RW.main(argv) == 0
    got, rate = read_wav(tmp_path / "out.wav")
    assert rate == out_rate
    want = reference_restatement(pcm, channels, width * 8, in_rate, out_rate, quality, np.float32 if fast else np.float64)
    assert got.shape == want.shape  # frame count bit-exact
    # truncation toward zero: a float result within rounding distance of an integer boundary may land one LSB apart
    diff = np.abs(got.astype(np.int64) - want)
    assert int(diff.max()) <= (1 if not fast else 2)
    assert np.mean(diff > 0) < (1e-3 if not fast else 0.2)


@pytest.mark.gpu
def test_batch_mode_equals_single_file_runs(tmp_path):
    specs = [("a.wav", 90000, 2, 2, 44100), ("b.wav", 131072, 2, 2, 44100), ("c.wav", 65536, 2, 2, 44100),
             ("d.wav", 1000, 2, 2, 44100), ("e.wav", 50000, 1, 3, 48000), ("f.wav", 70001, 1, 3, 48000)]
    for i, (name, frames, ch, width, rate) in enumerate(specs):
        make_wav(tmp_path / name, frames, ch, width, rate, 10 + i)
    srcs = [str(tmp_path / s[0]) for s in specs]
    assert RW.main(["-rate", "48000", "-quality", "high", "-outdir", str(tmp_path / "batch")] + srcs[:4]) == 0
    assert RW.main(["-rate", "44100", "-quality", "high", "-outdir", str(tmp_path / "batch")] + srcs[4:]) == 0
    for name, frames, ch, width, rate in specs:
        target = 48000 if rate == 44100 else 44100
        RW.resample_file(tmp_path / name, tmp_path / ("single_" + name), target, "high")
        one, _ = read_wav(tmp_path / ("single_" + name))
        many, r2 = read_wav(tmp_path / "batch" / name)
        assert r2 == target and one.shape == many.shape, (name, one.shape, many.shape)
        # same samples; >= 8 rows may take the tensor-core kernels (taps grouped in fours: last-bit differences in float64),
        # which can move a value across a truncation boundary once in a long while
        diff = np.abs(one.astype(np.int64) - many.astype(np.int64))
        assert int(diff.max()) <= 1 and np.mean(diff > 0) < 1e-3, (name, int(diff.max()), float(np.mean(diff > 0)))


def test_expected_frames_is_chunking_independent_and_matches_the_oracle():
    for ir, orr, q in ((44100, 48000, "high"), (48000, 16000, "medium"), (96000, 44100, "veryhigh")):
        for frames in (1000, 65536, 70001):
            e = O.Engine(ir, orr, QUAL[q])
            n = len(e.process(np.zeros(frames))) + len(e.flush())
            assert RW.expected_frames(ir, orr, q, False, frames) == n
            assert abs(RW.expected_frames(ir, orr, q, False, frames, block=4096) - n) <= 0  # greedy stages: same total
