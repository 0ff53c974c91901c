#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/gar_oracle.cpp), the restatement of the Go path that is
pinned to the reference's own known answers (tests/test_oracle_kat.py). The reference ships no sample vectors
(SURVEY.md §8c), so these are this repo's goldens: output counts (bit-exact) and excerpts of the output samples of the five
BASELINE configs on the documented synthetic inputs (tests/helpers.py).

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from helpers import O, sig_c1, sig_c2, sig_c3, sig_c4, sig_c5a  # noqa: E402

HEAD, TAIL = 3000, 600  # samples kept from the start and the end of every output (flush included)


def excerpt(y):
    y = np.asarray(y)
    return np.concatenate([y[:HEAD], y[-TAIL:]]) if len(y) > HEAD + TAIL else y.copy()


def main():
    g = {}
    # C1: ResampleMono 44.1k -> 48k QualityHigh float64
    x = sig_c1()
    e = O.Engine(44100, 48000, O.preset_to_engine_quality(O.PRESET_HIGH))
    y, f = e.process(x), e.flush()
    g["c1_counts"] = np.array([len(y), len(f)], np.int64)
    g["c1_out"] = excerpt(np.concatenate([y, f]))
    # C2: stereo 48k -> 44.1k High, float32 I/O, 4096-frame chunks + Flush (two mono instances)
    for name, ch in zip(("l", "r"), sig_c2()):
        p = O.Pipeline(48000, 44100, 1, O.PRESET_HIGH)
        buf = np.empty(p.estimate_output(4096), dtype=np.float32)
        outs, counts = [], []
        for i in range(0, len(ch), 4096):
            n = p.process_f32_into(ch[i:i + 4096], buf)
            outs.append(buf[:n].copy())
            counts.append(n)
        fl = p.flush()
        g[f"c2{name}_chunk_counts"] = np.array(counts + [len(fl)], np.int64)
        g[f"c2{name}_out"] = excerpt(np.concatenate(outs))
        g[f"c2{name}_flush"] = np.asarray(fl, np.float64)
    # C3: 8 channels 96k -> 48k VeryHigh float64, ProcessMulti + FlushMulti
    xs = sig_c3()
    p = O.Pipeline(96000, 48000, 8, O.PRESET_VERYHIGH)
    ys, fs = p.process_multi(xs), p.flush_multi()
    g["c3_counts"] = np.array([[len(a), len(b)] for a, b in zip(ys, fs)], np.int64)
    g["c3_out"] = np.stack([excerpt(np.concatenate([a, b])) for a, b in zip(ys, fs)])
    # C4: mono streams 48k -> 16k float32 (streams 0..3 of the 4096), Low / Medium / High
    x4 = sig_c4(4, 480000)
    for nm, q in (("low", O.Q_LOW), ("medium", O.Q_MEDIUM), ("high", O.Q_HIGH)):
        y4, c4 = O.batch_resample(x4, 48000, 16000, q, n_threads=4)
        g[f"c4_{nm}_counts"] = np.asarray(c4, np.int64)
        g[f"c4_{nm}_out"] = np.stack([excerpt(y4[i, :c4[i]]) for i in range(4)])
    # C5a: 8k -> 192k High (path A, multistage); C5b: 44.1k -> 47.999k High (cubic coefficient interpolation)
    p = O.Pipeline(8000, 192000, 1, O.PRESET_HIGH)
    y, f = p.process(sig_c5a()), p.flush()
    g["c5a_counts"] = np.array([len(y), len(f)], np.int64)
    g["c5a_out"] = excerpt(np.concatenate([y, f]))
    e = O.Engine(44100, 47999, O.preset_to_engine_quality(O.PRESET_HIGH))
    y, f = e.process(sig_c1()), e.flush()
    g["c5b_counts"] = np.array([len(y), len(f)], np.int64)
    g["c5b_out"] = excerpt(np.concatenate([y, f]))
    np.savez_compressed(HERE / "baseline_configs.npz", **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
