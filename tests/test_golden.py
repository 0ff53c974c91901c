"""Committed golden vectors (tests/golden/baseline_configs.npz, made by tests/golden/make_golden.py from the pinned
oracle): the oracle must keep reproducing them (no silent drift), their counts are the ones SURVEY.md §8a predicts from
the reference's formulas, and the GPU engine must match them on all five BASELINE configs."""
from pathlib import Path

import numpy as np
import pytest

from helpers import G, O, sig_c1, sig_c2, sig_c3, sig_c4, sig_c5a

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "baseline_configs.npz")
HEAD, TAIL = 3000, 600
TOL64, TOL32 = 1e-12, 1e-6


def excerpt(y):
    y = np.asarray(y)
    return np.concatenate([y[:HEAD], y[-TAIL:]]) if len(y) > HEAD + TAIL else y.copy()


def test_golden_counts_are_the_predicted_ones():  # SURVEY.md §8a "restatement-predicted sample counts"
    assert GOLD["c1_counts"].tolist() == [479787, 215]
    for ch in ("l", "r"):
        c = GOLD[f"c2{ch}_chunk_counts"]
        assert c[:4].tolist() == [3556, 3763, 3763, 3764] and int(c[:-1].sum()) == 440793 and int(c[-1]) == 209
    assert all(row.tolist() == [479389, 612] for row in GOLD["c3_counts"])
    assert GOLD["c4_medium_counts"].tolist() == [159708 + 293] * 4
    assert GOLD["c4_high_counts"].tolist() == [159626 + 375] * 4
    assert GOLD["c5a_counts"].tolist() == [1910673, 9375]
    assert GOLD["c5b_counts"].tolist() == [479789, 203]


def test_oracle_reproduces_the_golden_vectors():
    e = O.Engine(44100, 48000, O.preset_to_engine_quality(O.PRESET_HIGH))
    y = np.concatenate([e.process(sig_c1()), e.flush()])
    np.testing.assert_array_equal(excerpt(y), GOLD["c1_out"])
    x4 = sig_c4(4, 480000)
    y4, c4 = O.batch_resample(x4, 48000, 16000, O.Q_MEDIUM, n_threads=4)
    assert c4.tolist() == GOLD["c4_medium_counts"].tolist()
    np.testing.assert_array_equal(np.stack([excerpt(y4[i, :c4[i]]) for i in range(4)]), GOLD["c4_medium_out"])
    e = O.Engine(44100, 47999, O.preset_to_engine_quality(O.PRESET_HIGH))
    y = np.concatenate([e.process(sig_c1()), e.flush()])
    np.testing.assert_array_equal(excerpt(y), GOLD["c5b_out"])


def _cfg(ir, orr, ch=1, preset=G.QualityHigh):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset))


@pytest.mark.gpu
def test_gpu_matches_golden_c1_c5a_c5b():
    r = G.NewEngine(44100, 48000, G.QualityHigh)
    y, f = r.Process(sig_c1()), r.Flush()
    assert [len(y), len(f)] == GOLD["c1_counts"].tolist()
    assert np.max(np.abs(excerpt(np.concatenate([y, f])) - GOLD["c1_out"])) <= TOL64
    r = G.New(_cfg(8000, 192000))
    y, f = r.Process(sig_c5a()), r.Flush()
    assert [len(y), len(f)] == GOLD["c5a_counts"].tolist()
    assert np.max(np.abs(excerpt(np.concatenate([y, f])) - GOLD["c5a_out"])) <= TOL64
    r = G.NewEngine(44100, 47999, G.QualityHigh)
    y, f = r.Process(sig_c1()), r.Flush()
    assert [len(y), len(f)] == GOLD["c5b_counts"].tolist()
    assert np.max(np.abs(excerpt(np.concatenate([y, f])) - GOLD["c5b_out"])) <= TOL64


@pytest.mark.gpu
def test_gpu_matches_golden_c2_streaming_float32():
    for name, ch in zip(("l", "r"), sig_c2()):
        r = G.New(_cfg(48000, 44100))
        out = np.empty(r.EstimateOutput(4096), dtype=np.float32)
        outs, counts = [], []
        for i in range(0, len(ch), 4096):
            n = r.ProcessFloat32Into(ch[i:i + 4096], out)
            outs.append(out[:n].copy())
            counts.append(n)
        fl = r.Flush()
        assert counts + [len(fl)] == GOLD[f"c2{name}_chunk_counts"].tolist()
        got = excerpt(np.concatenate(outs)).astype(np.float64)
        assert np.max(np.abs(got - GOLD[f"c2{name}_out"].astype(np.float64))) <= TOL32
        assert np.max(np.abs(np.asarray(fl, np.float64) - GOLD[f"c2{name}_flush"])) <= TOL64


@pytest.mark.gpu
def test_gpu_matches_golden_c3_and_c4():
    r = G.New(_cfg(96000, 48000, 8, G.QualityVeryHigh))
    ys = r.ProcessMulti(sig_c3())
    fs = r.FlushMulti()
    assert [[len(a), len(b)] for a, b in zip(ys, fs)] == GOLD["c3_counts"].tolist()
    got = np.stack([excerpt(np.concatenate([a, b])) for a, b in zip(ys, fs)])
    assert np.max(np.abs(got - GOLD["c3_out"])) <= TOL64
    x4 = sig_c4(4, 480000)
    for nm, preset in (("low", G.QualityLow), ("medium", G.QualityMedium), ("high", G.QualityHigh)):
        b = G.NewBatch(48000, 16000, preset, 4, np.float32)
        y, ny = b.ProcessBatch(x4)
        f, nf = b.FlushBatch()
        assert [ny + nf] * 4 == GOLD[f"c4_{nm}_counts"].tolist()
        got = np.stack([excerpt(np.concatenate([y[i], f[i]])) for i in range(4)]).astype(np.float64)
        assert np.max(np.abs(got - GOLD[f"c4_{nm}_out"].astype(np.float64))) <= TOL32
