"""Shared test helpers: package import path, BASELINE configs as synthetic signals."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "go-audio-resampler_b200" / "python"
if str(PKG) not in sys.path:
    sys.path.insert(0, str(PKG))
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import gar_b200 as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

KIND = {G.STAGE_UP: "up", G.STAGE_DECIM: "dec", G.STAGE_POLY: "poly", G.STAGE_CUBIC: "cubic"}


def geometry_config(in_rate, out_rate, channels=1, preset=G.QualityHigh, precision=0):
    q = G.QualitySpec(Preset=preset, Precision=precision, PhaseResponse=50.0, PassbandEnd=0.9, StopbandBegin=0.99)
    return G.Config(InputRate=in_rate, OutputRate=out_rate, Channels=channels, Quality=q, Device=-1)


def oracle_chain_desc(pipe: "O.Pipeline", ch=0):
    """Flatten the oracle pipeline's per-engine descriptions into primitive stages like the product's."""
    out = []
    for i, _ in enumerate(pipe.stages()):
        out.extend(flatten_engine_desc(pipe.stage_describe(ch, i)))
    return out


def flatten_engine_desc(d):
    out = []
    if d["cubic"]:
        out.append(dict(kind="cubic"))
    if d["pre_factor"] > 1:
        out.append(dict(kind="up", factor=d["pre_factor"], taps=d["pre_tpp"], hist_len=d["pre_hist"]))
    if d["dec_factor"] > 1:
        out.append(dict(kind="dec", factor=d["dec_factor"], taps=d["dec_taps"], hist_len=d["dec_hist"],
                        decim_phase=d["dec_phase"]))
    if d["poly_L"] > 0:
        out.append(dict(kind="poly", factor=d["poly_L"], taps=d["poly_tpp"], step=d["poly_step"], at=d["poly_at"],
                        hist_len=d["poly_hist"]))
    return out


def product_chain_desc(h, stream=0):
    out = []
    for d in h.describe(stream):
        k = KIND[d["kind"]]
        e = dict(kind=k)
        if k != "cubic":
            e.update(factor=d["factor"], taps=d["taps"], hist_len=d["hist_len"])
        if k == "dec":
            e["decim_phase"] = d["decim_phase"]
        if k == "poly":
            e.update(step=d["step"], at=d["at"])
        out.append(e)
    return out


# ---- BASELINE.md §3 synthetic inputs -----------------------------------------------------------
def sig_c1(n=441000):
    return np.sin(2 * np.pi * 1000.0 * np.arange(n) / 44100.0)


def sig_c2(n=480000):
    t = np.arange(n) / 48000.0
    left = (np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 1320 * t)).astype(np.float32)
    right = (np.sin(2 * np.pi * 554.37 * t) + 0.1 * np.sin(2 * np.pi * 3 * 554.37 * t)).astype(np.float32)
    return left, right


def sig_c3(n=960000, channels=8):
    out = []
    t = np.arange(n) / 96000.0
    for c in range(channels):
        rng = np.random.default_rng(4242 + c)
        p1, p2 = rng.uniform(0, 2 * np.pi, 2)
        out.append(0.7 * np.sin(2 * np.pi * 440 * t + p1) + 0.2 * np.sin(2 * np.pi * 1750 * t + p2)
                   + 0.1 * (rng.random(n) - 0.5))
    return out


def sig_c4(n_streams, n, dtype=np.float32, first_stream=0):
    t = np.arange(n) / 48000.0
    x = np.empty((n_streams, n), dtype=dtype)
    for i in range(n_streams):
        s = first_stream + i
        rng = np.random.default_rng(s)
        x[i] = (np.sin(2 * np.pi * (200 + 1.7 * s) * t) + 0.05 * (rng.random(n) - 0.5) * 2).astype(dtype)
    return x


def sig_c5a(n=80000):
    return np.sin(2 * np.pi * 1000.0 * np.arange(n) / 8000.0)
