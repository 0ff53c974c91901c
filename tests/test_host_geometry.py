"""Host logic of the product (planner, filter design, integer streaming state machine) against the
oracle — runs without a GPU through geometry-only handles (gar_config.device = -1)."""
import numpy as np
import pytest

from helpers import G, O, geometry_config, oracle_chain_desc, flatten_engine_desc, product_chain_desc

PIPE_CASES = [  # (in, out, channels, preset, custom precision)
    (48000, 44100, 1, G.QualityHigh, 0),       # BASELINE C2
    (96000, 48000, 8, G.QualityVeryHigh, 0),   # C3
    (48000, 16000, 1, G.QualityHigh, 0),       # C4'
    (8000, 192000, 1, G.QualityHigh, 0),       # C5a
    (44100, 48000, 2, G.QualityMedium, 0),
    (44100, 48000, 1, G.QualityLow, 0),
    (44100, 48000, 1, G.QualityQuick, 0),
    (48000, 32000, 1, G.QualityHigh, 0),
    (44100, 22050, 1, G.QualityVeryHigh, 0),
    (22050, 44100, 1, G.QualityHigh, 0),
    (48000, 48000, 1, G.QualityHigh, 0),
    (48000, 6000, 1, G.QualityMedium, 0),
    (48000, 7350, 1, G.QualityHigh, 0),
    (11025, 96000, 1, G.QualityHigh, 0),
    (44100, 48000, 1, G.QualityCustom, 20),
    (44100, 48000, 1, G.QualityCustom, 28),
    (44100, 48000, 1, G.QualityCustom, 33),
]
ENGINE_CASES = [  # (in, out, preset, dtype)
    (44100, 48000, G.QualityHigh, np.float64),     # C1
    (48000, 16000, G.QualityMedium, np.float32),   # C4
    (48000, 16000, G.QualityHigh, np.float32),
    (48000, 16000, G.QualityLow, np.float32),
    (44100, 47999, G.QualityHigh, np.float64),     # C5b
    (8000, 192000, G.QualityHigh, np.float64),     # x24 single stage
    (48000, 8000, G.QualityMedium, np.float64),    # /6
    (48000, 44100, G.QualityVeryHigh, np.float32),
    (44100, 44100, G.QualityHigh, np.float64),     # pass-through
    (48000, 11025, G.QualityHigh, np.float64),     # steep non-integer down ratio
    (44100, 48000, G.QualityQuick, np.float64),    # path B maps Quick -> engine Low (not cubic)
]


def _chunks(rng, total, lo=1, hi=9000):
    out, left = [], total
    while left > 0:
        n = int(min(left, rng.integers(lo, hi)))
        out.append(n)
        left -= n
    return out


@pytest.mark.parametrize("ir,orr,ch,preset,prec", PIPE_CASES)
def test_pipeline_geometry_matches_oracle(ir, orr, ch, preset, prec):
    h = G.New(geometry_config(ir, orr, ch, preset, prec))
    p = O.Pipeline(ir, orr, ch, preset, prec)
    assert h.plan_types() == [t for t, _ in p.stages()]
    assert product_chain_desc(h) == oracle_chain_desc(p)
    assert h.EstimateOutput(4096) == p.estimate_output(4096)
    assert h.GetLatency() == p.latency()
    assert h.GetRatio() == p.ratio
    rng = np.random.default_rng(7)
    total = 40000
    for n in _chunks(rng, total) + [0, 1, 2]:
        want = len(p.process(np.zeros(n)))
        assert h.advance_geometry(n) == want
        assert product_chain_desc(h) == oracle_chain_desc(p)
    assert h.advance_geometry(0, flush=True) == len(p.flush())
    assert product_chain_desc(h) == oracle_chain_desc(p)
    # the stream continues after a flush without a reset
    for n in (5000, 17):
        assert h.advance_geometry(n) == len(p.process(np.zeros(n)))
    assert h.advance_geometry(0, flush=True) == len(p.flush())
    assert product_chain_desc(h) == oracle_chain_desc(p)


@pytest.mark.parametrize("ir,orr,preset,dt", ENGINE_CASES)
def test_engine_geometry_matches_oracle(ir, orr, preset, dt):
    h = G.SimpleResampler(ir, orr, preset, dt, device=-1)
    e = O.Engine(ir, orr, O.preset_to_engine_quality(preset), dt)
    assert product_chain_desc(h) == flatten_engine_desc(e.describe())
    rng = np.random.default_rng(11)
    for n in _chunks(rng, 60000) + [0, 3]:
        assert h.advance_geometry(n) == len(e.process(np.zeros(n, dtype=dt)))
        assert product_chain_desc(h) == flatten_engine_desc(e.describe())
    assert h.GetStatistics() == dict(zip(("samplesIn", "samplesOut"), e.stats()))
    assert h.advance_geometry(0, flush=True) == len(e.flush())
    assert product_chain_desc(h) == flatten_engine_desc(e.describe())
    assert h.GetStatistics() == dict(zip(("samplesIn", "samplesOut"), e.stats()))


def test_baseline_sample_counts():
    """BASELINE.md §3 / SURVEY §8a counts, asserted on the product's own state machine."""
    h = G.SimpleResampler(44100, 48000, G.QualityHigh, np.float64, device=-1)
    assert (h.advance_geometry(441000), h.advance_geometry(0, True)) == (479787, 215)
    h = G.New(geometry_config(48000, 44100))
    counts = [h.advance_geometry(min(4096, 480000 - i)) for i in range(0, 480000, 4096)]
    assert counts[:5] == [3556, 3763, 3763, 3764, 3763] and sum(counts) == 440793
    assert h.advance_geometry(0, True) == 209 and h.EstimateOutput(4096) == 3827
    h = G.New(geometry_config(96000, 48000, 8, G.QualityVeryHigh))
    assert (h.advance_geometry(960000), h.advance_geometry(0, True)) == (479389, 612)
    for preset, a, b in ((G.QualityMedium, 159708, 293), (G.QualityHigh, 159626, 375)):
        h = G.SimpleResampler(48000, 16000, preset, np.float32, device=-1)
        assert (h.advance_geometry(480000), h.advance_geometry(0, True)) == (a, b)
    h = G.New(geometry_config(8000, 192000))
    assert (h.advance_geometry(80000), h.advance_geometry(0, True)) == (1910673, 9375)
    h = G.SimpleResampler(44100, 47999, G.QualityHigh, np.float64, device=-1)
    assert (h.advance_geometry(441000), h.advance_geometry(0, True)) == (479789, 203)


@pytest.mark.parametrize("ir,orr,preset,dt", ENGINE_CASES)
def test_banks_bit_identical_to_oracle(ir, orr, preset, dt):
    h = G.SimpleResampler(ir, orr, preset, dt, device=-1)
    e = O.Engine(ir, orr, O.preset_to_engine_quality(preset), dt)
    for s, d in enumerate(product_chain_desc(h)):
        if d["kind"] == "up":
            np.testing.assert_array_equal(h.bank(s, 0), e.bank(0))
        elif d["kind"] == "dec":
            np.testing.assert_array_equal(h.bank(s, 0), e.bank(1))
        elif d["kind"] == "poly":
            for w in range(4):
                np.testing.assert_array_equal(h.bank(s, w), e.bank(2 + w))


def test_no_state_change_when_peeking():
    h = G.SimpleResampler(44100, 48000, G.QualityHigh, np.float64, device=-1)
    h.advance_geometry(1000)
    before = product_chain_desc(h)
    G.lib().gar_next_output_count(h._h, 0, 4096)
    G.lib().gar_next_flush_count(h._h, 0)
    assert product_chain_desc(h) == before


def test_reset_clears_state():
    h = G.New(geometry_config(48000, 44100))
    fresh = product_chain_desc(h)
    first = h.advance_geometry(4096)
    h.Reset()
    assert product_chain_desc(h) == fresh
    assert h.advance_geometry(4096) == first


@pytest.mark.parametrize("ir,orr,preset", [(44100, 48000, G.QualityHigh), (48000, 44100, G.QualityHigh),
                                          (44100, 47999, G.QualityHigh), (8000, 192000, G.QualityHigh),
                                          (96000, 48000, G.QualityVeryHigh), (48000, 16000, G.QualityHigh),
                                          (44100, 16000, G.QualityMedium), (8000, 22050, G.QualityLow)])
def test_time_sliced_process_produces_the_one_shot_counts_and_state(ir, orr, preset):
    """The engine runs a long multi-stage Process call as a sequence of shorter ones (inter-stage buffers stay
    L2-resident). Every stage is greedy, so the summed counts, the carried state and the flush length are those of
    the single call — checked here on the integer state machine (geometry-only handles, no device)."""
    n = 1_000_003
    for make in (lambda: G.New(geometry_config(ir, orr, 1, preset)),
                 lambda: G.SimpleResampler(ir, orr, preset, np.float64, device=-1)):
        one, cut = make(), make()
        total = one.advance_geometry(n)
        got, off = 0, 0
        for slice_len in (16384, 20480, 65536, 12288):
            while off < n and (slice_len != 12288 or True):
                k = min(slice_len, n - off)
                got += cut.advance_geometry(k)
                off += k
                if slice_len != 12288 and off > n // 2:
                    break
            if off >= n:
                break
        assert off == n and got == total
        assert product_chain_desc(one) == product_chain_desc(cut)
        assert one.advance_geometry(0, flush=True) == cut.advance_geometry(0, flush=True)
