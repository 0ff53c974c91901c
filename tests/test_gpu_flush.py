"""Flush (engine.Resampler.Flush, resampler.go:275-322; the pipeline's cascade, constant.go:360-386) is a sequence of small stage
calls. With fusion on, an x2 stage call and the polyphase stage call behind it run as one launch (K4 family); with fusion off
every stage call is its own launch. Same samples either way (bit for bit), same counts as the oracle, state usable afterwards.
(A whole Flush as ONE launch — one block per row running the stage calls in order — was measured and dropped: 18 against 12 us
for config 1, 350 against 80 us for the 10-call cascade of config 5a, whose later calls are large enough to want many SMs.)"""
import numpy as np
import pytest

from helpers import G, O

pytestmark = pytest.mark.gpu


def _cfg(ir, orr, preset=G.QualityHigh):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=preset))


@pytest.mark.parametrize("ir,orr,rows,n", [
    (44100, 48000, 1, 30000),     # x2 + polyphase engine: 3 stage calls
    (44100, 47999, 2, 20000),     # cubic coefficient interpolation live
    (8000, 192000, 1, 9000),      # five x2 stages + polyphase: the longest flush cascade of the BASELINE configs
    (48000, 44100, 3, 25000),     # pre-stage + polyphase, three lock-step rows
    (96000, 48000, 2, 40000),     # a single decimator: one op, stays on the ordinary launch
    (48000, 8000, 1, 50000),      # two-engine pipeline going down
])
def test_fused_flush_equals_one_launch_per_stage_call_and_the_oracle(ir, orr, rows, n):
    rng = np.random.default_rng(11)
    x = 0.5 * rng.standard_normal((rows, n))

    def run(fused):
        h = G.Resampler(_cfg(ir, orr), n_streams=rows)
        h.set_fusion(fused)
        y = h.ProcessBatch(x)[0].copy()
        G.kernel_launches(reset=True)
        f = h.FlushBatch()[0].copy()
        lf = G.kernel_launches()
        # the state after a Flush is still usable (resampler.go allows Process after Flush): same samples either way
        y2 = h.ProcessBatch(x[:, :5000])[0].copy()
        f2 = h.FlushBatch()[0].copy()
        return y, f, y2, f2, lf, h.last_kernels()

    ya, fa, ya2, fa2, la, ka = run(True)
    yb, fb, yb2, fb2, lb, kb = run(False)
    assert np.array_equal(ya, yb) and np.array_equal(ya2, yb2)
    assert fa.shape == fb.shape and fa2.shape == fb2.shape
    assert np.array_equal(fa, fb), float(np.max(np.abs(fa - fb)))
    assert np.array_equal(fa2, fb2), float(np.max(np.abs(fa2 - fb2)))
    assert la <= lb, (la, lb)
    p = O.Pipeline(ir, orr, 1, O.PRESET_HIGH)
    want = np.concatenate([p.process(x[0]), p.flush()])
    got = np.concatenate([ya[0], fa[0]])
    assert len(got) == len(want)
    assert np.max(np.abs(got - want)) <= 1e-12
