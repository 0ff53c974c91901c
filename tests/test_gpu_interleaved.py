"""SURVEY.md §8f N1 — interleaved / integer-PCM boundary fused around the resampling pass, checked against a
numpy restatement of cmd/resample-wav/main.go:358-520 driving the oracle engine per channel
(cmd/resample-wav/helpers.go:77-334: one engine.Resampler per channel, block loop, flush + pad + interleave)."""
import numpy as np
import pytest

from helpers import G, O

pytestmark = pytest.mark.gpu

MAXV = {16: 32767.0, 24: 8388607.0, 32: 2147483647.0}


def ref_deinterleave(data, channels, bit_depth, F):  # main.go:444-470: F(float64(v) * invMaxVal)
    inv = 1.0 / MAXV.get(bit_depth, 32767.0)
    x = data.reshape(-1, channels).astype(np.float64) * inv
    return [np.ascontiguousarray(x[:, c]).astype(F) for c in range(channels)]


def ref_interleave(chans, bit_depth):  # main.go:474-520: clamp, int(sample * maxVal) (truncation)
    mv = MAXV.get(bit_depth, 32767.0)
    cols = [np.trunc(np.clip(c.astype(np.float64), -1.0, 1.0) * mv).astype(np.int64) for c in chans]
    return np.stack(cols, axis=1)


def pcm_signal(n, channels, bit_depth, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 44100.0
    cols = [0.6 * np.sin(2 * np.pi * (300 + 211 * c) * t) + 0.3 * (rng.random(n) - 0.5) for c in range(channels)]
    x = np.stack(cols, axis=1)
    x[n // 2, 0] = 1.5  # exercises the clamp on the way out for ratios near 1
    return np.trunc(np.clip(x, -1, 1) * MAXV[bit_depth]).astype(np.int64)


@pytest.mark.parametrize("channels,bit_depth,dt_int,F,ir,orr", [
    (2, 16, np.int16, np.float64, 44100, 48000),
    (2, 24, np.int32, np.float64, 48000, 44100),
    (6, 32, np.int64, np.float64, 96000, 48000),
    (1, 16, np.int64, np.float32, 48000, 16000),
    (2, 24, np.int32, np.float32, 44100, 48000)])
def test_pcm_interleaved_block_loop_matches_reference_restatement(channels, bit_depth, dt_int, F, ir, orr):
    n, block = 30000, 4096
    pcm = pcm_signal(n, channels, bit_depth, 3).astype(dt_int)
    h = G.SimpleResampler(ir, orr, G.QualityHigh, F, n_streams=channels)
    engines = [O.Engine(ir, orr, O.Q_HIGH, F) for _ in range(channels)]
    # 1 LSB of slack where a float result sits within rounding distance of an integer boundary
    tol_lsb = 1
    worst = 0
    for i in range(0, n, block):
        blk = pcm[i:i + block]
        got = h.ProcessInterleaved(blk, bit_depth)
        chans = ref_deinterleave(blk.astype(np.int64), channels, bit_depth, F)
        res = [e.process(c) for e, c in zip(engines, chans)]
        want = ref_interleave(res, bit_depth)
        assert got.shape == want.shape and got.dtype == dt_int
        if len(want):
            worst = max(worst, int(np.max(np.abs(got.astype(np.int64) - want))))
    gf = h.FlushInterleaved(dt_int, bit_depth)
    wf = ref_interleave([e.flush() for e in engines], bit_depth)
    assert gf.shape == wf.shape
    worst = max(worst, int(np.max(np.abs(gf.astype(np.int64) - wf))))
    # float64 engines differ by <= 1e-12 (<< 1 LSB even at 32 bit: 2e-3 LSB); float32 by <= 1e-6 (2147 LSB at 32 bit)
    limit = tol_lsb if F == np.float64 else max(tol_lsb, int(1e-6 * MAXV[bit_depth]) + 1)
    assert worst <= limit, worst


def test_conversion_kernels_are_bit_exact_at_ratio_one():
    """Rate 1:1 is a pass-through engine (DFTStage(1)), so the output is exactly the conversion round trip."""
    pcm = pcm_signal(5000, 2, 24, 9).astype(np.int32)
    h = G.SimpleResampler(48000, 48000, G.QualityHigh, np.float64, n_streams=2)
    got = h.ProcessInterleaved(pcm, 24)
    want = ref_interleave(ref_deinterleave(pcm.astype(np.int64), 2, 24, np.float64), 24)
    np.testing.assert_array_equal(got.astype(np.int64), want)
    h32 = G.SimpleResampler(48000, 48000, G.QualityHigh, np.float32, n_streams=2)
    got32 = h32.ProcessInterleaved(pcm, 24)
    want32 = ref_interleave(ref_deinterleave(pcm.astype(np.int64), 2, 24, np.float32), 24)
    np.testing.assert_array_equal(got32.astype(np.int64), want32)


@pytest.mark.parametrize("F", [np.float64, np.float32])
def test_float_interleaved_stereo_equals_planar_path(F):  # convenience.go:261-282,463-486
    rng = np.random.default_rng(4)
    l, r = (rng.random(20000) - 0.5).astype(F), (rng.random(20000) - 0.5).astype(F)
    h = G.SimpleResampler(44100, 48000, G.QualityMedium, F, n_streams=2)
    il = G.InterleaveToStereo(l, r)
    got = np.concatenate([h.ProcessInterleaved(il), h.FlushInterleaved(F)])
    lo, ro = (G.ResampleMono if F == np.float64 else G.ResampleMonoFloat32)(l, 44100, 48000, G.QualityMedium), \
        (G.ResampleMono if F == np.float64 else G.ResampleMonoFloat32)(r, 44100, 48000, G.QualityMedium)
    gl, gr = G.DeinterleaveFromStereo(got.reshape(-1))
    np.testing.assert_array_equal(gl, lo)
    np.testing.assert_array_equal(gr, ro)
