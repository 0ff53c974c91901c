"""K5 — the persistent x2 -> polyphase chain kernel (kernels_chain.cu): one launch per Process for large float64 batches, the
intermediate-rate samples in an L2-resident ring. Same MMA cores as the two stand-alone launches (K1m + K3p): every sample
must be bit-identical to them, and within 1e-12 of the oracle's restatement of resampler.go:182-227 / polyphase_stage.go:186-312."""
import numpy as np
import pytest

from helpers import G, O

pytestmark = pytest.mark.gpu


def _run(ir, orr, x, cuts, chain):
    G.set_chain_kernel(chain)
    try:
        h = G.NewBatch(ir, orr, G.QualityHigh, x.shape[0], np.float64)
        ys, ks = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            ys.append(h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy())
            ks.append(h.last_kernels())
        ys.append(h.FlushBatch()[0].copy())
        return np.concatenate(ys, axis=1), h.last_kernels(), G.kernel_launches()
    finally:
        G.set_chain_kernel(True)


@pytest.mark.parametrize("ir,orr,rows,n,cuts", [
    (44100, 48000, 256, 100000, None),              # rational 147/80, 16-row stages, ring wraps ~3 times
    (44100, 47999, 40, 150000, None),               # cubic coefficient interpolation live
    (48000, 44100, 70, 120000, None),               # 2.18 samples per output; ragged last groups (70 = 8*8 + 6 = 2*32 + 6)
    (44100, 48000, 64, 200001, [0, 90001, 200001]), # two large calls: the second starts from carried tails and a non-zero phase
    (44100, 48000, 33, 260000, [0, 1000, 260000]),  # a small call (stand-alone kernels) in front of a chain call
])
def test_chain_kernel_bit_identical_to_stand_alone_launches_and_oracle(ir, orr, rows, n, cuts):
    rng = np.random.default_rng(5)
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = cuts or [0, n]
    G.kernel_launches(reset=True)
    ya, ka, la = _run(ir, orr, x, cuts, True)
    G.kernel_launches(reset=True)
    yb, kb, lb = _run(ir, orr, x, cuts, False)
    assert "chain_up2_poly_f64_mma" in ka, ka
    assert "chain_up2_poly_f64_mma" not in kb and "fir_f64_mma_up2" in kb, kb
    assert la < lb, (la, lb)  # one launch instead of two per large Process call
    assert ya.shape == yb.shape
    assert np.array_equal(ya, yb), float(np.max(np.abs(ya - yb)))
    pick = sorted(set([0, rows // 2, rows - 1]))
    want, counts = O.batch_resample(x[pick], ir, orr, O.Q_HIGH, n_threads=4)
    assert np.all(counts == ya.shape[1])
    assert np.max(np.abs(ya[pick] - want[:, :ya.shape[1]])) <= 1e-12


def test_chain_kernel_keeps_streaming_state():
    """A chain call followed by small calls and a flush: the carried tails written by the chain kernel (x2 stage from the
    input, polyphase stage from the ring) are what the stand-alone kernels would have left."""
    rng = np.random.default_rng(6)
    rows, n = 48, 140000
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = [0, 120000, 120300, 125000, n]
    ya, ka, _ = _run(44100, 48000, x, cuts, True)
    yb, _, _ = _run(44100, 48000, x, cuts, False)
    assert "chain_up2_poly_f64_mma" in ka, ka
    assert np.array_equal(ya, yb)
