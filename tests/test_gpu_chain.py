"""K5 — the persistent x2 -> polyphase chain kernel (kernels_chain.cu): one launch per Process for large float64 batches, the
intermediate-rate samples in an L2-resident ring instead of a full-size device buffer. Same MMA cores as the two stand-alone
launches (K1m + K3p): every sample must be bit-identical to them, and within 1e-12 of the oracle's restatement of
resampler.go:182-227 / polyphase_stage.go:186-312. Calls go through the C ABI's device entry points (gar_process_batch_dev)."""
import numpy as np
import pytest
import torch

from helpers import G, O

pytestmark = pytest.mark.gpu


def _run(ir, orr, x, cuts, mode, budget=None):
    """Process the column ranges `cuts` of x as successive batch calls on device buffers, then Flush."""
    G.set_chain_kernel(mode)
    try:
        dev = torch.device("cuda", 0)
        rows = x.shape[0]
        h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
        if budget is not None:
            h.set_slice_budget(budget)
        G.kernel_launches(reset=True)
        ys = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            dx = torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).to(dev)
            n = hi - lo
            cap = (h.EstimateOutput(n) + 64 + 3) & ~3
            dy = torch.zeros((rows, cap), dtype=torch.float64, device=dev)
            k = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), cap, cap, 0, np.float64)
            torch.cuda.synchronize()
            ys.append(dy[:, :k].cpu().numpy())
        ys.append(h.FlushBatch()[0].copy())
        return np.concatenate(ys, axis=1), h.last_kernels(), G.kernel_launches(), h
    finally:
        G.set_chain_kernel(2)


@pytest.mark.parametrize("ir,orr,rows,n,cuts", [
    (44100, 48000, 256, 60000, None),               # rational 147/80; the ring wraps several times
    (44100, 47999, 40, 150000, None),               # cubic coefficient interpolation live; 40 rows = 2 full 16-row stages + 8
    (48000, 44100, 70, 120000, None),               # 2.18 samples per output; ragged last groups (70 = 8*8 + 6 = 4*16 + 6)
    (44100, 48000, 64, 200001, [0, 90001, 200001]), # two large calls: the second starts from carried tails and a non-zero phase
    (44100, 48000, 33, 260000, [0, 1000, 260000]),  # a small call (stand-alone kernels) in front of a chain call
])
def test_chain_kernel_bit_identical_to_stand_alone_launches_and_oracle(ir, orr, rows, n, cuts):
    rng = np.random.default_rng(5)
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = cuts or [0, n]
    ya, ka, la, _ = _run(ir, orr, x, cuts, 1)
    yb, kb, lb, _ = _run(ir, orr, x, cuts, 0)
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
    ya, ka, _, _ = _run(44100, 48000, x, cuts, 1)
    yb, _, _, _ = _run(44100, 48000, x, cuts, 0)
    assert "chain_up2_poly_f64_mma" in ka, ka
    assert np.array_equal(ya, yb)


def test_chain_kernel_replaces_time_slices_above_the_memory_budget():
    """Default policy (mode 2): a call over all rows whose intermediate buffer exceeds the inter-stage budget by so much that time
    slices would be shorter than 64 K samples runs as ONE chain launch without that buffer instead; below the budget (and with
    longer slices) the two stand-alone launches stay. Identical samples either way."""
    rng = np.random.default_rng(7)
    rows, n = 64, 200000
    x = 0.5 * rng.standard_normal((rows, n))
    budget = 32 << 20                                  # the intermediate buffer would be 64 x 400k x 8 = 205 MB: 32 K-sample slices
    ya, ka, la, ha = _run(44100, 48000, x, [0, n], 2, budget)
    yb, kb, lb, hb = _run(44100, 48000, x, [0, n], 0, budget)   # time slices
    yc, kc, _, hc = _run(44100, 48000, x, [0, n], 2)            # default budget (2 GiB): not exceeded
    assert "chain_up2_poly_f64_mma" in ka, ka
    assert "chain_up2_poly_f64_mma" not in kb and "chain_up2_poly_f64_mma" not in kc, (kb, kc)
    assert la < lb, (la, lb)
    assert np.array_equal(ya, yb) and np.array_equal(ya, yc)
    # no full-size intermediate buffer (205 MB here): the ring and its counters instead (<= 45 MB)
    ma, mc = ha.GetInfo()["MemoryUsage"], hc.GetInfo()["MemoryUsage"]
    assert ma + (150 << 20) < mc, (ma, mc)


@pytest.mark.parametrize("ir,orr,rows,n,make", [
    (48000, 44100, 64, 90000, "pipeline"),    # BASELINE config 2's shape: float64 pipeline behind the float32 API
    (44100, 48000, 40, 120001, "engine32"),   # NewEngineFloat32 at a non-integer ratio (float64 tensor-core arithmetic inside)
    (44100, 47999, 33, 100000, "pipeline"),   # interpolated coefficients, ragged row groups
    (8000, 192000, 32, 70000, "pipeline"),    # five x2 stages + polyphase: only the first and the last launch touch float32
    (96000, 48000, 16, 140000, "pipeline"),   # one /2 decimator (K2m): float32 in and float32 out in the same launch
    (48000, 16000, 40, 100001, "pipeline"),   # /2 -> x2 -> polyphase: float32 in on the decimator (K2m), float32 out of K3p
    (48000, 8000, 33, 120000, "pipeline"),    # two engines going down
    (192000, 48000, 9, 150001, "pipeline"),   # /4 decimator, ragged column group (9 rows)
])
def test_float32_io_folded_into_the_tensor_core_pair_equals_the_cast_launches(ir, orr, rows, n, make):
    """Large float32-I/O batches on float64 arithmetic: K1m widens its float32 sample windows in shared memory and K3p narrows on
    the store (no cast launches, no float64 copies of input and output in HBM). Same samples as casting through scratch buffers
    (gar_set_fusion(h, 0) keeps the casts), and within 1e-6 of the oracle."""
    rng = np.random.default_rng(8)
    x = (0.5 * rng.standard_normal((rows, n))).astype(np.float32)
    dev = torch.device("cuda", 0)

    def run(fold):
        if make == "pipeline":
            h = G.Resampler(G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=G.QualityHigh)),
                            n_streams=rows)
        else:
            h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float32)
        h.set_fusion(fold)
        ys = []
        for lo, hi in [(0, n // 2 + 3), (n // 2 + 3, n)]:
            dx = torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).to(dev)
            m = hi - lo
            istride = (m + 3) & ~3
            dxp = torch.zeros((rows, istride), dtype=torch.float32, device=dev)
            dxp[:, :m] = dx
            cap = (h.EstimateOutput(m) + 64 + 3) & ~3
            dy = torch.zeros((rows, cap), dtype=torch.float32, device=dev)
            G.kernel_launches(reset=True)
            k = h.process_batch_dev(dxp.data_ptr(), istride, m, dy.data_ptr(), cap, cap, 0, np.float32)
            torch.cuda.synchronize()
            ys.append(dy[:, :k].cpu().numpy())
        launches = G.kernel_launches()
        ys.append(h.FlushBatch(io_dtype=np.float32)[0].copy() if make == "pipeline" else h.FlushBatch()[0].copy())
        return np.concatenate(ys, axis=1), launches, h.last_kernels()

    ya, la, ka = run(True)
    yb, lb, kb = run(False)
    assert la <= lb - 2 and la >= 1, (la, lb, ka, kb)  # no cast launch in front, none behind
    assert ya.dtype == np.float32 and ya.shape == yb.shape
    assert np.array_equal(ya, yb), float(np.max(np.abs(ya.astype(np.float64) - yb)))
    if make == "pipeline":
        p = O.Pipeline(ir, orr, 1, O.PRESET_HIGH)
        want = np.concatenate([p.process(x[0].astype(np.float64)), p.flush()])
        assert len(want) == ya.shape[1]
        assert np.max(np.abs(ya[0].astype(np.float64) - want)) <= 1e-6
