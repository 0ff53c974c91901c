"""API-level invariants the reference tests assert, re-run on the GPU engine, plus the
edge cases of its engine tests (empty / tiny inputs, flush of an unfed stage, reset, unusual factors)."""
import numpy as np
import pytest

from helpers import G, O, sig_c3

pytestmark = pytest.mark.gpu


def _cfg(ir, orr, ch=1, preset=G.QualityHigh, precision=0):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch,
                    Quality=G.QualitySpec(Preset=preset, Precision=precision, PhaseResponse=50, PassbandEnd=0.9,
                                          StopbandBegin=0.99))


def _noise(n, seed=0, dt=np.float64):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    return (0.6 * np.sin(2 * np.pi * 0.01 * t) + 0.3 * (rng.random(n) - 0.5)).astype(dt)


RATES = [(44100, 48000), (48000, 44100), (96000, 48000), (48000, 32000), (22050, 44100), (48000, 16000),
         (8000, 192000), (48000, 48000)]


@pytest.mark.parametrize("ir,orr", RATES)
def test_process_into_equals_process_and_oracle(ir, orr):  # processinto_test.go:36-105,562-618
    x = _noise(30000, 1)
    a, b = G.New(_cfg(ir, orr)), G.New(_cfg(ir, orr))
    p = O.Pipeline(ir, orr, 1, O.PRESET_HIGH)
    out = np.empty(a.EstimateOutput(5000))
    for i in range(0, len(x), 5000):
        n = a.ProcessInto(x[i:i + 5000], out)
        y = b.Process(x[i:i + 5000])
        w = p.process(x[i:i + 5000])
        np.testing.assert_array_equal(out[:n], y)
        assert len(y) == len(w)
        if len(w):
            assert np.max(np.abs(y - w)) <= 1e-12
    fa, fb, fw = a.Flush(), b.Flush(), p.flush()
    np.testing.assert_array_equal(fa, fb)
    assert len(fa) == len(fw) and (len(fw) == 0 or np.max(np.abs(fa - fw)) <= 1e-12)


@pytest.mark.parametrize("ir,orr,dt", [(44100, 48000, np.float64), (48000, 44100, np.float32),
                                       (48000, 16000, np.float32), (44100, 47999, np.float64),
                                       (48000, 8000, np.float64), (8000, 192000, np.float64),
                                       (48000, 11025, np.float64), (16000, 48000, np.float32),
                                       (12000, 48000, np.float64)])
def test_engine_random_chunking_matches_oracle(ir, orr, dt):  # processinto_test.go:258-449
    tol = 1e-12 if dt == np.float64 else 1e-6
    x = _noise(50000, 2, dt)
    r = G.SimpleResampler(ir, orr, G.QualityHigh, dt)
    e = O.Engine(ir, orr, O.Q_HIGH, dt)
    rng = np.random.default_rng(3)
    i = 0
    while i < len(x):
        n = int(rng.integers(1, 7000))
        y, w = r.Process(x[i:i + n]), e.process(x[i:i + n])
        assert len(y) == len(w) <= r.EstimateOutput(len(x[i:i + n]))
        if len(w):
            assert np.max(np.abs(y.astype(np.float64) - w.astype(np.float64))) <= tol
        i += n
    fy, fw = r.Flush(), e.flush()
    assert len(fy) == len(fw)
    if len(fw):
        assert np.max(np.abs(fy.astype(np.float64) - fw.astype(np.float64))) <= tol
    assert r.GetStatistics() == dict(zip(("samplesIn", "samplesOut"), e.stats()))


def test_buffer_too_small_does_not_advance_state():  # processinto_test.go:176-224
    x = _noise(8192, 4)
    a, b = G.New(_cfg(44100, 48000)), G.New(_cfg(44100, 48000))
    a.Process(x[:4096])
    b.Process(x[:4096])
    with pytest.raises(G.ErrBufferTooSmall):
        a.ProcessInto(x[4096:], np.empty(100))
    np.testing.assert_array_equal(a.Process(x[4096:]), b.Process(x[4096:]))


def test_parallel_equals_sequential_and_channels_independent():  # parallel_test.go:12-89
    xs = [c[:60000] for c in sig_c3(60000, 4)]
    m = G.New(_cfg(96000, 48000, 4, G.QualityVeryHigh))
    ys = m.ProcessMulti(xs)
    fs = m.FlushMulti()
    for c in range(4):
        s = G.New(_cfg(96000, 48000, 1, G.QualityVeryHigh))
        np.testing.assert_array_equal(ys[c], s.Process(xs[c]))
        np.testing.assert_array_equal(fs[c], s.Flush())


def test_process_multi_unequal_channel_lengths_and_flush_is_channel0_only():
    xs = [_noise(20000, 5), _noise(12345, 6), _noise(20000, 7)]
    m = G.New(_cfg(44100, 48000, 3))
    p = O.Pipeline(44100, 48000, 3, O.PRESET_HIGH)
    ys, ws = m.ProcessMulti(xs), p.process_multi(xs)
    for y, w in zip(ys, ws):
        assert len(y) == len(w) and np.max(np.abs(y - w)) <= 1e-12
    f0, w0 = m.Flush(), p.flush(0)  # Flush drains channel 0 only (constant.go:349-354)
    assert len(f0) == len(w0) and np.max(np.abs(f0 - w0)) <= 1e-12
    fs, wf = m.FlushMulti(), p.flush_multi()
    for y, w in zip(fs, wf):
        assert len(y) == len(w)
        if len(w):
            assert np.max(np.abs(y - w)) <= 1e-12


def test_stereo_equals_two_mono():  # convenience_stereo_test.go:40-71
    l, r = _noise(20000, 8), _noise(20000, 9)
    lo, ro = G.ResampleStereo(l, r, 44100, 48000, G.QualityHigh)
    np.testing.assert_array_equal(lo, G.ResampleMono(l, 44100, 48000, G.QualityHigh))
    np.testing.assert_array_equal(ro, G.ResampleMono(r, 44100, 48000, G.QualityHigh))
    l32, r32 = l.astype(np.float32), r.astype(np.float32)
    lo, ro = G.ResampleStereoFloat32(l32, r32, 48000, 44100, G.QualityMedium)
    np.testing.assert_array_equal(ro, G.ResampleMonoFloat32(r32, 48000, 44100, G.QualityMedium))
    assert lo.dtype == np.float32


def test_float32_vs_float64_within_1e5():  # convenience_float32_test.go:222-266
    x = _noise(30000, 10)
    a = G.ResampleMono(x, 44100, 48000, G.QualityHigh)
    b = G.ResampleMonoFloat32(x.astype(np.float32), 44100, 48000, G.QualityHigh)
    assert len(a) == len(b) and np.max(np.abs(a - b)) <= 1e-5


def test_empty_tiny_inputs_and_flush_of_unfed_stage():  # edge_cases_test.go, polyphase_flush_test.go:94-109
    for ir, orr in RATES:
        r = G.New(_cfg(ir, orr))
        p = O.Pipeline(ir, orr, 1, O.PRESET_HIGH)
        assert len(r.Flush()) == len(p.flush())  # nothing fed: nothing (or passthrough nothing) to drain
        assert len(r.Process(np.zeros(0))) == 0
        for n in (1, 1, 2, 7):
            y, w = r.Process(np.full(n, 0.25)), p.process(np.full(n, 0.25))
            assert len(y) == len(w)
        fy, fw = r.Flush(), p.flush()
        assert len(fy) == len(fw)
        if len(fw):
            assert np.max(np.abs(fy - fw)) <= 1e-12


def test_reset_reproduces_first_run():  # reset_state_test.go
    x = _noise(20000, 11)
    r = G.NewEngine(48000, 44100, G.QualityHigh)
    a = np.concatenate([r.Process(x), r.Flush()])
    r.Reset()
    assert r.GetStatistics() == {"samplesIn": 0, "samplesOut": 0}
    b = np.concatenate([r.Process(x), r.Flush()])
    np.testing.assert_array_equal(a, b)


def test_zero_input_gives_zero_output_and_dc_gain():  # regression_test.go:160-185, quality_regression_test.go:26-55
    r = G.NewEngine(44100, 48000, G.QualityHigh)
    assert np.max(np.abs(r.Process(np.zeros(20000)))) <= 1e-10
    r.Reset()
    y = r.Process(np.ones(40000))
    assert abs(float(np.mean(y[5000:-5000])) - 1.0) <= 1e-3


@pytest.mark.parametrize("preset,prec", [(G.QualityQuick, 0), (G.QualityLow, 0), (G.QualityMedium, 0),
                                         (G.QualityVeryHigh, 0), (G.QualityCustom, 20), (G.QualityCustom, 28)])
def test_quality_presets_path_a(preset, prec):  # includes the QualityQuick cubic stage (cubic.go)
    x = _noise(25000, 12)
    for ir, orr in ((44100, 48000), (48000, 22050)):
        r = G.New(_cfg(ir, orr, 1, preset, prec))
        p = O.Pipeline(ir, orr, 1, preset, prec)
        outs = 0
        for i in range(0, len(x), 6000):
            y, w = r.Process(x[i:i + 6000]), p.process(x[i:i + 6000])
            assert len(y) == len(w)
            tol = 0.0 if preset == G.QualityQuick else 1e-12  # the cubic path is evaluated op for op
            if len(w):
                assert np.max(np.abs(y - w)) <= tol
            outs += len(y)
        fy, fw = r.Flush(), p.flush()
        assert len(fy) == len(fw)
        assert outs > 0


def test_uploaded_bank_replaces_host_design():
    """gar_upload_bank: coefficients designed elsewhere (e.g. by the Go internal/filter) drive the kernels."""
    x = _noise(10000, 13)
    r = G.NewEngine(48000, 24000, G.QualityLow)
    base = r.Process(x)
    r.Reset()
    bank = r.bank(0, 0)
    r.upload_bank(0, 0, bank * 0.5)
    half = r.Process(x)
    assert np.max(np.abs(half - 0.5 * base)) <= 1e-15


@pytest.mark.parametrize("ir,orr,dt", [(44100, 48000, np.float64), (48000, 44100, np.float64),
                                       (44100, 47999, np.float64), (8000, 12000, np.float64),
                                       (48000, 44100, np.float32), (44100, 48000, np.float32)])
def test_fused_x2_polyphase_kernel_equals_unfused(ir, orr, dt):
    """K4: the fused launch keeps the intermediate-rate samples in shared memory; float64 results are
    bit-identical to the two stand-alone launches (same summation order), float32 to rounding."""
    x = _noise(60000, 21, dt)
    a = G.SimpleResampler(ir, orr, G.QualityHigh, dt)
    b = G.SimpleResampler(ir, orr, G.QualityHigh, dt)
    b.set_fusion(False)
    rng = np.random.default_rng(5)
    i, la, lb = 0, 0, 0
    G.kernel_launches(reset=True)
    while i < len(x):
        n = int(rng.integers(1, 9000))
        k0 = G.kernel_launches()
        ya = a.Process(x[i:i + n])
        k1 = G.kernel_launches()
        yb = b.Process(x[i:i + n])
        k2 = G.kernel_launches()
        la, lb = la + (k1 - k0), lb + (k2 - k1)
        if dt == np.float64:
            np.testing.assert_array_equal(ya, yb)
        else:
            assert len(ya) == len(yb) and (len(ya) == 0 or np.max(np.abs(ya - yb)) <= 2e-7)
        i += n
    fa, fb = a.Flush(), b.Flush()
    if dt == np.float64:
        np.testing.assert_array_equal(fa, fb)
    else:
        assert len(fa) == len(fb) and np.max(np.abs(fa - fb)) <= 2e-7
    assert la < lb  # fewer launches when fused


@pytest.fixture
def vector_fir_only():
    """The bit-for-bit comparisons below are between kernels that sum taps strictly in order; the FP64 tensor-core FIR
    (taps grouped in fours) is switched off for them and has its own test."""
    G.set_tensor_fir(False)
    yield
    G.set_tensor_fir(True)


@pytest.mark.parametrize("ir,orr,rows,n,kernel", [
    (44100, 48000, 200, 60000, "fused_up2_rat_f64"),   # Mi/L = 147/80  -> slot stride 2, persistent multi-tile blocks
    (48000, 44100, 40, 70000, "fused_up2_rat_f64"),    # 320/147 -> slot stride 3
    (48000, 30000, 24, 50000, "fused_up2_rat_f64"),    # 3.2     -> slot stride 4
    (8000, 12000, 33, 20000, "fused_up2_rat_f64"),     # 4/3
    (44100, 32000, 16, 41000, "fused_up2_rat_f64"),
])
def test_rational_fused_kernel_batched_rows_vs_unfused_and_oracle(ir, orr, rows, n, kernel, vector_fir_only):
    """K4r (register-tiled rational-ratio fused kernel): many lock-step rows, ragged chunking (every carried
    phase/tail state), flush. Bit-identical to the stand-alone launches; <= 1e-12 against the oracle."""
    rng = np.random.default_rng(77)
    x = (0.5 * np.sin(2 * np.pi * rng.random((rows, 1)) * 0.05 * np.arange(n)[None, :]) +
         0.4 * (rng.random((rows, n)) - 0.5))
    a = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
    b = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
    b.set_fusion(False)
    cuts = [0, 7, 7 + 311, n // 3, n - 1234, n]
    ya, yb = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        xa = np.ascontiguousarray(x[:, lo:hi])
        ya.append(a.ProcessBatch(xa)[0].copy())
        yb.append(b.ProcessBatch(xa)[0].copy())
    ya.append(a.FlushBatch()[0].copy())
    yb.append(b.FlushBatch()[0].copy())
    assert kernel in a.last_kernels(), a.last_kernels()
    # the stand-alone stage runs K3r (odd period length) or K3i (even period length, >= 8 rows)
    assert any(k in b.last_kernels() for k in ("poly_rat_f64", "poly_rows_f64")), b.last_kernels()
    ya, yb = np.concatenate(ya, axis=1), np.concatenate(yb, axis=1)
    np.testing.assert_array_equal(ya, yb)
    pick = sorted(set([0, 1, rows // 2, rows - 1]))
    want, counts = O.batch_resample(x[pick], ir, orr, O.Q_HIGH, n_threads=4)
    assert np.all(counts == ya.shape[1])
    assert np.max(np.abs(ya[pick] - want[:, :ya.shape[1]])) <= 1e-12


def test_rational_kernels_random_geometry_stress(vector_fir_only):
    """Many random (rows, length, chunking) cases through the barrier-free K4r / K3r pipelines: fused, unfused and
    one-shot runs must agree bit for bit (float64 sums are strictly sequential in every kernel)."""
    rng = np.random.default_rng(2024)
    for case in range(12):
        ir, orr = [(44100, 48000), (48000, 44100), (8000, 12000), (48000, 30000)][case % 4]
        rows = int(rng.integers(1, 70))
        n = int(rng.integers(3000, 90000))
        x = rng.standard_normal((rows, n))
        a = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
        b = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
        c = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
        b.set_fusion(False)
        cuts = np.unique(np.concatenate([[0, n], rng.integers(0, n, size=int(rng.integers(0, 4)))]))
        ya, yb = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            xa = np.ascontiguousarray(x[:, lo:hi])
            ya.append(a.ProcessBatch(xa)[0].copy())
            yb.append(b.ProcessBatch(xa)[0].copy())
        ya.append(a.FlushBatch()[0].copy())
        yb.append(b.FlushBatch()[0].copy())
        yc = np.concatenate([c.ProcessBatch(x)[0].copy(), c.FlushBatch()[0].copy()], axis=1)
        ya, yb = np.concatenate(ya, axis=1), np.concatenate(yb, axis=1)
        np.testing.assert_array_equal(ya, yb, err_msg=f"case {case}: fused vs unfused")
        np.testing.assert_array_equal(ya, yc, err_msg=f"case {case}: chunked vs one-shot")


@pytest.mark.parametrize("ir,orr,rows,n", [
    (44100, 47999, 40, 50000),    # BASELINE config 5b's ratio: cubic coefficient interpolation live, 1.84 samples/output
    (48000, 44099, 24, 60000),    # irrational down-conversion, 2.18 samples/output
    (8000, 22051, 19, 30000),     # fewer than one intermediate sample per output (window slot stride 1)
    (44100, 16000, 17, 70000),    # rational, 5.5 samples/output: beyond the rational kernel's slot strides
])
def test_rows_kernel_batched_any_ratio_vs_thread_per_output_kernels_and_oracle(ir, orr, rows, n, vector_fir_only):
    """K3i (lanes = lock-step rows, interpolated coefficients evaluated once per batch): bit-identical to the
    one-thread-per-output kernels (same float64 summation order), <= 1e-12 against the oracle."""
    rng = np.random.default_rng(31)
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = [0, 5, n // 2 + 3, n]

    def run(tiled):
        G.set_tiled_polyphase(tiled)
        try:
            h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
            ys = [h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy() for lo, hi in zip(cuts[:-1], cuts[1:])]
            ys.append(h.FlushBatch()[0].copy())
            return np.concatenate(ys, axis=1), h.last_kernels()
        finally:
            G.set_tiled_polyphase(True)

    ya, ka = run(True)
    yb, kb = run(False)
    assert any(k.startswith("poly_rows_f64") for k in ka), ka
    assert not any(k.startswith(("poly_rows", "poly_rat", "fused_up2_rat")) for k in kb), kb
    np.testing.assert_array_equal(ya, yb)
    pick = sorted(set([0, rows // 2, rows - 1]))
    want, counts = O.batch_resample(x[pick], ir, orr, O.Q_HIGH, n_threads=4)
    assert np.all(counts == ya.shape[1])
    assert np.max(np.abs(ya[pick] - want[:, :ya.shape[1]])) <= 1e-12


@pytest.mark.parametrize("ir,orr,preset,rows,n,kernel", [
    (22050, 44100, G.QualityHigh, 8, 30000, "fir_f64_mma_up2"),        # x2 up-sampler, 166 taps per phase
    (96000, 48000, G.QualityVeryHigh, 8, 120000, "fir_f64_mma_s2"),    # /2 (path B maps VeryHigh to the 751-tap High filter)
    (48000, 16000, G.QualityHigh, 19, 90000, "fir_f64_mma_s3"),        # /3 with a ragged last group of streams
    (192000, 48000, G.QualityMedium, 9, 100000, "fir_f64_mma_s4"),     # /4
    (44100, 48000, G.QualityHigh, 70, 16000, "fir_f64_mma_up2"),       # x2 stage in front of the polyphase stage (>= 32 rows: unfused)
    # fewer than 8 rows: time segments of the rows are the MMA columns
    (96000, 48000, G.QualityVeryHigh, 2, 2100000, "fir_f64_mma_s2"),   # stereo, 2 rows x 4 segments
    (22050, 44100, G.QualityHigh, 1, 2050000, "fir_f64_mma_up2"),      # mono, 8 segments
    (48000, 16000, G.QualityHigh, 3, 2100003, "fir_f64_mma_s3"),       # 3 rows x 8 segments = 24 columns
    (192000, 48000, G.QualityMedium, 5, 1700003, "fir_f64_mma_s4"),    # 5 rows x 8 segments, ragged last segment
    (96000, 48000, G.QualityHigh, 7, 600001, "fir_f64_mma_s2"),        # 7 rows x 8 segments = 56 columns
])
def test_tensor_core_fir_matches_vector_kernels_and_oracle(ir, orr, preset, rows, n, kernel):
    """K1m/K2m (FP64 tensor cores, DMMA): same samples as the vector-FMA kernels to 1e-13 (taps grouped in fours
    instead of strictly sequential), <= 1e-12 against the oracle, identical counts; ragged chunking exercises the
    carried tails and the mix with the small-call vector kernels."""
    rng = np.random.default_rng(9)
    x = 0.6 * rng.standard_normal((rows, n))
    cuts = [0, 11, n // 3, n // 3 + 40000 if n // 3 + 40000 < n else n - 7, n]

    def run(tensor, chunks):
        G.set_tensor_fir(tensor)
        try:
            h = G.NewBatch(ir, orr, preset, rows, np.float64)
            ys = [h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy() for lo, hi in zip(chunks[:-1], chunks[1:])]
            ys.append(h.FlushBatch()[0].copy())
            return np.concatenate(ys, axis=1), h.last_kernels()
        finally:
            G.set_tensor_fir(True)

    ya, ka = run(True, [0, n])
    yb, kb = run(False, [0, n])
    yc, _ = run(True, cuts)
    assert kernel in ka, ka
    assert not any("mma" in k for k in kb), kb
    assert ya.shape == yb.shape == yc.shape
    assert np.max(np.abs(ya - yb)) <= 1e-13
    assert np.max(np.abs(yc - yb)) <= 1e-13
    pick = sorted(set([0, rows // 2, rows - 1]))
    want, counts = O.batch_resample(x[pick], ir, orr, O.preset_to_engine_quality(preset), n_threads=4)
    assert np.all(counts == ya.shape[1])
    assert np.max(np.abs(ya[pick] - want[:, :ya.shape[1]])) <= 1e-12


@pytest.mark.parametrize("ir,orr,rows,n", [
    (44100, 47999, 40, 50000),    # cubic coefficient interpolation live, 1.84 samples/output
    (48000, 44100, 70, 60000),    # rational 320/147, ragged last group of rows
    (44100, 48000, 64, 30000),    # rational 147/80
    (8000, 22051, 19, 30000),     # fewer than one intermediate sample per output
    (192000, 44100, 33, 120000),  # steep: 8.7 intermediate samples per output (one block per SM)
])
def test_tensor_core_polyphase_rows_kernel_vs_thread_per_output_kernels_and_oracle(ir, orr, rows, n):
    """K3m (polyphase stage as 8-output x K coefficient matrices times the rows' windows, DMMA): same samples as the
    one-thread-per-output kernels to 1e-13, <= 1e-12 against the oracle, ragged chunking + flush."""
    rng = np.random.default_rng(32)
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = [0, 9, n // 2 + 5, n]

    def run(tiled):
        G.set_tiled_polyphase(tiled)
        G.set_tensor_fir(tiled)
        try:
            h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
            ys = [h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy() for lo, hi in zip(cuts[:-1], cuts[1:])]
            ys.append(h.FlushBatch()[0].copy())
            return np.concatenate(ys, axis=1), h.last_kernels()
        finally:
            G.set_tiled_polyphase(True)
            G.set_tensor_fir(True)

    ya, ka = run(True)
    yb, kb = run(False)
    assert any(k.startswith("poly_rows_mma_f64") for k in ka), ka
    assert not any("mma" in k or k.startswith(("poly_rows", "poly_rat", "fused_up2_rat")) for k in kb), kb
    assert ya.shape == yb.shape
    assert np.max(np.abs(ya - yb)) <= 1e-13
    pick = sorted(set([0, rows // 2, rows - 1]))
    want, counts = O.batch_resample(x[pick], ir, orr, O.Q_HIGH, n_threads=4)
    assert np.all(counts == ya.shape[1])
    assert np.max(np.abs(ya[pick] - want[:, :ya.shape[1]])) <= 1e-12


@pytest.mark.parametrize("ir,orr,rows", [(44100, 48000, 12), (8000, 192000, 3), (44100, 47999, 9), (48000, 44100, 1)])
def test_time_sliced_calls_are_bit_identical_to_unsliced(ir, orr, rows):
    """A long multi-stage call runs as a sequence of shorter ones (inter-stage buffers L2-resident): identical samples
    and counts, BUFFER_TOO_SMALL still decided for the whole call before any state changes."""
    rng = np.random.default_rng(3)
    n = 150000 if orr < 100000 else 40000
    x = rng.standard_normal((rows, n))
    a = G.Resampler(_cfg(ir, orr), n_streams=rows)
    b = G.Resampler(_cfg(ir, orr), n_streams=rows)
    a.set_slice_budget(1 << 20)   # forces many slices
    b.set_slice_budget(0)
    with pytest.raises(G.ErrBufferTooSmall):
        a.ProcessBatch(x, np.empty((rows, 1000)))
    ya, na = a.ProcessBatch(x)
    yb, nb = b.ProcessBatch(x)
    assert na == nb
    np.testing.assert_array_equal(ya, yb)
    fa, fb = a.FlushBatch()[0], b.FlushBatch()[0]
    np.testing.assert_array_equal(fa, fb)


def test_flush_right_after_the_first_process_on_a_busy_gpu_and_mixed_streams():
    """the shared zero row and grown tail buffers are cleared on the launch stream, and calls on different streams
    of one handle are ordered by the engine's event. A large matmul keeps the GPU busy on torch's stream while fresh handles
    do Process -> Flush as their first device work; then device-batch calls alternate between a side stream, torch's stream
    and the handle's own stream and must equal the one-stream result bit for bit."""
    import torch
    a = torch.randn(8192, 8192, device="cuda")
    x = np.sin(2 * np.pi * 1000 * np.arange(6000) / 44100.0)
    want = O.resample_mono(x, 44100, 48000, O.PRESET_HIGH)
    for _ in range(12):
        (a @ a).sum()  # asynchronous: stays in flight while the handle works
        h = G.NewEngine(44100, 48000, G.QualityHigh)
        got = np.concatenate([h.Process(x), h.Flush()])
        assert len(got) == len(want) and np.max(np.abs(got - want)) <= 1e-12
    torch.cuda.synchronize()
    rows, n = 8, 20000
    xb = np.random.default_rng(5).standard_normal((rows, 4 * n))
    def run(streams):
        hb = G.NewBatch(44100, 47999, G.QualityHigh, rows, np.float64)
        dx = torch.from_numpy(xb).cuda()
        ost = (hb.EstimateOutput(n) + 3) & ~3
        dy = torch.zeros((5, rows, ost), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        counts = []
        for k in range(4):
            counts.append(hb.process_batch_dev(dx.data_ptr() + k * n * 8, 4 * n, n, dy[k].data_ptr(), ost, ost, streams[k % len(streams)], np.float64))
        counts.append(hb.flush_batch_dev(dy[4].data_ptr(), ost, ost, streams[0], np.float64))
        torch.cuda.synchronize()
        return np.concatenate([dy[k, :, :c].cpu().numpy() for k, c in enumerate(counts)], axis=1)
    side = torch.cuda.Stream()
    one = run([side.cuda_stream])
    mixed = run([side.cuda_stream, torch.cuda.current_stream().cuda_stream, 0])
    np.testing.assert_array_equal(one, mixed)


@pytest.mark.parametrize("ir,orr,rows,n", [(44100, 47999, 1, 200000), (48000, 44101, 1, 150000), (44100, 32001, 3, 90000),
                                            (32000, 47999, 7, 50000), (96000, 44101, 2, 120000), (22050, 47999, 1, 60000)])
def test_phase_sorted_fused_kernel_irrational_ratios_vs_thread_per_output_kernel_and_oracle(ir, orr, rows, n):
    """K4s (fused x2 -> polyphase with cubic coefficient interpolation, outputs sorted by phase inside a tile, a half-warp per
    phase): same operation order as the thread-per-output kernel => bit-identical, in one shot and in chunks; <= 1e-12 vs the
    oracle; counts exact (polyphase_stage.go:186-312)."""
    rng = np.random.default_rng(ir + orr + rows)
    x = 0.5 * rng.standard_normal((rows, n))
    cuts = [0, n // 3 + 17, n]

    def run(fast):
        G.set_tiled_polyphase(fast)
        try:
            h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
            ys = [h.ProcessBatch(np.ascontiguousarray(x[:, a:b]))[0].copy() for a, b in zip(cuts[:-1], cuts[1:])]
            ys.append(h.FlushBatch()[0].copy())
            return np.concatenate(ys, axis=1), h.last_kernels()
        finally:
            G.set_tiled_polyphase(True)

    fast, kf = run(True)
    slow, ks = run(False)
    assert "fused_up2_poly_sorted_f64" in kf and "fused_up2_poly_sorted_f64" not in ks, (kf, ks)
    np.testing.assert_array_equal(fast, slow)
    h1 = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
    one = np.concatenate([h1.ProcessBatch(x)[0], h1.FlushBatch()[0]], axis=1)
    np.testing.assert_array_equal(one, fast)  # chunked == one shot, bit for bit
    want, counts = O.batch_resample(x, ir, orr, O.Q_HIGH, n_threads=min(rows, 4))
    assert np.all(counts == fast.shape[1])
    assert np.max(np.abs(fast - want[:, :fast.shape[1]])) <= 1e-12


@pytest.mark.parametrize("ir,orr,rows,n", [(44100, 48000, 40, 60000), (48000, 44100, 33, 50000), (44100, 47999, 8, 80000),
                                            (32000, 44101, 12, 40000)])
def test_float32_engine_batches_run_wide_on_the_tensor_cores_within_1e6_of_the_float32_oracle(ir, orr, rows, n):
    """NewEngineFloat32 batches at non-integer ratios (convenience.go:329-366; polyphase_stage.go / dft_stage.go are generic
    over F): float32 samples and float32-rounded coefficients, float64 arithmetic on the FP64 tensor-core kernels
    (gar_handle::wide_f32). Counts exact, <= 1e-6 against the oracle's float32 path, chunked == one shot to 1e-6,
    and the handle still behaves as a float32 engine at the API."""
    rng = np.random.default_rng(rows)
    x = (0.5 * rng.standard_normal((rows, n))).astype(np.float32)
    h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float32)
    y = np.concatenate([h.ProcessBatch(x)[0], h.FlushBatch()[0]], axis=1)
    assert y.dtype == np.float32
    assert any("mma" in k for k in h.last_kernels()), h.last_kernels()
    want, counts = O.batch_resample(x[:3], ir, orr, O.Q_HIGH, n_threads=3)
    assert np.all(counts == y.shape[1])
    assert np.max(np.abs(y[:3].astype(np.float64) - want[:, :y.shape[1]].astype(np.float64))) <= 1e-6
    h.Reset()
    parts = [h.ProcessBatch(np.ascontiguousarray(c))[0].copy() for c in np.array_split(x, 3, axis=1)]
    y2 = np.concatenate(parts + [h.FlushBatch()[0]], axis=1)
    assert y2.shape == y.shape and np.max(np.abs(y2.astype(np.float64) - y)) <= 1e-6
    with pytest.raises(G.ErrNotSupported):  # still a float32 engine
        G.lib()  # noqa: B018
        h._process(0, x[0].astype(np.float64), np.float64)
    bank = h.bank(0)
    assert np.array_equal(bank, bank.astype(np.float32).astype(np.float64))  # coefficients are float32 values


@pytest.mark.parametrize("ir,orr,rows,n", [(44100, 47999, 1, 120000), (48000, 44101, 2, 90000), (22050, 47999, 5, 60000)])
def test_float32_engines_with_a_fractional_phase_step_run_wide_at_any_row_count(ir, orr, rows, n):
    """A float32 engine whose polyphase stage interpolates its coefficients (fractional phase step) computes in float64 at every row
    count: the phase-sorted kernel K4s serves 1-7 rows (a 10 s stream: 77 us against 199 us for the float32 thread-per-output
    kernel). Counts exact, <= 1e-6 against the oracle's float32 path, 4096-frame streaming chunks == one shot to 1e-6."""
    rng = np.random.default_rng(100 + rows)
    x = (0.5 * rng.standard_normal((rows, n))).astype(np.float32)
    h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float32)
    y = np.concatenate([h.ProcessBatch(x)[0], h.FlushBatch()[0]], axis=1)
    assert y.dtype == np.float32
    assert "fused_up2_poly_sorted_f64" in h.last_kernels(), h.last_kernels()
    want, counts = O.batch_resample(x, ir, orr, O.Q_HIGH, n_threads=2)
    assert np.all(counts == y.shape[1])
    assert np.max(np.abs(y.astype(np.float64) - want[:, :y.shape[1]].astype(np.float64))) <= 1e-6
    h.Reset()
    parts = [h.ProcessBatch(np.ascontiguousarray(x[:, i:i + 4096]))[0].copy() for i in range(0, n, 4096)]
    y2 = np.concatenate(parts + [h.FlushBatch()[0]], axis=1)
    assert y2.shape == y.shape and np.max(np.abs(y2.astype(np.float64) - y)) <= 1e-6


@pytest.mark.parametrize("ir,orr,rows,n,dt", [
    (8000, 192000, 1, 12000, np.float64),    # x24 (BASELINE config 5a's rates as ONE engine: path B)
    (8000, 56000, 2, 20000, np.float64),     # x7
    (48000, 4000, 3, 120000, np.float64),    # /12 decimator
    (44100, 6300, 1, 120000, np.float64),    # /7
    (8000, 96000, 2, 20000, np.float32),     # x12, float32 engine
    (48000, 4800, 2, 90000, np.float32),     # /10, float32 engine
])
def test_integer_factors_without_a_tiled_variant_shared_memory_kernel_vs_generic_and_oracle(ir, orr, rows, n, dt):
    """Path-B engines at integer ratios without a register-tiled variant (dft_stage.go with any factor): the shared-memory kernel
    (bank + window staged, one thread per output, sequential taps) against the generic one-thread-per-output kernel — bit-identical
    in float64, identical fold points in float32 — and the oracle."""
    rng = np.random.default_rng(77)
    x = (0.5 * rng.standard_normal((rows, n))).astype(dt)
    cuts = [0, n // 3 + 1, n]

    def run(fast):
        G.set_tensor_fir(fast)
        try:
            h = G.NewBatch(ir, orr, G.QualityHigh, rows, dt)
            ys = [h.ProcessBatch(np.ascontiguousarray(x[:, lo:hi]))[0].copy() for lo, hi in zip(cuts[:-1], cuts[1:])]
            ys.append(h.FlushBatch()[0].copy())
            return np.concatenate(ys, axis=1), h.last_kernels()
        finally:
            G.set_tensor_fir(True)

    ya, ka = run(True)
    yb, kb = run(False)
    assert any(k.endswith("_smem") for k in ka), ka
    assert not any(k.endswith("_smem") for k in kb) and any(k.endswith("_generic") for k in kb), kb
    assert np.array_equal(ya, yb), float(np.max(np.abs(ya.astype(np.float64) - yb)))
    want, counts = O.batch_resample(x[:1], ir, orr, O.Q_HIGH, n_threads=1)
    assert counts[0] == ya.shape[1]
    assert np.max(np.abs(ya[0].astype(np.float64) - want[0, :ya.shape[1]].astype(np.float64))) <= (1e-12 if dt == np.float64 else 1e-6)


@pytest.mark.parametrize("ir,orr,rows,n,dt,kernel", [
    (8000, 48000, 2, 30000, np.float64, "fir_f64_up6_r2"),
    (8000, 40000, 1, 30000, np.float64, "fir_f64_up5_r2"),
    (8000, 64000, 3, 20000, np.float64, "fir_f64_up8_r2"),
    (48000, 8000, 3, 90000, np.float64, "fir_f64_s6_r3"),
    (44100, 8820, 1, 120000, np.float64, "fir_f64_s5_r2"),
    (64000, 8000, 2, 120000, np.float64, "fir_f64_s8_r2"),
    (8000, 40000, 2, 20000, np.float32, "fir_f32_up5_r4"),
    (8000, 48000, 2, 20000, np.float32, "fir_f32_up6_r4"),
    (48000, 8000, 2, 60000, np.float32, "fir_f32_s6_r6"),
    (44100, 8820, 2, 60000, np.float32, "fir_f32_s5_r4"),
])
def test_register_tiled_fir_variants_for_factors_5_6_8(ir, orr, rows, n, dt, kernel):
    """x5 x6 x8 and /5 /6 /8 (8k <-> 48k, 44.1k -> 8.82k ...): the register-tiled kernel's variants for these factors against the
    oracle; float64 chunked == one shot bit for bit (same tap order whatever the chunking), counts exact."""
    rng = np.random.default_rng(78)
    x = (0.5 * rng.standard_normal((rows, n))).astype(dt)
    h = G.NewBatch(ir, orr, G.QualityHigh, rows, dt)
    y = np.concatenate([h.ProcessBatch(x)[0], h.FlushBatch()[0]], axis=1)
    assert kernel in h.last_kernels(), h.last_kernels()
    want, counts = O.batch_resample(x[:1], ir, orr, O.Q_HIGH, n_threads=1)
    assert counts[0] == y.shape[1]
    assert np.max(np.abs(y[0].astype(np.float64) - want[0, :y.shape[1]].astype(np.float64))) <= (1e-12 if dt == np.float64 else 1e-6)
    h.Reset()
    parts = [h.ProcessBatch(np.ascontiguousarray(c))[0].copy() for c in np.array_split(x, 3, axis=1)]
    y2 = np.concatenate(parts + [h.FlushBatch()[0]], axis=1)
    assert y2.shape == y.shape
    if dt == np.float64:
        assert np.array_equal(y2, y)
    else:
        assert np.max(np.abs(y2.astype(np.float64) - y)) <= 1e-6
