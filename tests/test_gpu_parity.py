"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the five BASELINE configs.

Tolerances are the north star's: sample counts / flush lengths bit-exact; float64 path <= 1e-12 max abs
error; float32 path <= 1e-6 max abs error.
"""
import numpy as np
import pytest

from helpers import G, O, sig_c1, sig_c2, sig_c3, sig_c4, sig_c5a

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-6


def _cfg(ir, orr, ch=1, preset=G.QualityHigh, parallel=False):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset),
                    EnableParallel=parallel)


def _maxerr(a, b):
    assert len(a) == len(b), (len(a), len(b))
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if len(a) else 0.0


def test_c1_resample_mono_44k1_to_48k_high_f64():
    x = sig_c1()
    got = G.ResampleMono(x, 44100, 48000, G.QualityHigh)
    want = O.resample_mono(x, 44100, 48000, O.PRESET_HIGH)
    assert len(got) == 479787 + 215
    assert _maxerr(got, want) <= TOL64


def test_c2_stereo_streaming_48k_to_44k1_f32_chunks_with_flush():
    left, right = sig_c2()
    for x in (left, right):  # ProcessFloat32Into is channel-0 only: two mono instances
        r = G.New(_cfg(48000, 44100))
        p = O.Pipeline(48000, 44100, 1, O.PRESET_HIGH)
        out = np.empty(r.EstimateOutput(4096), dtype=np.float32)
        ref = np.empty(p.estimate_output(4096), dtype=np.float32)
        assert len(out) == 3827
        total, worst = 0, 0.0
        for i in range(0, len(x), 4096):
            n = r.ProcessFloat32Into(x[i:i + 4096], out)
            m = p.process_f32_into(x[i:i + 4096], ref)
            assert n == m  # per-chunk counts are bit-exact
            worst = max(worst, _maxerr(out[:n], ref[:m]))
            total += n
        assert total == 440793
        fl, fr = r.Flush(), p.flush()
        assert len(fl) == len(fr) == 209
        assert _maxerr(fl, fr) <= TOL64  # Flush returns float64 (resample.go:32)
        assert worst <= TOL32


def test_c3_surround_8ch_96k_to_48k_veryhigh_f64_process_multi():
    xs = sig_c3()
    r = G.New(_cfg(96000, 48000, 8, G.QualityVeryHigh, parallel=True))
    p = O.Pipeline(96000, 48000, 8, O.PRESET_VERYHIGH)
    got = r.ProcessMulti(xs)
    gf = r.FlushMulti()
    for c in range(8):
        want = p.process(xs[c], c)
        wf = p.flush(c)
        assert len(got[c]) == 479389 and len(gf[c]) == 612
        assert _maxerr(got[c], want) <= TOL64
        assert _maxerr(gf[c], wf) <= TOL64


@pytest.mark.parametrize("preset,taps", [(G.QualityMedium, 877), (G.QualityHigh, 1125), (G.QualityLow, 245)])
def test_c4_batched_mono_streams_48k_to_16k_f32(preset, taps):
    ns, n = 96, 48000
    x = sig_c4(ns, n)
    b = G.NewBatch(48000, 16000, preset, ns, np.float32)
    assert b.describe()[0]["taps"] == taps
    y, ny = b.ProcessBatch(x)
    f, nf = b.FlushBatch()
    want, counts = O.batch_resample(x, 48000, 16000, O.preset_to_engine_quality(preset), n_threads=8)
    assert np.all(counts == ny + nf)
    got = np.concatenate([y, f], axis=1)
    # referee: float64 arithmetic on the same float32 operands (bank as stored on the device, reversed taps)
    bank = b.bank(0, 0)
    e_gpu = e_orc = 0.0
    for r in range(0, ns, 8):
        v = np.concatenate([x[r].astype(np.float64), np.zeros(taps)])
        exact = np.correlate(v, bank, mode="valid")[::3][:ny + nf]
        e_gpu = max(e_gpu, float(np.max(np.abs(got[r].astype(np.float64) - exact))))
        e_orc = max(e_orc, float(np.max(np.abs(want[r, :ny + nf].astype(np.float64) - exact))))
    err = float(np.max(np.abs(got.astype(np.float64) - want[:, :ny + nf].astype(np.float64))))
    print(f"taps={taps}: |gpu-exact|={e_gpu:.3e} |oracle-exact|={e_orc:.3e} |gpu-oracle|={err:.3e}")
    assert e_gpu <= 2.5e-7, (e_gpu, e_orc)  # f32 output rounding (6e-8 at |y|~1) + short f32 partial sums
    assert e_orc <= TOL32, e_orc
    assert err <= TOL32, (err, e_gpu, e_orc)


def test_c4_full_length_rows_chunked_vs_one_shot_and_oracle():
    """Full 480 000-sample rows: counts (159708 + 293), a row subset against the oracle, and chunked vs
    one-shot. The float32 kernels pair taps into packed FMAs by memory alignment, so different chunkings may
    round differently (<= 2.5e-7, both within 1.5e-7 of exact); identical call sequences are bit-identical."""
    ns, n = 32, 480000
    x = sig_c4(ns, n)
    b = G.NewBatch(48000, 16000, G.QualityMedium, ns, np.float32)
    y, ny = b.ProcessBatch(x)
    f, nf = b.FlushBatch()
    assert (ny, nf) == (159708, 293)
    one = np.concatenate([y, f], axis=1).copy()
    want, _ = O.batch_resample(x[:4], 48000, 16000, O.Q_MEDIUM, n_threads=4)
    assert float(np.max(np.abs(one[:4].astype(np.float64) - want[:, :ny + nf]))) <= TOL32

    def chunked(step):
        b.Reset()
        parts = []
        for i in range(0, n, step):
            yy, k = b.ProcessBatch(np.ascontiguousarray(x[:, i:i + step]))
            parts.append(yy[:, :k].copy())
        ff, k = b.FlushBatch()
        parts.append(ff[:, :k].copy())
        return np.concatenate(parts, axis=1)

    c1 = chunked(65536 + 17)   # 65553 = 3 * 21851: the decimation phase stays 0
    c2 = chunked(50000 + 1)    # phase cycles through 0, 1, 2: tiles start at every alignment
    assert c1.shape == c2.shape == one.shape
    assert float(np.max(np.abs(c1.astype(np.float64) - one))) <= 2.5e-7
    assert float(np.max(np.abs(c2.astype(np.float64) - one))) <= 2.5e-7
    np.testing.assert_array_equal(chunked(50000 + 1), c2)  # deterministic for a given call sequence


def test_f64_decimator_chunked_equals_one_shot_bit_for_bit():
    """The float64 kernels sum taps strictly in order, independent of tile position and alignment."""
    x = np.stack(sig_c3(200000, 2))
    b = G.NewBatch(96000, 48000, G.QualityHigh, 2, np.float64)
    y, ny = b.ProcessBatch(x)
    f, nf = b.FlushBatch()
    one = np.concatenate([y, f], axis=1).copy()
    b.Reset()
    parts = []
    for i in range(0, x.shape[1], 33333):
        yy, k = b.ProcessBatch(np.ascontiguousarray(x[:, i:i + 33333]))
        parts.append(yy[:, :k].copy())
    ff, k = b.FlushBatch()
    parts.append(ff[:, :k].copy())
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), one)
    e = O.Engine(96000, 48000, O.Q_HIGH)
    want = np.concatenate([e.process(x[0]), e.flush()])
    assert _maxerr(one[0], want) <= TOL64


def test_c5a_extreme_upsampling_8k_to_192k_multistage_f64():
    x = sig_c5a()
    r = G.New(_cfg(8000, 192000))
    p = O.Pipeline(8000, 192000, 1, O.PRESET_HIGH)
    got, want = r.Process(x), p.process(x)
    gf, wf = r.Flush(), p.flush()
    assert (len(got), len(gf)) == (1910673, 9375)
    assert _maxerr(got, want) <= TOL64 and _maxerr(gf, wf) <= TOL64


def test_c5b_irrational_ratio_cubic_coefficient_interpolation_f64():
    x = sig_c1()
    r = G.NewEngine(44100, 47999, G.QualityHigh)
    d = r.describe()[1]
    assert d["factor"] == 197 and d["taps"] == 41 and d["step"] == 23723707 and d["step"] & 0xFFFF == 65211
    got = np.concatenate([r.Process(x), r.Flush()])
    want = O.resample_mono(x, 44100, 47999, O.PRESET_HIGH)
    assert len(got) == 479789 + 203
    assert _maxerr(got, want) <= TOL64


def test_c4_alt_path_a_float32_io_float64_inside():
    """BASELINE C4' : New(Config{48k->16k, High}) + ProcessFloat32Into: half-band /2 -> x2 -> polyphase."""
    x = sig_c4(1, 480000)[0]
    r = G.New(_cfg(48000, 16000))
    p = O.Pipeline(48000, 16000, 1, O.PRESET_HIGH)
    assert [d["kind"] for d in r.describe()] == [G.STAGE_DECIM, G.STAGE_UP, G.STAGE_POLY]
    out = np.empty(r.EstimateOutput(len(x)), dtype=np.float32)
    ref = np.empty(p.estimate_output(len(x)), dtype=np.float32)
    n, m = r.ProcessFloat32Into(x, out), p.process_f32_into(x, ref)
    assert n == m == 159531
    assert _maxerr(out[:n], ref[:m]) <= TOL32
    gf, wf = r.Flush(), p.flush()
    assert len(gf) == 471 and _maxerr(gf, wf) <= TOL64


@pytest.mark.parametrize("ir,orr,counts", [(44100, 48000, (479787, 215)), (44100, 47999, (479789, 203))])
def test_c1_c5b_full_length_batched_on_the_tensor_cores_equals_single_streams_and_oracle(ir, orr, counts):
    """BASELINE configs 1 and 5b at full length as ONE lock-step batch call on device buffers (128 rows x 441 000 samples:
    K1m + K3p on the FP64 tensor cores): every row equals the single-stream engine's result for its signal to 1e-13 (rows are
    independent; the single stream takes the fused vector kernels), counts are the reference's, and the sine row is within
    1e-12 of the oracle."""
    import torch
    rows, n = 128, 441000
    rng = np.random.default_rng(77)
    base = [sig_c1(), 0.7 * rng.standard_normal(n), np.where(np.arange(n) % 1000 == 0, 1.0, 0.0), rng.uniform(-1, 1, n)]
    x = np.empty((rows, n))
    for r in range(rows):
        x[r] = base[r % 4]
    h = G.NewBatch(ir, orr, G.QualityHigh, rows, np.float64)
    dx = torch.from_numpy(x).cuda()
    ostride = (h.EstimateOutput(n) + 3) & ~3
    dy = torch.zeros((rows, ostride), dtype=torch.float64, device="cuda")
    df = torch.zeros((rows, 1024), dtype=torch.float64, device="cuda")
    ts = torch.cuda.current_stream().cuda_stream  # the stream that produced dx/dy: the calls are ordered after it
    c1 = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), ostride, ostride, ts, np.float64)
    used = set(h.last_kernels())
    c2 = h.flush_batch_dev(df.data_ptr(), 1024, 1024, ts, np.float64)
    torch.cuda.synchronize()
    assert (c1, c2) == counts
    assert "fir_f64_mma_up2" in used and any(k.startswith("poly_rows_mma_f64") and k.endswith("pipe") for k in used), used
    got = np.concatenate([dy[:, :c1].cpu().numpy(), df[:, :c2].cpu().numpy()], axis=1)
    for s in range(4):
        e = G.NewEngine(ir, orr, G.QualityHigh)
        single = np.concatenate([e.Process(base[s]), e.Flush()])
        assert single.shape[0] == got.shape[1]
        for r in range(s, rows, 4):
            assert np.max(np.abs(got[r] - single)) <= 1e-13, (s, r)
    want = O.resample_mono(base[0], ir, orr, O.PRESET_HIGH)
    assert _maxerr(got[0], want) <= TOL64
