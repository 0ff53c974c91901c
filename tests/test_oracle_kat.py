"""Pins the CPU oracle against every exact known answer the reference's own
tests hold for the hot path. Paths are relative to /root/reference.
"""
import math

import numpy as np
import pytest


# --- internal/simdops/ops_test.go:25-70 ------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_simdops_known_answers(oracle, dt):
    L = oracle.lib()
    sfx = "f32" if dt == np.float32 else "f64"
    p = oracle._ptr
    a = np.array([1, 2, 3], dtype=dt)
    b = np.array([4, 5, 6], dtype=dt)
    assert abs(getattr(L, f"orc_dot_{sfx}")(p(a), p(b), 3) - 32.0) < 1e-5
    sig = np.array([1, 2, 3, 4], dtype=dt)
    ker = np.array([1, 0.5], dtype=dt)
    dst = np.zeros(3, dtype=dt)
    getattr(L, f"orc_convolve_valid_{sfx}")(p(dst), p(sig), 4, p(ker), 2)
    np.testing.assert_allclose(dst, [2, 3.5, 5], atol=1e-5)
    ker2 = np.array([0, 2], dtype=dt)  # ConvolveValidMulti == one ConvolveValid per kernel
    getattr(L, f"orc_convolve_valid_{sfx}")(p(dst), p(sig), 4, p(ker2), 2)
    np.testing.assert_allclose(dst, [4, 6, 8], atol=1e-5)
    il = np.zeros(6, dtype=dt)
    getattr(L, f"orc_interleave2_{sfx}")(p(il), p(a), p(b), 3)
    np.testing.assert_allclose(il, [1, 4, 2, 5, 3, 6], atol=1e-5)
    assert abs(getattr(L, f"orc_sum_{sfx}")(p(a), 3) - 6.0) < 1e-5
    sc = np.zeros(3, dtype=dt)
    getattr(L, f"orc_scale_{sfx}")(p(sc), p(a), 3, 2.0)
    np.testing.assert_allclose(sc, [2, 4, 6], atol=1e-5)
    h = np.array([0.5, 0.5], dtype=dt)
    ca, cb, cc, cd = (np.array(v, dtype=dt) for v in ([1, 2], [3, 4], [5, 6], [7, 8]))
    got = getattr(L, f"orc_cubic_interp_dot_{sfx}")(p(h), p(ca), p(cb), p(cc), p(cd), 0.5, 2)
    assert abs(got - 5.5625) < 1e-5


def test_dot_long_vectors_match_numpy(oracle):
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 31, 32, 33, 877, 1223):
        a = rng.standard_normal(n)
        b = rng.standard_normal(n)
        got = oracle.lib().orc_dot_f64(oracle._ptr(a), oracle._ptr(b), n)
        assert abs(got - float(np.dot(a, b))) < 1e-12 * max(1, n)
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        got32 = oracle.lib().orc_dot_f32(oracle._ptr(a32), oracle._ptr(b32), n)
        assert abs(got32 - float(np.dot(a32.astype(np.float64), b32.astype(np.float64)))) < 2e-5


# --- internal/mathutil/bessel_test.go:14-42, :89-110 -------------------------
@pytest.mark.parametrize("x,exp,tol", [
    (0.0, 1.0, 1e-15), (0.5, 1.063483344, 1e-7), (1.0, 1.266065848, 1e-7), (2.0, 2.279585307, 1e-7),
    (3.0, 4.880792565, 1e-7), (3.75, 9.118945994, 1e-7), (4.0, 11.30192217, 1e-7), (5.0, 27.23987183, 1e-7),
    (10.0, 2815.716628, 1e-6), (20.0, 4.355826e7, 1e-1), (-0.5, 1.063483344, 1e-7), (-1.0, 1.266065848, 1e-7)])
def test_bessel_i0_table(oracle, x, exp, tol):
    got = oracle.lib().orc_bessel_i0(x)
    assert abs(got - exp) / abs(exp) <= tol


@pytest.mark.parametrize("att,lo,hi", [(20, 0, 0.1), (50, 4.5, 4.6), (60, 5.6, 5.7), (80, 7.8, 7.9),
                                       (100, 10.0, 10.1), (120, 12.2, 12.3)])
def test_kaiser_beta_ranges(oracle, att, lo, hi):
    assert lo <= oracle.lib().orc_kaiser_beta(float(att)) <= hi


# --- internal/engine/critical_functions_test.go:18-53 ------------------------
@pytest.mark.parametrize("r,exp", [
    (1.0, True), (2.0, True), (3.0, True), (4.0, True), (2.0000000001, True), (0.5, False), (0.333333, False),
    (1.5, False), (1.088435374, False), (0.91875, False), (2.1768707, False), (0.0, False), (0.999999, False),
    (0.9999999999, True), (1.0000000001, True)])
def test_is_integer_ratio_table(oracle, r, exp):
    assert bool(oracle.lib().orc_is_integer_ratio(r)) == exp


# --- critical_functions_test.go:59-98 ----------------------------------------
@pytest.mark.parametrize("ratio", [1.088435374, 0.91875, 1.0, 0.5, 0.25, 2.0])
def test_find_rational_approx(oracle, ratio):
    import ctypes as C
    L, s = C.c_int(0), C.c_int(0)
    oracle.lib().orc_find_rational_approx(ratio, C.byref(L), C.byref(s))
    assert 64 <= L.value <= 256 and s.value > 0
    assert abs(s.value / L.value - 1 / ratio) / (1 / ratio) < 0.01


# --- critical_functions_test.go:104-165 --------------------------------------
@pytest.mark.parametrize("drop,att", [(-0.01, 180), (-0.01, 140), (-0.01, 100), (-0.1, 180), (-1, 180), (-3, 180),
                                      (-6, 180), (-0.01, 1), (-0.01, 300), (0, 180), (-20, 180)])
def test_lsx_inv_f_resp_valid(oracle, drop, att):
    v = oracle.lib().orc_lsx_inv_f_resp(float(drop), float(att))
    assert math.isfinite(v) and 0.0 <= v <= 1.0


def test_lsx_inv_f_resp_monotonic(oracle):
    vals = [oracle.lib().orc_lsx_inv_f_resp(d, 180.0) for d in (-0.001, -0.01, -0.1, -1.0, -3.0, -6.0)]
    assert all(b > a for a, b in zip(vals, vals[1:]))


# --- critical_functions_test.go:183-310 (Fn normalisation) -------------------
@pytest.mark.parametrize("L,ratio,tio,pre,fn,up", [
    (147, 48000 / 44100, 44100 / 48000, True, 1.0, True),
    (147, 96000 / 44100, 44100 / 96000, True, 1.0, True),
    (160, 44100 / 48000, 48000 / 44100, False, 1.0, False),
    (1, 0.5, 2.0, False, 1.0, False),
    (2, 32000 / 48000, 1.5, False, 1.0, False),
    (160, 44100 / 48000, 48000 / 44100, True, 2.0 * 1.088, False),
    (1, 0.5, 2.0, True, 4.0, False)])
def test_polyphase_params_fn(oracle, L, ratio, tio, pre, fn, up):
    dv = np.zeros(10)
    iv = np.zeros(3, dtype=np.int32)
    oracle.lib().orc_polyphase_params(L, ratio, tio, int(pre), 126.0, 0.912, oracle._ptr(dv), oracle._ptr(iv))
    assert bool(iv[0]) == up
    assert abs(dv[1] - fn) <= fn * 0.01
    if (not up) and pre:
        assert abs(dv[5] - (3.0 + abs(ratio - 1.0))) < 0.01
    assert abs(dv[4] / dv[1] - dv[6]) < 1e-4 and abs(dv[5] / dv[1] - dv[7]) < 1e-4
    assert 0.0 < dv[9] < 1.0


# --- critical_functions_test.go:628-646 --------------------------------------
def test_quality_attenuation(oracle):
    Q = oracle
    for q, bits in [(Q.Q_LOW, 16), (Q.Q_MEDIUM, 16), (Q.Q_HIGH, 20), (999, 20)]:
        assert abs(Q.lib().orc_quality_attenuation(q) - (bits + 1) * 6.0206) < 1e-10


# --- internal/pipeline/pipeline_test.go:13-175 -------------------------------
@pytest.mark.parametrize("ratio,prec,types", [
    (1.5, 8, ["cubic"]),
    (0.125, 16, ["hb", "hb", "fft"]),
    (0.15, 16, ["hb", "hb", "poly"]),
    (8.0, 16, ["hb", "hb", "fft"]),
    (10.0, 16, ["hb", "hb", "hb", "poly"]),
    (44100.0 / 48000.0, 16, ["fft"]),
    (48000.0 / 44100.0, 16, ["fft"]),
    (1.5, 28, ["fft"]),
    (1.5, 32, ["fft"]),
    (1.5, 24, ["poly"]),
    (1.0, 16, [])])
def test_planner_stage_sequences(oracle, ratio, prec, types):
    names = {oracle.ST_CUBIC: "cubic", oracle.ST_HALFBAND: "hb", oracle.ST_POLYPHASE: "poly", oracle.ST_FFT: "fft"}
    assert [names[t] for t, _ in oracle.build_plan(ratio, prec)] == types


@pytest.mark.parametrize("ratio", [0.0, -1.5])
def test_planner_rejects_bad_ratio(oracle, ratio):
    with pytest.raises(ValueError):
        oracle.build_plan(ratio, 16)


# --- internal/engine/extra_engine_test.go:85-121 (exact lengths) --------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_exact_lengths_upsample_2x(oracle, dt):
    e = oracle.Engine(22050, 44100, oracle.Q_HIGH, dt)
    d = e.describe()
    assert d["pre_factor"] == 2 and d["pre_tpp"] == 166 and d["pre_proto_taps"] == 331
    assert len(e.process(np.full(5000, 0.05, dtype=dt))) == 9670


def test_exact_lengths_decimate_2x(oracle):
    e = oracle.Engine(44100, 22050, oracle.Q_HIGH)
    assert e.describe()["dec_taps"] == 751
    assert len(e.process(np.full(5000, 0.05))) == 2125


# --- README.md:466-471 (High: 166x2 + 64x80) ----------------------------------
def test_readme_filter_complexity_high(oracle):
    d = oracle.Engine(44100, 48000, oracle.Q_HIGH).describe()
    assert (d["pre_tpp"], d["pre_factor"], d["poly_tpp"], d["poly_L"]) == (166, 2, 64, 80)
    d = oracle.Engine(44100, 48000, oracle.Q_LOW).describe()
    assert (d["pre_tpp"], d["poly_tpp"], d["poly_L"]) == (132, 32, 80)
    # README's VeryHigh row (166x2 + 100x80) matches neither preset path in the code (path B maps
    # VeryHigh->engine.QualityHigh; att>=160 dB lifts the cap to 8191/80=102): documentation drift, not pinned.
    d = oracle.Engine(44100, 48000, oracle.Q_24BIT).describe()
    assert (d["pre_tpp"], d["poly_tpp"], d["poly_L"]) == (200, 100, 80)


# --- the two preset maps (stages.go:92-108, convenience.go:189-200)
def test_preset_maps(oracle):
    O = oracle
    assert [O.lib().orc_preset_to_engine_quality(p) for p in range(5)] == [O.Q_LOW, O.Q_LOW, O.Q_MEDIUM, O.Q_HIGH, O.Q_HIGH]
    assert [O.lib().orc_precision_to_engine_quality(O.lib().orc_preset_precision(p)) for p in range(5)] == \
        [O.Q_QUICK, O.Q_LOW, O.Q_LOW, O.Q_24BIT, O.Q_32BIT]
    for prec, q in [(8, O.Q_QUICK), (16, O.Q_LOW), (17, O.Q_HIGH), (20, O.Q_HIGH), (21, O.Q_24BIT), (24, O.Q_24BIT),
                    (25, O.Q_VERYHIGH), (28, O.Q_VERYHIGH), (29, O.Q_32BIT), (33, O.Q_32BIT)]:
        assert O.lib().orc_precision_to_engine_quality(prec) == q


# --- filter/soxr_filter_test.go style properties ------------------------------
def test_lowpass_symmetry_and_dc_gain(oracle):
    out = np.zeros(8191)
    n = oracle.lib().orc_design_lowpass_auto(0.4778321 / 2, 0.05 / 2, 126.4326, 1.0, oracle._ptr(out), 8191)
    h = out[:n]
    assert n == 331
    np.testing.assert_allclose(h, h[::-1], atol=1e-15)
    assert abs(h.sum() - 1.0) < 1e-12


def test_half_band_shortcut_never_fires(oracle):
    # with cutoff 0.4778321/2 phase 0 is never a single-tap passthrough
    for q in (oracle.Q_LOW, oracle.Q_MEDIUM, oracle.Q_HIGH, oracle.Q_VERYHIGH, oracle.Q_24BIT, oracle.Q_32BIT):
        assert oracle.Engine(48000, 96000, q).describe()["pre_half_band"] == 0
