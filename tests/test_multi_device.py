"""Multi-device handle (gar_create_multi): ONE handle shards its rows over several devices — the GPU analogue of the channel
fan-out inside one call (constant.go:223-241). CPU tests run it with geometry-only shards (device -1: partition, routing of
per-row calls, lock-step checks, error behaviour); GPU tests compare a multi handle with the single-device handle sample by
sample (1 device always, 2 devices in one process when the box shows >= 2)."""
import numpy as np
import pytest

from helpers import G, O, sig_c3, sig_c4
from gar_b200.shard import partition


def _cfg(ir, orr, ch, preset=G.QualityHigh):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset))


def test_shard_layout_matches_the_rank_partition():
    for streams, k in ((37, 2), (4096, 8), (5, 8), (9, 3)):
        h = G.NewBatch(48000, 16000, G.QualityMedium, streams, np.float32, devices=[-1] * k)
        sh = h.shards()
        assert len(sh) == min(k, streams)
        assert [(r0, n) for _, r0, n, _ in sh] == [partition(streams, len(sh), r) for r in range(len(sh))]
        assert sum(n for _, _, n, _ in sh) == h.rows == streams
    # a single stream with channels shards by channel; stereo streams stay whole
    assert [(r0, n) for _, r0, n, _ in G.Resampler(_cfg(96000, 48000, 8), devices=[-1, -1, -1]).shards()] == [(0, 3), (3, 3), (6, 2)]
    assert [(r0, n) for _, r0, n, _ in G.Resampler(_cfg(48000, 44100, 2), n_streams=5, devices=[-1, -1]).shards()] == [(0, 6), (6, 4)]


def test_per_row_state_is_routed_to_the_owning_shard():
    multi = G.Resampler(_cfg(44100, 48000, 6), devices=[-1, -1])
    single = G.Resampler(G.Config(InputRate=44100, OutputRate=48000, Channels=6, Quality=G.QualitySpec(Preset=G.QualityHigh), Device=-1))
    for ch, n in ((0, 1000), (4, 777), (5, 12345), (4, 3), (2, 50000)):
        assert multi.advance_geometry(n, stream=ch) == single.advance_geometry(n, stream=ch)
    for ch in range(6):
        assert multi.describe(ch) == single.describe(ch)
        assert multi.GetStatistics(ch) == single.GetStatistics(ch)
        assert multi.advance_geometry(0, flush=True, stream=ch) == single.advance_geometry(0, flush=True, stream=ch)
    assert multi.GetLatency() == single.GetLatency() and multi.EstimateOutput(4096) == single.EstimateOutput(4096)
    multi.Reset()
    assert multi.describe(5) == G.Resampler(G.Config(InputRate=44100, OutputRate=48000, Channels=6,
                                                     Quality=G.QualitySpec(Preset=G.QualityHigh), Device=-1)).describe(5)


def test_multi_handle_errors():
    with pytest.raises(G.ErrInvalidConfig):
        G.NewBatch(48000, 16000, G.QualityMedium, 4, np.float32, devices=[])
    with pytest.raises(G.ErrInvalidConfig):  # Config.Validate runs once, on the whole config
        G.Resampler(_cfg(48000, 16000, 300), devices=[-1, -1])
    h = G.NewBatch(48000, 16000, G.QualityMedium, 4, np.float32, devices=[-1, -1])
    with pytest.raises(G.CudaError):        # geometry-only shards refuse samples: no CPU fallback
        h.ProcessBatch(np.zeros((4, 4800), np.float32))
    with pytest.raises(G.ErrNotSupported):  # device pointers belong to one device
        h.process_batch_dev(0, 4800, 4800, 0, 2000, 2000)
    h.advance_geometry(4800, stream=3)      # rows out of lock step are rejected before anything moves
    with pytest.raises(G.ErrNotSupported):
        h.ProcessBatch(np.zeros((4, 4800), np.float32))


def _devices():
    n = G.device_count()
    return [[0]] + ([[0, 1]] if n >= 2 else [])


@pytest.mark.gpu
@pytest.mark.parametrize("devices", _devices() if G.device_count() else [[0]])
def test_multi_handle_batch_equals_single_device_handle(devices):
    rows, n = 37, 48000
    x = sig_c4(rows, n)
    ref = G.NewBatch(48000, 16000, G.QualityMedium, rows, np.float32)
    want = np.concatenate([ref.ProcessBatch(x)[0], ref.FlushBatch()[0]], axis=1)
    h = G.NewBatch(48000, 16000, G.QualityMedium, rows, np.float32, devices=devices)
    assert [d for d, _, _, _ in h.shards()] == devices
    xh, px = h.host_alloc_rows(n, np.float32)  # pinned, first-touched by the shards' bound workers
    xh[:] = x
    for chunks in (1, 3):  # one shot, and streaming chunks (carry state stays with the shard that owns the row)
        h.Reset()
        parts = [h.ProcessBatch(np.ascontiguousarray(c))[0].copy() for c in np.array_split(xh, chunks, axis=1)]
        got = np.concatenate(parts + [h.FlushBatch()[0]], axis=1)
        if chunks == 1 and len(devices) == 1:
            np.testing.assert_array_equal(got, want)
        else:
            assert got.shape == want.shape and np.max(np.abs(got - want)) <= 2.5e-7
    G.host_free(px)
    owant, _ = O.batch_resample(x[:2], 48000, 16000, O.Q_MEDIUM, n_threads=2)
    assert np.max(np.abs(got[:2].astype(np.float64) - owant[:, :got.shape[1]])) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("devices", _devices() if G.device_count() else [[0]])
def test_multi_handle_process_multi_and_per_channel_calls(devices):
    xs = [c[:96000] for c in sig_c3(96000, 8)]
    ref = G.New(_cfg(96000, 48000, 8, G.QualityVeryHigh))
    want = [np.concatenate([a, b]) for a, b in zip(ref.ProcessMulti(xs), ref.FlushMulti())]
    h = G.Resampler(_cfg(96000, 48000, 8, G.QualityVeryHigh), devices=devices)
    got = [np.concatenate([a, b]) for a, b in zip(h.ProcessMulti(xs), h.FlushMulti())]
    # 8 rows on one device take the tensor-core decimator, 4 + 4 rows the vector kernel: same taps, different grouping
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.max(np.abs(g - w)) <= 1e-12
    h.Reset()
    a = h._process(7, xs[7], np.float64)  # a channel owned by the last shard
    assert np.max(np.abs(np.concatenate([a, h._flush(7, np.float64)]) - want[7])) <= 1e-12
    assert any("mma" in k or "fir" in k for k in h.last_kernels())
