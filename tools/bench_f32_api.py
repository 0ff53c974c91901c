import sys, numpy as np, torch
sys.path.insert(0, 'go-audio-resampler_b200/python')
import gar_b200 as G
dev = torch.device('cuda', 0); ts = torch.cuda.Stream(device=dev)
for (ir, orr, preset, rows, n) in [(96000, 48000, G.QualityVeryHigh, 8, 960000), (48000, 16000, G.QualityHigh, 256, 480000), (48000, 44100, G.QualityHigh, 256, 480000)]:
    x = np.random.default_rng(0).standard_normal((rows, n)).astype(np.float32)
    dx = torch.from_numpy(x).to(dev)
    for fold in (True, False):
        h = G.Resampler(G.Config(InputRate=ir, OutputRate=orr, Channels=1, Quality=G.QualitySpec(Preset=preset)), n_streams=rows)
        h.set_fusion(fold)
        ostride = (h.EstimateOutput(n) + 8192 + 3) & ~3
        dy = torch.zeros((rows, ostride), dtype=torch.float32, device=dev)
        def one():
            h.Reset()
            n1 = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), ostride, ostride, ts.cuda_stream, np.float32)
            n2 = h.flush_batch_dev(dy.data_ptr() + n1 * 4, ostride, ostride - n1, ts.cuda_stream, np.float32)
        for _ in range(2): one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(5): one()
        e1.record(ts); torch.cuda.synchronize()
        print(f"{ir}->{orr} rows {rows} float32 API, fold {fold}: {e0.elapsed_time(e1)/5:.3f} ms, kernels {h.last_kernels()}")
