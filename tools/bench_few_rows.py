import sys, numpy as np, torch
sys.path.insert(0, 'go-audio-resampler_b200/python')
import gar_b200 as G
dev = torch.device('cuda', 0); ts = torch.cuda.Stream(device=dev)
for (ir, orr) in [(44100, 48000), (48000, 44100)]:
  for rows in (1, 2, 4, 8, 16, 24):
    for dt in (np.float32, np.float64):
        n = 441000 if ir == 44100 else 480000
        h = G.NewBatch(ir, orr, G.QualityHigh, rows, dt)
        x = np.random.default_rng(0).standard_normal((rows, n)).astype(dt)
        tdt = torch.float32 if dt == np.float32 else torch.float64
        esz = 4 if dt == np.float32 else 8
        dx = torch.from_numpy(x).to(dev)
        ostride = (h.EstimateOutput(n) + 8192 + 3) & ~3
        dy = torch.zeros((rows, ostride), dtype=tdt, device=dev)
        def one():
            h.Reset()
            n1 = h.process_batch_dev(dx.data_ptr(), n, n, dy.data_ptr(), ostride, ostride, ts.cuda_stream, dt)
            n2 = h.flush_batch_dev(dy.data_ptr() + n1 * esz, ostride, ostride - n1, ts.cuda_stream, dt)
        for _ in range(3): one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(20): one()
        e1.record(ts); torch.cuda.synchronize()
        print(f"{ir}->{orr} rows {rows:2d} {np.dtype(dt).name}: {e0.elapsed_time(e1)/20*1e3:8.1f} us  {h.last_kernels()}")
