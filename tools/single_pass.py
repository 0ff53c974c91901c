#!/usr/bin/env python
"""One Process+Flush pass of a single-stream BASELINE config (after warm-up passes) — the target of an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/single_pass.py c1
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import gar_b200 as G  # noqa: E402
from helpers import sig_c1, sig_c2, sig_c5a  # noqa: E402


def cfg(ir, orr, ch, preset):
    return G.Config(InputRate=ir, OutputRate=orr, Channels=ch, Quality=G.QualitySpec(Preset=preset))


which = sys.argv[1] if len(sys.argv) > 1 else "c1"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
io = np.float64
if which == "c1":
    h, x = G.NewEngine(44100, 48000, G.QualityHigh), sig_c1()
elif which == "c5b":
    h, x = G.NewEngine(44100, 47999, G.QualityHigh), sig_c1()
elif which == "c5a":
    h, x = G.New(cfg(8000, 192000, 1, G.QualityHigh)), sig_c5a()
elif which == "c2":
    h, x, io = G.New(cfg(48000, 44100, 1, G.QualityHigh)), sig_c2()[0][:4096 * 6], np.float32
else:
    raise SystemExit("c1 | c5a | c5b | c2")
esz = 4 if io == np.float32 else 8
dx = torch.from_numpy(np.ascontiguousarray(x, dtype=io)).cuda()
ost = (h.EstimateOutput(len(x)) + 8192 + 3) & ~3
dy = torch.zeros(ost, dtype=torch.float32 if io == np.float32 else torch.float64, device="cuda")
ts = torch.cuda.Stream()
for it in range(warm + 1):
    h.Reset()
    torch.cuda.synchronize()
    if it == warm:
        torch.cuda.cudart().cudaProfilerStart()
    if which == "c2":
        for k in range(len(x) // 4096):
            h.process_batch_dev(dx.data_ptr() + k * 4096 * esz, 4096, 4096, dy.data_ptr(), ost, ost, ts.cuda_stream, io)
    else:
        n1 = h.process_batch_dev(dx.data_ptr(), len(x), len(x), dy.data_ptr(), ost, ost, ts.cuda_stream, io)
        h.flush_batch_dev(dy.data_ptr() + n1 * esz, ost, ost - n1, ts.cuda_stream, io)
    torch.cuda.synchronize()
    if it == warm:
        torch.cuda.cudart().cudaProfilerStop()
print(which, h.last_kernels())
