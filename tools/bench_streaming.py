#!/usr/bin/env python
"""BASELINE config 2 as the reference runs it: streaming ProcessFloat32Into in 4096-frame chunks + Flush, two
mono instances (L/R). Reports per-chunk latency of the host-facing call and Msamples/s, next to the oracle."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "go-audio-resampler_b200" / "python"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import gar_b200 as G  # noqa: E402
from helpers import sig_c2  # noqa: E402
from oracle import oracle as O  # noqa: E402

left, right = sig_c2()
cfg = G.Config(InputRate=48000, OutputRate=44100, Channels=1, Quality=G.QualitySpec(Preset=G.QualityHigh))
for chunk in (4096, 65536):
    r = G.New(cfg)
    out = np.empty(r.EstimateOutput(chunk), dtype=np.float32)
    for rep in range(3):
        r.Reset()
        G.kernel_launches(reset=True)
        t0 = time.perf_counter()
        tot = 0
        for x in (left, right):
            r.Reset()
            for i in range(0, len(x), chunk):
                tot += r.ProcessFloat32Into(x[i:i + chunk], out)
            tot += len(r.Flush())
        dt = time.perf_counter() - t0
    nchunks = 2 * ((len(left) + chunk - 1) // chunk)
    print(f"GPU  chunk={chunk:6d}: {dt*1e3:8.2f} ms total, {dt/nchunks*1e6:7.1f} us/chunk, {tot/dt/1e6:8.1f} Msamples/s out, "
          f"{G.kernel_launches()/nchunks:.1f} launches/chunk")
p = O.Pipeline(48000, 44100, 1, O.PRESET_HIGH)
ref = np.empty(p.estimate_output(4096), dtype=np.float32)
t0 = time.perf_counter()
tot = 0
for x in (left, right):
    p.reset()
    for i in range(0, len(x), 4096):
        tot += p.process_f32_into(x[i:i + 4096], ref)
    tot += len(p.flush())
dt = time.perf_counter() - t0
print(f"CPU oracle (1 thread) chunk=4096: {dt*1e3:8.2f} ms total, {tot/dt/1e6:8.1f} Msamples/s out")
