"""Go-reference vector harness (shared by tests/golden/make_go_inputs.py and tests/test_go_vectors.py).

The reference ships no sample vectors and no Go toolchain exists in the build image, so sample-level
parity against the *Go binary* is closed by one command on any machine that has Go:

    python tests/golden/make_go_inputs.py                 # writes tests/golden/go_vectors/{manifest.json,in/*.f64|f32}
    (cd go-audio-resampler_b200/go/paritydump && go mod tidy && go run . -dir ../../../tests/golden/go_vectors)
    python -m pytest tests/test_go_vectors.py             # oracle (CPU) and B200 engine (-m gpu) vs the Go outputs

Inputs travel as raw little-endian files because Go's math.Sin and numpy's sin may differ in the last ulp: both sides must
read the same bits. `CASES` are the five BASELINE configs on SURVEY.md §8(d)'s synthetic inputs (C4: streams 0..3 of the 4096,
all three presets) plus the reference's published-THD cases and two small streaming cases.
"""
import hashlib
import json
from pathlib import Path

import numpy as np

from helpers import G, O, sig_c1, sig_c2, sig_c3, sig_c4, sig_c5a

DEFAULT_DIR = Path(__file__).resolve().parent / "golden" / "go_vectors"
TOL = {"f64": 1e-12, "f32": 1e-6}  # BASELINE.json north_star tolerances
DT = {"f64": np.float64, "f32": np.float32}


def _thd_sine(n=65536, rate=44100.0):  # measureTHDInternal's input (quality_regression_test.go:298-303)
    return 0.9 * np.sin(2.0 * np.pi * 1000.0 * np.arange(n) / rate)


def cases(small=False):
    """name, api, rates, preset, dtype, chunk, multi, inputs (arrays). small=True: short signals for the self-check."""
    k = 20 if small else 1
    c2l, c2r = sig_c2(480000 // k)
    out = [
        dict(name="c1", api="engine", in_rate=44100, out_rate=48000, preset=3, dtype="f64", chunk=0, multi=False,
             inputs=[sig_c1(441000 // k)]),
        dict(name="c2", api="pipeline", in_rate=48000, out_rate=44100, preset=3, dtype="f32", chunk=4096, multi=False,
             inputs=[c2l, c2r]),
        dict(name="c3", api="pipeline", in_rate=96000, out_rate=48000, preset=4, dtype="f64", chunk=0, multi=True,
             inputs=sig_c3(960000 // k, 8)),
        dict(name="c5a", api="pipeline", in_rate=8000, out_rate=192000, preset=3, dtype="f64", chunk=0, multi=False,
             inputs=[sig_c5a(80000 // k)]),
        dict(name="c5b", api="engine", in_rate=44100, out_rate=47999, preset=3, dtype="f64", chunk=0, multi=False,
             inputs=[sig_c1(441000 // k)]),
    ]
    x4 = sig_c4(4, 480000 // k)
    for nm, p in (("low", 1), ("medium", 2), ("high", 3)):
        out.append(dict(name=f"c4_{nm}", api="engine", in_rate=48000, out_rate=16000, preset=p, dtype="f32", chunk=0,
                        multi=False, inputs=[x4[i] for i in range(4)]))
    # the reference's THD procedure (README.md:303-308 publishes the Go results) through the public API
    for nm, p in (("low", 1), ("medium", 2), ("high", 3), ("veryhigh", 4)):
        out.append(dict(name=f"thd_{nm}", api="engine", in_rate=44100, out_rate=48000, preset=p, dtype="f64", chunk=0,
                        multi=False, inputs=[_thd_sine(65536 // k)]))
    # float32 engine at a non-integer ratio (NewEngineFloat32, convenience.go:329-366) and chunked streaming cases
    out.append(dict(name="f32_cd2dat", api="engine", in_rate=44100, out_rate=48000, preset=3, dtype="f32", chunk=0,
                    multi=False, inputs=[sig_c1(88200 // k).astype(np.float32)]))
    out.append(dict(name="chunk_dec", api="engine", in_rate=48000, out_rate=32000, preset=4, dtype="f64", chunk=1000,
                    multi=False, inputs=[_thd_sine(48000 // k, 48000.0)]))
    out.append(dict(name="quick", api="pipeline", in_rate=44100, out_rate=48000, preset=0, dtype="f64", chunk=777,
                    multi=False, inputs=[_thd_sine(44100 // k)]))
    return out


def write_inputs(dirpath=DEFAULT_DIR, small=False):
    d = Path(dirpath)
    (d / "in").mkdir(parents=True, exist_ok=True)
    manifest = []
    for c in cases(small):
        files = []
        for i, x in enumerate(c["inputs"]):
            a = np.ascontiguousarray(x, dtype="<" + ("f4" if c["dtype"] == "f32" else "f8"))
            rel = f"in/{c['name']}_{i}.{c['dtype']}"
            (d / rel).write_bytes(a.tobytes())
            files.append(rel)
        m = {k: v for k, v in c.items() if k != "inputs"}
        m["inputs"] = files
        m["sha256"] = [hashlib.sha256((d / f).read_bytes()).hexdigest() for f in files]
        manifest.append(m)
    (d / "manifest.json").write_text(json.dumps(manifest, indent=1))
    return manifest


def load(dirpath=DEFAULT_DIR):
    """-> list of (case dict with 'inputs' arrays, reference outputs [arrays], chunk_counts) or None if no Go run present."""
    d = Path(dirpath)
    if not (d / "result.json").exists() or not (d / "manifest.json").exists():
        return None
    manifest = json.loads((d / "manifest.json").read_text())
    result = {r["name"]: r for r in json.loads((d / "result.json").read_text())["cases"]}
    out = []
    for m in manifest:
        r = result[m["name"]]
        dt = np.dtype("<f4" if m["dtype"] == "f32" else "<f8")
        ins = [np.frombuffer((d / f).read_bytes(), dtype=dt).astype(DT[m["dtype"]]) for f in m["inputs"]]
        refs = [np.frombuffer((d / f).read_bytes(), dtype=dt).astype(DT[m["dtype"]]) for f in r["outputs"]]
        out.append((dict(m, inputs=ins), refs, r["chunk_counts"]))
    return out


def _chunks(x, chunk):
    return [x] if chunk <= 0 else [x[i:i + chunk] for i in range(0, len(x), chunk)]


def run_oracle(c):
    """The case through the CPU oracle, the way paritydump/main.go drives the Go API -> (outputs, chunk_counts)."""
    dt = DT[c["dtype"]]
    outs, counts = [], []
    if c["api"] == "pipeline" and c["multi"]:
        p = O.Pipeline(c["in_rate"], c["out_rate"], len(c["inputs"]), c["preset"])
        ys, fs = p.process_multi(c["inputs"]), p.flush_multi()
        return [np.concatenate([a, b]) for a, b in zip(ys, fs)], [[len(a), len(b)] for a, b in zip(ys, fs)]
    for x in c["inputs"]:
        parts, cnt = [], []
        if c["api"] == "engine":
            e = O.Engine(c["in_rate"], c["out_rate"], O.preset_to_engine_quality(c["preset"]), dt)
            for ch in _chunks(x, c["chunk"]):
                y = e.process(ch)
                parts.append(y)
                cnt.append(len(y))
            f = e.flush()
        else:
            p = O.Pipeline(c["in_rate"], c["out_rate"], 1, c["preset"])
            for ch in _chunks(x, c["chunk"]):
                if dt == np.float32:
                    buf = np.empty(p.estimate_output(len(ch)), dtype=np.float32)
                    n = p.process_f32_into(ch, buf)
                    y = buf[:n].copy()
                else:
                    y = p.process(ch)
                parts.append(y)
                cnt.append(len(y))
            f = p.flush().astype(dt)
        outs.append(np.concatenate(parts + [f]))
        counts.append(cnt + [len(f)])
    return outs, counts


def run_gpu(c):
    """The same case through the B200 engine's Go-API mirror."""
    dt = DT[c["dtype"]]
    outs, counts = [], []
    if c["api"] == "pipeline" and c["multi"]:
        r = G.New(G.Config(InputRate=c["in_rate"], OutputRate=c["out_rate"], Channels=len(c["inputs"]),
                           Quality=G.QualitySpec(Preset=c["preset"]), EnableParallel=True))
        ys, fs = r.ProcessMulti(c["inputs"]), r.FlushMulti()
        return [np.concatenate([a, b]) for a, b in zip(ys, fs)], [[len(a), len(b)] for a, b in zip(ys, fs)]
    for x in c["inputs"]:
        parts, cnt = [], []
        if c["api"] == "engine":
            r = G.SimpleResampler(c["in_rate"], c["out_rate"], c["preset"], dt)
            for ch in _chunks(x, c["chunk"]):
                y = r.Process(ch)
                parts.append(y)
                cnt.append(len(y))
            f = r.Flush()
        else:
            r = G.New(G.Config(InputRate=c["in_rate"], OutputRate=c["out_rate"], Channels=1,
                               Quality=G.QualitySpec(Preset=c["preset"])))
            for ch in _chunks(x, c["chunk"]):
                y = r.ProcessFloat32(ch) if dt == np.float32 else r.Process(ch)
                parts.append(y)
                cnt.append(len(y))
            f = r.Flush().astype(dt)
        outs.append(np.concatenate(parts + [f]))
        counts.append(cnt + [len(f)])
    return outs, counts


def compare(c, got, got_counts, refs, ref_counts):
    """Counts bit-exact, samples within the north-star tolerance. Returns max |err| over the case."""
    assert [list(map(int, a)) for a in got_counts] == [list(map(int, a)) for a in ref_counts], c["name"]
    worst = 0.0
    for g, r in zip(got, refs):
        assert len(g) == len(r), (c["name"], len(g), len(r))
        if len(g):
            worst = max(worst, float(np.max(np.abs(g.astype(np.float64) - r.astype(np.float64)))))
    assert worst <= TOL[c["dtype"]], (c["name"], worst)
    return worst
