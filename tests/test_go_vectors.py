"""Sample-level parity against outputs of the Go reference itself (tests/go_vectors.py explains how they are made).

When tests/golden/go_vectors/result.json is present (one `go run` on a machine with Go), the CPU oracle and — under
`-m gpu` — the B200 engine are compared with the Go outputs on every case: counts bit-exact, float64 <= 1e-12,
float32 <= 1e-6. Without the file those tests skip; the harness itself is always exercised by a self-check in which a
stand-in result is produced from the oracle in a temporary directory (this proves the comparator and the file formats, not parity)."""
import json

import numpy as np
import pytest

import go_vectors as GV

HAVE = GV.load() is not None
need_go = pytest.mark.skipif(not HAVE, reason="no Go-produced vectors under tests/golden/go_vectors (see tests/go_vectors.py)")


def _write_result_from(dirpath, runner):
    """Mimics paritydump/main.go's output files using `runner` (self-check only)."""
    manifest = json.loads((dirpath / "manifest.json").read_text())
    (dirpath / "out").mkdir(exist_ok=True)
    by_name = {c["name"]: c for c in GV.cases(small=True)}
    res = []
    for m in manifest:
        c = by_name[m["name"]]
        outs, counts = runner(c)
        files = []
        for i, o in enumerate(outs):
            rel = f"out/{m['name']}_{i}.{m['dtype']}"
            (dirpath / rel).write_bytes(np.ascontiguousarray(o, dtype="<f4" if m["dtype"] == "f32" else "<f8").tobytes())
            files.append(rel)
        res.append(dict(name=m["name"], outputs=files, chunk_counts=counts, latency=0, ratio=0.0, info=""))
    (dirpath / "result.json").write_text(json.dumps({"reference": "self-check (oracle stand-in)", "cases": res}))


def test_harness_selfcheck(tmp_path):
    m = GV.write_inputs(tmp_path, small=True)
    assert len(m) == len(GV.cases(small=True)) and all(len(c["sha256"]) == len(c["inputs"]) for c in m)
    assert GV.load(tmp_path) is None  # no result.json yet
    _write_result_from(tmp_path, GV.run_oracle)
    loaded = GV.load(tmp_path)
    assert len(loaded) == len(m)
    for c, refs, ref_counts in loaded:
        got, counts = GV.run_oracle(c)
        assert GV.compare(c, got, counts, refs, ref_counts) == 0.0
    # a perturbed sample or a changed count must be caught
    c, refs, ref_counts = loaded[0]
    got, counts = GV.run_oracle(c)
    bad = [g.copy() for g in got]
    bad[0][100] += 1e-9
    with pytest.raises(AssertionError):
        GV.compare(c, bad, counts, refs, ref_counts)
    with pytest.raises(AssertionError):
        GV.compare(c, got, [[n + 1 for n in counts[0]]] + counts[1:], refs, ref_counts)


def test_paritydump_program_is_committed():
    from helpers import ROOT
    src = (ROOT / "go-audio-resampler_b200" / "go" / "paritydump" / "main.go").read_text()
    assert 'resampler "github.com/tphakala/go-audio-resampler"' in src and "manifest.json" in src
    assert (ROOT / "go-audio-resampler_b200" / "go" / "paritydump" / "go.mod").exists()


@need_go
def test_oracle_matches_the_go_reference():
    for c, refs, ref_counts in GV.load():
        got, counts = GV.run_oracle(c)
        GV.compare(c, got, counts, refs, ref_counts)


@need_go
@pytest.mark.gpu
def test_gpu_matches_the_go_reference():
    for c, refs, ref_counts in GV.load():
        got, counts = GV.run_gpu(c)
        GV.compare(c, got, counts, refs, ref_counts)


@pytest.mark.gpu
def test_gpu_runs_every_go_vector_case_like_the_oracle():
    """The same drivers on short signals, B200 engine vs oracle (always runs; the Go files only replace the referee)."""
    for c in GV.cases(small=True):
        got, counts = GV.run_gpu(c)
        want, wcounts = GV.run_oracle(c)
        GV.compare(c, got, counts, want, wcounts)
