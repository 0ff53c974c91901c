#!/usr/bin/env python
"""Writes the input files + manifest the Go parity-dump program (go-audio-resampler_b200/go/paritydump) consumes.

    python tests/golden/make_go_inputs.py [dir]      # default tests/golden/go_vectors
See tests/go_vectors.py for the whole procedure."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import go_vectors  # noqa: E402

if __name__ == "__main__":
    d = Path(sys.argv[1]) if len(sys.argv) > 1 else go_vectors.DEFAULT_DIR
    m = go_vectors.write_inputs(d)
    print(f"{len(m)} cases, {sum(len(c['inputs']) for c in m)} input files under {d}/in; now run the Go program:")
    print(f"  (cd go-audio-resampler_b200/go/paritydump && go mod tidy && go run . -dir {d})")
