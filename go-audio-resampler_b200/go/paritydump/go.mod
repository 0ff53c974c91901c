// Parity-dump program: runs the UNMODIFIED reference (github.com/tphakala/go-audio-resampler) on the input files written by
// tests/golden/make_go_inputs.py and writes its outputs for tests/test_go_vectors.py.
//
//   cd go-audio-resampler_b200/go/paritydump
//   go mod tidy            # fetches the reference at the pinned tag (or: go mod edit -replace below for a local checkout)
//   go run . -dir ../../../tests/golden/go_vectors
//
// With a local checkout of the reference:
//   go mod edit -replace github.com/tphakala/go-audio-resampler=/path/to/go-audio-resampler
module gar-b200/paritydump

go 1.26

require github.com/tphakala/go-audio-resampler v1.4.0
