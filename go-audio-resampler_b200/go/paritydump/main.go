// paritydump runs the unmodified Go reference on the inputs listed in <dir>/manifest.json and writes, per case,
// the raw little-endian output samples and the per-call sample counts into <dir>/out/ plus <dir>/result.json.
// tests/test_go_vectors.py then compares the CPU oracle and the B200 engine with these files sample by sample
// (float64: 1e-12, float32: 1e-6, counts exact) — the sample-level pin the reference itself does not ship.
//
// Only the reference's public API is used (resample.go, convenience.go); nothing here is product code.
package main

import (
	"encoding/binary"
	"encoding/json"
	"flag"
	"fmt"
	"math"
	"os"
	"path/filepath"
	"runtime"

	resampler "github.com/tphakala/go-audio-resampler"
)

// One case of manifest.json (written by tests/golden/make_go_inputs.py).
type manifestCase struct {
	Name     string   `json:"name"`
	API      string   `json:"api"`      // "engine" (NewEngine / NewEngineFloat32) or "pipeline" (New(Config))
	InRate   float64  `json:"in_rate"`
	OutRate  float64  `json:"out_rate"`
	Preset   int      `json:"preset"`   // resampler.QualityPreset value
	Dtype    string   `json:"dtype"`    // "f64" or "f32" (I/O type; path A computes in float64 either way)
	Chunk    int      `json:"chunk"`    // 0: one Process call; >0: ProcessInto / ProcessFloat32Into in chunks of this many frames
	Multi    bool     `json:"multi"`    // pipeline only: ProcessMulti + FlushMulti over all inputs as channels (EnableParallel)
	Inputs   []string `json:"inputs"`   // raw little-endian files, one per channel / stream
}

type caseResult struct {
	Name        string    `json:"name"`
	Outputs     []string  `json:"outputs"`       // one raw file per input: Process output followed by the Flush output
	ChunkCounts [][]int   `json:"chunk_counts"`  // per input: samples returned by every Process call, then by Flush
	Latency     int       `json:"latency"`
	Ratio       float64   `json:"ratio"`
	Info        string    `json:"info"`
}

func readF64(path string) ([]float64, error) {
	b, err := os.ReadFile(path)
	if err != nil {
		return nil, err
	}
	out := make([]float64, len(b)/8)
	for i := range out {
		out[i] = math.Float64frombits(binary.LittleEndian.Uint64(b[8*i:]))
	}
	return out, nil
}

func readF32(path string) ([]float32, error) {
	b, err := os.ReadFile(path)
	if err != nil {
		return nil, err
	}
	out := make([]float32, len(b)/4)
	for i := range out {
		out[i] = math.Float32frombits(binary.LittleEndian.Uint32(b[4*i:]))
	}
	return out, nil
}

func writeF64(path string, v []float64) error {
	b := make([]byte, 8*len(v))
	for i, x := range v {
		binary.LittleEndian.PutUint64(b[8*i:], math.Float64bits(x))
	}
	return os.WriteFile(path, b, 0o644)
}

func writeF32(path string, v []float32) error {
	b := make([]byte, 4*len(v))
	for i, x := range v {
		binary.LittleEndian.PutUint32(b[4*i:], math.Float32bits(x))
	}
	return os.WriteFile(path, b, 0o644)
}

// ---- path B: NewEngine / NewEngineFloat32 (convenience.go:125-186, 329-395) ----

func runEngineF64(c *manifestCase, in []float64) (out []float64, counts []int, ratio float64, err error) {
	r, err := resampler.NewEngine(c.InRate, c.OutRate, resampler.QualityPreset(c.Preset))
	if err != nil {
		return nil, nil, 0, err
	}
	if c.Chunk <= 0 {
		y, perr := r.Process(in)
		if perr != nil {
			return nil, nil, 0, perr
		}
		out = append(out, y...)
		counts = append(counts, len(y))
	} else {
		buf := make([]float64, r.EstimateOutput(c.Chunk))
		for off := 0; off < len(in); off += c.Chunk {
			end := min(off+c.Chunk, len(in))
			n, perr := r.ProcessInto(in[off:end], buf)
			if perr != nil {
				return nil, nil, 0, perr
			}
			out = append(out, buf[:n]...)
			counts = append(counts, n)
		}
	}
	f, err := r.Flush()
	if err != nil {
		return nil, nil, 0, err
	}
	out = append(out, f...)
	counts = append(counts, len(f))
	return out, counts, r.GetRatio(), nil
}

func runEngineF32(c *manifestCase, in []float32) (out []float32, counts []int, ratio float64, err error) {
	r, err := resampler.NewEngineFloat32(c.InRate, c.OutRate, resampler.QualityPreset(c.Preset))
	if err != nil {
		return nil, nil, 0, err
	}
	if c.Chunk <= 0 {
		y, perr := r.Process(in)
		if perr != nil {
			return nil, nil, 0, perr
		}
		out = append(out, y...)
		counts = append(counts, len(y))
	} else {
		buf := make([]float32, r.EstimateOutput(c.Chunk))
		for off := 0; off < len(in); off += c.Chunk {
			end := min(off+c.Chunk, len(in))
			n, perr := r.ProcessInto(in[off:end], buf)
			if perr != nil {
				return nil, nil, 0, perr
			}
			out = append(out, buf[:n]...)
			counts = append(counts, n)
		}
	}
	f, err := r.Flush()
	if err != nil {
		return nil, nil, 0, err
	}
	out = append(out, f...)
	counts = append(counts, len(f))
	return out, counts, r.GetRatio(), nil
}

// ---- path A: New(Config) (resample.go:272, constant.go:88-404) ----

type intoF64 interface {
	ProcessInto(input, output []float64) (int, error)
	EstimateOutput(inputLen int) int
}
type intoF32 interface {
	ProcessFloat32Into(input, output []float32) (int, error)
	EstimateOutput(inputLen int) int
}

func newPipeline(c *manifestCase, channels int) (resampler.Resampler, error) {
	return resampler.New(&resampler.Config{
		InputRate:      c.InRate,
		OutputRate:     c.OutRate,
		Channels:       channels,
		Quality:        resampler.QualitySpec{Preset: resampler.QualityPreset(c.Preset)},
		EnableParallel: c.Multi,
	})
}

func runPipelineF64(c *manifestCase, in []float64) (out []float64, counts []int, lat int, ratio float64, info string, err error) {
	r, err := newPipeline(c, 1)
	if err != nil {
		return
	}
	if c.Chunk <= 0 {
		y, perr := r.Process(in)
		if perr != nil {
			err = perr
			return
		}
		out = append(out, y...)
		counts = append(counts, len(y))
	} else {
		pi, ok := r.(intoF64)
		if !ok {
			err = fmt.Errorf("resampler does not implement ProcessInto")
			return
		}
		buf := make([]float64, pi.EstimateOutput(c.Chunk))
		for off := 0; off < len(in); off += c.Chunk {
			end := min(off+c.Chunk, len(in))
			n, perr := pi.ProcessInto(in[off:end], buf)
			if perr != nil {
				err = perr
				return
			}
			out = append(out, buf[:n]...)
			counts = append(counts, n)
		}
	}
	f, ferr := r.Flush()
	if ferr != nil {
		err = ferr
		return
	}
	out = append(out, f...)
	counts = append(counts, len(f))
	return out, counts, r.GetLatency(), r.GetRatio(), fmt.Sprintf("%+v", resampler.GetInfo(r)), nil
}

// float32 I/O through the float64 pipeline; Flush returns float64 (constant.go:349), stored here as float32 like the
// chunks so that one file holds the whole stream (the conversion is exact for the comparison: the test casts the same way).
func runPipelineF32(c *manifestCase, in []float32) (out []float32, counts []int, lat int, ratio float64, info string, err error) {
	r, err := newPipeline(c, 1)
	if err != nil {
		return
	}
	if c.Chunk <= 0 {
		y, perr := r.ProcessFloat32(in)
		if perr != nil {
			err = perr
			return
		}
		out = append(out, y...)
		counts = append(counts, len(y))
	} else {
		pi, ok := r.(intoF32)
		if !ok {
			err = fmt.Errorf("resampler does not implement ProcessFloat32Into")
			return
		}
		buf := make([]float32, pi.EstimateOutput(c.Chunk))
		for off := 0; off < len(in); off += c.Chunk {
			end := min(off+c.Chunk, len(in))
			n, perr := pi.ProcessFloat32Into(in[off:end], buf)
			if perr != nil {
				err = perr
				return
			}
			out = append(out, buf[:n]...)
			counts = append(counts, n)
		}
	}
	f, ferr := r.Flush()
	if ferr != nil {
		err = ferr
		return
	}
	for _, v := range f {
		out = append(out, float32(v))
	}
	counts = append(counts, len(f))
	return out, counts, r.GetLatency(), r.GetRatio(), fmt.Sprintf("%+v", resampler.GetInfo(r)), nil
}

func runPipelineMulti(c *manifestCase, in [][]float64) (outs [][]float64, counts [][]int, lat int, ratio float64, info string, err error) {
	r, err := newPipeline(c, len(in))
	if err != nil {
		return
	}
	ys, err := r.ProcessMulti(in)
	if err != nil {
		return
	}
	mf, ok := r.(resampler.MultiFlusher)
	if !ok {
		err = fmt.Errorf("resampler does not implement FlushMulti")
		return
	}
	fs, err := mf.FlushMulti()
	if err != nil {
		return
	}
	for ch := range ys {
		o := append([]float64{}, ys[ch]...)
		o = append(o, fs[ch]...)
		outs = append(outs, o)
		counts = append(counts, []int{len(ys[ch]), len(fs[ch])})
	}
	return outs, counts, r.GetLatency(), r.GetRatio(), fmt.Sprintf("%+v", resampler.GetInfo(r)), nil
}

func main() {
	dir := flag.String("dir", "go_vectors", "directory holding manifest.json and in/ (outputs go to out/ and result.json)")
	flag.Parse()
	mb, err := os.ReadFile(filepath.Join(*dir, "manifest.json"))
	if err != nil {
		fmt.Fprintln(os.Stderr, "manifest:", err)
		os.Exit(1)
	}
	var cases []manifestCase
	if err := json.Unmarshal(mb, &cases); err != nil {
		fmt.Fprintln(os.Stderr, "manifest:", err)
		os.Exit(1)
	}
	outDir := filepath.Join(*dir, "out")
	if err := os.MkdirAll(outDir, 0o755); err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	var results []caseResult
	for ci := range cases {
		c := &cases[ci]
		res := caseResult{Name: c.Name}
		fail := func(e error) {
			fmt.Fprintf(os.Stderr, "case %s: %v\n", c.Name, e)
			os.Exit(1)
		}
		switch {
		case c.API == "pipeline" && c.Multi:
			var ins [][]float64
			for _, f := range c.Inputs {
				x, e := readF64(filepath.Join(*dir, f))
				if e != nil {
					fail(e)
				}
				ins = append(ins, x)
			}
			outs, counts, lat, ratio, info, e := runPipelineMulti(c, ins)
			if e != nil {
				fail(e)
			}
			for ch, o := range outs {
				name := fmt.Sprintf("%s_%d.f64", c.Name, ch)
				if e := writeF64(filepath.Join(outDir, name), o); e != nil {
					fail(e)
				}
				res.Outputs = append(res.Outputs, "out/"+name)
			}
			res.ChunkCounts, res.Latency, res.Ratio, res.Info = counts, lat, ratio, info
		default:
			for i, f := range c.Inputs {
				p := filepath.Join(*dir, f)
				if c.Dtype == "f32" {
					x, e := readF32(p)
					if e != nil {
						fail(e)
					}
					var o []float32
					var cnt []int
					if c.API == "engine" {
						o, cnt, res.Ratio, e = runEngineF32(c, x)
					} else {
						o, cnt, res.Latency, res.Ratio, res.Info, e = runPipelineF32(c, x)
					}
					if e != nil {
						fail(e)
					}
					name := fmt.Sprintf("%s_%d.f32", c.Name, i)
					if e := writeF32(filepath.Join(outDir, name), o); e != nil {
						fail(e)
					}
					res.Outputs = append(res.Outputs, "out/"+name)
					res.ChunkCounts = append(res.ChunkCounts, cnt)
				} else {
					x, e := readF64(p)
					if e != nil {
						fail(e)
					}
					var o []float64
					var cnt []int
					if c.API == "engine" {
						o, cnt, res.Ratio, e = runEngineF64(c, x)
					} else {
						o, cnt, res.Latency, res.Ratio, res.Info, e = runPipelineF64(c, x)
					}
					if e != nil {
						fail(e)
					}
					name := fmt.Sprintf("%s_%d.f64", c.Name, i)
					if e := writeF64(filepath.Join(outDir, name), o); e != nil {
						fail(e)
					}
					res.Outputs = append(res.Outputs, "out/"+name)
					res.ChunkCounts = append(res.ChunkCounts, cnt)
				}
			}
		}
		results = append(results, res)
		fmt.Printf("%-12s %d stream(s) done\n", c.Name, len(res.Outputs))
	}
	rb, _ := json.MarshalIndent(map[string]any{
		"reference": "github.com/tphakala/go-audio-resampler",
		"go":        runtime.Version(),
		"goarch":    runtime.GOARCH,
		"cases":     results,
	}, "", " ")
	if err := os.WriteFile(filepath.Join(*dir, "result.json"), rb, 0o644); err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
}
