// Compiled and run by tests/test_cpp_mirror.py: the C++ mirror (include/gar.hpp) links against the C ABI and
// its error mapping works; geometry-only handles (Device = -1) so no GPU is needed.
#include <cstdio>
#include "gar.hpp"
int main() {
    using namespace resampler;
    Config c;
    c.InputRate = 48000; c.OutputRate = 44100; c.Channels = 1; c.Quality.Preset = QualityHigh; c.Device = -1;
    auto r = New(c);
    if (r->EstimateOutput(4096) != 3827) return 1;
    if (gar_next_output_count(r->raw(), 0, 4096) != 3556) return 2;
    bool thrown = false;
    try { Config bad = c; bad.Channels = 0; New(bad); } catch (const ErrInvalidConfig&) { thrown = true; }
    if (!thrown) return 3;
    thrown = false;
    try { double in[8] = {0}, out[4]; r->ProcessInto(in, 8, out, 4); } catch (const ErrBufferTooSmall&) { thrown = true; }
    if (!thrown) return 4;
    thrown = false;
    try { r->Process(std::vector<double>(100, 0.0)); } catch (const CudaError&) { thrown = true; }  // no CPU fallback
    if (!thrown) return 5;
    std::puts("cpp mirror ok");
    return 0;
}
