"""K5 (kernels_chain.cu) schedules the polyphase stage in chunks of S intermediate samples: a 64-output tile belongs to the first
chunk that contains the end of everything its staging reads, and the kernel's dependency counters rely on three facts, checked
here on the CPU for random geometries (polyphase_stage.go:186-312 arithmetic: at_n = at0 + n*step, div_n = (at_n >> 16) / L):
  1. the assignment is monotone and covers every tile exactly once;
  2. a tile of chunk c < last reads nothing at or beyond intermediate sample (c + 1) * S  (it may start when the x2 items of
     chunks <= c have finished);
  3. a tile of chunk c reads nothing below (c - 1) * S  (S >= 1024 > 63 * 8 + K + 9; a ring slot may be overwritten once the
     polyphase chunks up to c - R + 2 have finished: the readers of slot c - R are chunks c - R and c - R + 1, one to spare)."""
import numpy as np
import pytest

from helpers import G

TO = 64  # outputs per tile


@pytest.mark.parametrize("seed", range(12))
def test_polyphase_tiles_fall_into_the_chunks_the_dependency_counters_assume(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.integers(2, 400))
    r = float(rng.uniform(0.3, 7.9))                       # intermediate samples per output
    step = int(r * L * 65536) | (int(rng.integers(0, 2)) * int(rng.integers(1, 65535)))
    r = step / (L * 65536.0)
    taps = int(rng.integers(8, 200))
    kp = ((int(np.ceil(7 * r)) + 1 + taps + 3) // 4) * 4   # K of the coefficient matrices (launch_chain_t)
    hp = int(rng.integers(0, taps + 4))                    # carried tail of the polyphase stage
    at0 = int(rng.integers(0, L << 16))
    S = int(rng.choice([1024, 2048, 4096]))
    n_mid = int(rng.integers(40 * S, 60 * S))
    NC = (n_mid + S - 1) // S
    num_in = hp + n_mid - taps + 1
    n_out = int((((num_in * L) << 16) - at0 + step - 1) // step)
    n_tiles = (n_out + TO - 1) // TO
    hi = [int(G.lib().gar_debug_chain_tile_hi(S, kp, NC, n_tiles, c, hp, L, at0, step, n_out)) for c in range(NC)]
    assert hi[-1] == n_tiles and all(b >= a for a, b in zip(hi, hi[1:])) and hi[0] >= 0
    span_max = int(np.ceil((TO - 1) * r)) + 1 + kp + 8

    def d(n):
        return ((at0 + n * step) >> 16) // L

    lo = 0
    for c in range(NC):
        for t in ({lo, hi[c] - 1} if hi[c] > lo else ()):
            n0, n1 = t * TO, min(n_out, t * TO + TO)
            first = d(n0) - hp                                            # intermediate index of the first staged sample
            last = d(n1 - 1) - hp + kp + 4 + 2                            # one past the last one (span + alignment pad)
            assert last - first <= span_max + 2
            if c < NC - 1:
                assert last <= (c + 1) * S, (c, t, last)
            if c > 0:
                assert first >= (c - 1) * S, (c, t, first)
        lo = hi[c]
