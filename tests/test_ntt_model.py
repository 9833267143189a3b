"""Validates the index algebra that csrc/ntt.cu implements (tests/ntt_model.py) against the O(n^2)
DFT with the reference's semantics (pyref.dft).  CPU only."""
import random

import pytest

import ntt_model
import pyref as P


@pytest.mark.parametrize("kind", ["fft", "ifft", "coset_fft", "coset_ifft"])
@pytest.mark.parametrize("log_n,max_deg,log_c", [(0, 3, 1), (1, 3, 2), (2, 3, 2), (3, 3, 1), (4, 3, 1), (5, 3, 2),
                                                 (6, 2, 1), (6, 3, 2), (7, 3, 2), (7, 4, 1), (5, 8, 2), (6, 5, 3)])
def test_model_matches_dft(kind, log_n, max_deg, log_c):
    rng = random.Random(log_n * 100 + max_deg)
    vals = [rng.randrange(P.R_MOD) for _ in range(1 << log_n)]
    assert ntt_model.run(vals, kind, max_deg, log_c) == P.dft(vals, kind)
