"""CPU tier: gkm_bitslice.h -- the arithmetic the sm_100a kernel runs per lane -- executed serially by the
lane emulator (tests/emu/diag_emu.cc) over the product's own packed image, against the golden histograms."""
import numpy as np
import pytest

import pyoracle
from conftest import golden_names, load_golden, random_seqs
from gkmqc_b200 import capi


@pytest.mark.parametrize("name", golden_names())
def test_emulated_lanes_match_reference_histograms(name, emu_lib):
    g, cfg, pos, neg = load_golden(name)
    P = capi.Problem(lib=emu_lib, **cfg)
    assert P.read(pos, neg) == int(g["npos"])
    n = P.n
    H = np.zeros((n, n, cfg["d"] + 1), np.int32)
    assert emu_lib.gkm_emu_hist_lower(P.h, H.ctypes.data_as(capi.c_i32_p)) == 0
    w = P.weights()
    sq = np.array([np.sqrt(sum(w[m] * H[a, a, m] for m in range(len(w)))) for a in range(n)])
    assert np.array_equal(sq, g["sqnorm"]), "diagonal histograms (sqnorm)"
    for a in range(n):
        H[a, a:, :] = 0
    assert np.array_equal(H, g["hist"])
    P.close()


@pytest.mark.parametrize("L,d", [(10, 3), (11, 3), (12, 4), (15, 5), (16, 4), (5, 5), (7, 2)])
def test_emulated_lanes_random_ragged(L, d, emu_lib):
    seqs = random_seqs(6, 130, seed=100 * L + d, ragged=True)
    seqs = [s if len(s) >= L else s + "ACGT" * 5 for s in seqs]
    k = L - d
    o = pyoracle.Oracle(4, L, k, d)
    P = capi.Problem(4, L, k, d, lib=emu_lib)
    for s in seqs:
        o.add(s)
        P.add(s)
    h = np.zeros(d + 1, np.int32)
    for a in range(len(seqs)):
        for b in range(len(seqs)):
            assert emu_lib.gkm_emu_hist(P.h, a, b, h.ctypes.data_as(capi.c_i32_p)) == 0
            assert np.array_equal(h, o.hist(a, b)), (a, b)
