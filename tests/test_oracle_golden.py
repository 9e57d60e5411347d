"""CPU tier: the oracle (oracle/gkm_oracle.c) against the golden vectors that the UNMODIFIED
reference produced (tests/golden/*.npz, generator oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import pyoracle
from conftest import GOLD, golden_names, load_golden


def test_weights_grid_bit_identical():
    W = np.load(GOLD + "/weights.npz")
    assert len(W.files) == 405
    for key in W.files:
        t, L, k = [int(x[1:]) for x in key.split("_")]
        assert np.array_equal(pyoracle.oracle_weights(t, L, k), W[key]), key


def test_known_answer_weights():
    # values probed from the reference at survey time
    assert np.array_equal(pyoracle.oracle_weights(0, 11, 7)[:4], [330, 120, 36, 8])
    np.testing.assert_array_equal(pyoracle.oracle_weights(2, 11, 7)[:4],
                                  [0.2200049901184684, 0.06041187953087501, 0.013319240119017195, 0.0017748346654116172])
    np.testing.assert_array_equal(pyoracle.oracle_weights(1, 11, 7)[:4],
                                  [0.2866954803466797, 0.06257057189941405, 0.004171371459960938, -0.0023174285888671905])
    np.testing.assert_array_equal(pyoracle.oracle_weights(4, 10, 6)[:4],
                                  [0.16959830600535497, 0.05573480820748955, 0.014950497890822588, 0.0027788987208623426])


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g, cfg, pos, neg = load_golden(name)
    o = pyoracle.Oracle(**cfg)
    n = o.read_problem(pos, neg)
    assert n == len(g["lens"]) and o.npos == int(g["npos"])
    assert [o.seqlen(i) for i in range(n)] == list(g["lens"])
    assert np.array_equal(o.weights()[: cfg["d"] + 1], g["weights"])
    assert np.array_equal(o.sqnorm(), g["sqnorm"])
    K, H = o.matrix_lower()
    assert np.array_equal(H, g["hist"]), "integer mismatch histograms"
    if cfg["kernel_type"] in (3, 5):  # exp() may differ in the last ulp between libm builds
        np.testing.assert_allclose(K, g["kmat"], rtol=1e-12, atol=0)
    else:
        assert np.array_equal(K, g["kmat"])
    for i in range(n):
        a, b = o.poswt(i)
        assert np.array_equal(a, g["poswt"][i, 0, : len(a)]) and np.array_equal(b, g["poswt"][i, 1, : len(b)])
    o.close()


def test_oracle_rect_equals_triangular():
    g, cfg, pos, neg = load_golden("mix_t4_L11k7d3")
    o = pyoracle.Oracle(**cfg)
    n = o.read_problem(pos, neg)
    K, H = o.rect(np.arange(5, n), 5)
    assert np.array_equal(K, g["kmat"][5:, :5]) and np.array_equal(H, g["hist"][5:, :5])


@pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_probe_still_agrees():
    g, cfg, pos, neg = load_golden("uni_t2_L11k7d3")
    h = pyoracle.RefHook(pos, neg, **cfg)
    try:
        assert np.array_equal(h.mmprofile(7, 7).T, g["hist"][7, :7])
        assert np.array_equal(h.row(7, 0, 7), g["kmat"][7, :7])
    finally:
        h.close()
