"""CPU tier: the pieces of bench.py that decide whether a bench line may be trusted -- the synthetic FASTA it writes,
the rows it picks for the in-line parity check, and the check itself (it must pass on the reference's own numbers and
fail on a single flipped bit)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyoracle  # noqa: E402
from gkmqc_b200 import capi  # noqa: E402


def test_written_fasta_is_what_the_reader_reads(tmp_path):
    pos, neg = bench.write_problem(str(tmp_path), 37, seed=5)
    arr = bench.synth(37, seed=5)
    with capi.Problem(2, 11, 7, 3) as P:
        assert P.read(pos, neg) == 18 and P.n == 37
        for i in (0, 17, 18, 36):
            assert P.sid(i) == b"s%d" % i and P.seqlen(i) == 300
            assert np.array_equal(P.codes(i)[0], np.searchsorted(np.frombuffer(b"ACGT", np.uint8), arr[i]) + 1)


def test_parity_rows_cover_both_sides_of_every_boundary():
    n, cols, nblk = 50000, 12512, 4
    rows = bench.parity_rows(n, cols, nblk)
    assert len(rows) >= 64 and rows.min() >= 1 and rows.max() == n - 1 and len(set(rows)) == len(rows)
    for k in (1, 2, 3):
        assert {k * cols - 1, k * cols, k * cols + 1} <= set(rows)
    assert any(r % 148 == 0 for r in rows) and any(r % 148 == 147 for r in rows) and any(r % 592 == 0 for r in rows)
    small = bench.parity_rows(40, 40, 1)
    assert len(small) == 39 and small.min() == 1 and small.max() == 39


def test_parity_check_passes_on_the_reference_and_catches_one_bit(tmp_path):
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    n = 60
    pos, neg = bench.write_problem(str(tmp_path), n, seed=9)
    h = bench.open_reference(pos, neg)
    try:
        rows = np.arange(1, n, dtype=np.int32)
        K, H, _ = h.rows_values(rows, 2)
        kmat = np.zeros((n, n))
        hist = {}
        for i, r in enumerate(rows):
            kmat[r, :r] = K[i, :r]
            hist[int(r)] = H[i, :, :r].T.copy()
        np.fill_diagonal(kmat, 1.0)
        par, _ = bench.check_parity(h, kmat, lambda r: hist[r], n, (1, n, 0), 2, owned_only=False)
        assert par["ok"] and par["hist_bit_exact"] and par["kmat_max_rel"] == 0.0 and par["rows"] == n - 1
        bad = kmat.copy()
        bad[37, 5] = np.nextafter(bad[37, 5], 2.0)          # one ulp
        par, _ = bench.check_parity(h, bad, lambda r: hist[r], n, (1, n, 0), 2, owned_only=False)
        assert not par["kmat_bit_identical"] and 37 in par["mismatching_rows"]
        hist[12][3, 2] += 1                                   # one count
        par, _ = bench.check_parity(h, kmat, lambda r: hist[r], n, (1, n, 0), 2, owned_only=False)
        assert not par["ok"] and not par["hist_bit_exact"]
        touched = kmat.copy()
        touched[20, 30] = 0.5                                 # the upper triangle belongs to the caller
        par, _ = bench.check_parity(h, touched, lambda r: hist[r], n, (1, n, 0), 2, owned_only=False)
        assert 20 in par["mismatching_rows"]
    finally:
        h.close()
