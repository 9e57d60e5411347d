"""GPU tier: the stand-alone `gkmkern` CLI on the new engine -- the reference CLI's
output format (src/gkmkern_main.c:221-228) with all rows present and a loss-free number format on request."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "gkmqc_b200", "bin", "gkmkern")


def parse_tsv(path, n):
    K = np.zeros((n, n))
    with open(path) as f:
        for a, line in enumerate(f):
            cells = line.rstrip("\n").split("\t")
            assert cells[-1] == "" and cells[-2] == "1.0" and len(cells) == a + 2
            K[a, :a] = [float(x) for x in cells[:a]]
            K[a, a] = 1.0
    assert a == n - 1, "every row must be written (the reference drops the last N mod 4 rows)"
    return K


def test_cli_lossless_and_reference_format(tmp_path):
    g, cfg, pos, neg = load_golden("mix_t4_L11k7d3")   # 15 sequences: 15 mod 4 != 0
    n = len(g["lens"])
    out = tmp_path / "k.tsv"
    args = [CLI, "-t", "4", "-l", "11", "-k", "7", "-d", "3", "-v", "0"]
    subprocess.check_call(args + ["-p", "17", pos, neg, str(out)])
    assert np.array_equal(parse_tsv(out, n), g["kmat"])
    subprocess.check_call(args + [pos, neg, str(out)])            # default "%e" like the reference
    np.testing.assert_allclose(parse_tsv(out, n), g["kmat"], rtol=1e-6)
    subprocess.check_call(args + ["-b", pos, neg, str(out)])
    raw = np.fromfile(out, dtype=np.uint8)
    assert np.frombuffer(raw[:4].tobytes(), np.int32)[0] == n
    tri = np.frombuffer(raw[4:].tobytes(), np.float64)
    assert np.array_equal(tri, g["kmat"][np.tril_indices(n, -1)])


def test_cli_defaults_are_the_reference_cli_defaults(tmp_path):
    # no options = EST_TRUNC, L=10 k=6 d=3 (src/gkmkern_main.c:99-107); expected values from the oracle
    import pyoracle
    g, cfg, pos, neg = load_golden("uni_t2_L11k7d3")
    o = pyoracle.Oracle(2, 10, 6, 3)
    n = o.read_problem(pos, neg)
    out = tmp_path / "k.tsv"
    subprocess.check_call([CLI, "-v", "0", "-p", "17", pos, neg, str(out)])
    assert np.array_equal(parse_tsv(out, n), o.matrix_lower(False)[0])


def test_cli_errors(tmp_path):
    out = tmp_path / "k.tsv"
    r = subprocess.run([CLI, "-l", "17", "a", "b", str(out)], capture_output=True)
    assert r.returncode == 1
    r = subprocess.run([CLI, "-v", "0", str(tmp_path / "missing.fa"), str(tmp_path / "missing.fa"), str(out)], capture_output=True)
    assert r.returncode == 1
    assert subprocess.run([CLI], capture_output=True).returncode == 1
