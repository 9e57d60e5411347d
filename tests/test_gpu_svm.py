"""GPU tier for the consumer of the kernel matrix: the C-SVC cross-validation of
scripts/gkmsvm.py:104-160 on the GPU against sklearn's SVC(kernel="precomputed") -- sklearn's bundled libsvm is the
reference implementation of this piece (third-party; scikit-learn 1.9.0 in this image and on the GPU box).

Bar: the GPU solver restates libsvm's SMO arithmetic (float-cached Q rows, tie rules, unfused updates), so the
iterates coincide: same iteration count and support vectors, alpha and decision values to 1e-9 relative.  Should the
paths ever part (they stop within eps = 1e-3 of the same optimum), the looser tier still holds: AUC within 2e-3."""
import numpy as np
import pytest

from conftest import random_seqs
from gkmqc_b200 import capi, driver

pytestmark = pytest.mark.gpu

SVC = pytest.importorskip("sklearn.svm").SVC
from sklearn.metrics import roc_auc_score
from sklearn.model_selection import StratifiedKFold


@pytest.fixture(scope="module")
def problem():
    """two classes that differ by planted motifs, so that the SVM has something to learn"""
    capi.load()
    rng = np.random.default_rng(5)
    npos = nneg = 300
    seqs = random_seqs(npos + nneg, 200, seed=21)
    motifs = ["GATAAGGCAT", "TTGACGTCAA", "CCCGCCCCTA"]
    for i in range(npos):
        s = list(seqs[i])
        for m in motifs[: 1 + i % 3]:
            p = int(rng.integers(0, 190))
            mm = list(m)
            if rng.random() < 0.5:
                mm[int(rng.integers(0, 10))] = "ACGT"[int(rng.integers(0, 4))]
            s[p:p + 10] = mm
        seqs[i] = "".join(s)
    P = capi.Problem(4, 10, 6, 3, 50, 50.0, 1.0)
    P.add_many(seqs)
    K = P.kernel_lower()
    K = np.maximum(K, K.T)
    y = np.concatenate((np.ones(npos, int), np.zeros(nneg, int)))
    yield P, K, y
    P.close()


def sklearn_fit(K, y, train, test, C, eps):
    sv = SVC(kernel="precomputed", C=C, tol=eps, shrinking=False, gamma=1.0, cache_size=500)
    sv.fit(K[train][:, train], y[train])
    return sv, sv.decision_function(K[test][:, train])


@pytest.mark.parametrize("C,eps", [(1.0, 1e-3), (0.1, 1e-3), (10.0, 1e-4)])
def test_fits_follow_libsvm(problem, C, eps):
    P, K, y = problem
    kf = StratifiedKFold(n_splits=5, shuffle=True, random_state=3)
    splits = list(kf.split(np.zeros(len(y)), y))
    scores, fits, alphas = capi.svm_cv(y, splits, kmat=K, C=C, eps=eps)
    for (train, test), s, f, a in zip(splits, scores, fits, alphas):
        sv, ref = sklearn_fit(K, y, train, test, C, eps)
        # loose tier first: whatever the path, the optimum is the same within the stopping tolerance
        assert abs(roc_auc_score(y[test], s) - roc_auc_score(y[test], ref)) < 2e-3
        np.testing.assert_allclose(s, ref, atol=5 * eps * max(1.0, C))
        # strict tier: same path
        assert f["n_sv"] == len(sv.support_)
        order = np.concatenate((train[y[train] == 0], train[y[train] == 1]))   # libsvm's grouping
        sv_ids = order[a > 0]
        assert np.array_equal(np.sort(sv_ids), np.sort(train[sv.support_]))
        # dual_coef_ = alpha * y' with sklearn's sign flip: negative for label 0
        coef = np.where(y[order] == 0, -a, a)[a > 0]
        by_id = dict(zip(train[sv.support_], sv.dual_coef_[0]))
        np.testing.assert_allclose(coef, [by_id[i] for i in sv_ids], rtol=1e-9, atol=1e-12)
        assert f["rho"] == pytest.approx(sv.intercept_[0], rel=1e-9, abs=1e-12)
        np.testing.assert_allclose(s, ref, rtol=1e-9, atol=1e-11)
        assert f["nu"] == pytest.approx(np.sum(np.abs(sv.dual_coef_[0])) / len(train), rel=1e-9)


def test_resident_matrix_equals_host_matrix(problem):
    """kmat = NULL: the kernel matrix is computed on the device, mirrored there and consumed there"""
    P, K, y = problem
    kf = StratifiedKFold(n_splits=4, shuffle=True, random_state=11)
    splits = list(kf.split(np.zeros(len(y)), y))
    s_host, f_host, _ = capi.svm_cv(y, splits, kmat=K, C=1.0, eps=1e-3)
    s_dev, f_dev, _ = capi.svm_cv(y, splits, problem=P, C=1.0, eps=1e-3)
    for a, b in zip(s_host, s_dev):
        assert np.array_equal(a, b)
    assert [f["n_iter"] for f in f_host] == [f["n_iter"] for f in f_dev]


def test_driver_crossValidate_matches_the_reference_flow(problem):
    """gkmsvm.crossValidate (gkmsvm.py:126-176) restated with sklearn fits, against the driver mirror"""
    P, K, y = problem
    npos = int(y.sum())
    args_svm = [1.0, 0.001, 0, 100, 5, 2, 0, 7, 4]
    mean_gpu, std_gpu = driver.crossValidate(args_svm, K, npos, len(y) - npos)
    aucs = []
    for _ in range(2):
        kf = StratifiedKFold(n_splits=5, shuffle=True, random_state=7)
        for train, test in kf.split(np.zeros(len(y)), y):
            _, ref = sklearn_fit(K, y, train, test, 1.0, 0.001)
            aucs.append(roc_auc_score(y[test], ref))
    assert mean_gpu == pytest.approx(np.mean(aucs), abs=1e-9)
    assert std_gpu == pytest.approx(np.std(aucs), abs=1e-9)
    assert 0.6 < mean_gpu <= 1.0


def test_bad_arguments(problem):
    P, K, y = problem
    with pytest.raises(ValueError):
        capi.svm_cv(y, [(np.arange(10), np.arange(10, 20))], kmat=K)   # one class only
    with pytest.raises(capi.GkmError):
        capi.svm_cv(y, [(np.arange(len(y)), np.arange(5))], kmat=K, C=-1.0)


def test_init_mirrors_one_bin_of_evaluate(tmp_path):
    """driver.init = gkmsvm.init (gkmsvm.py:182-222): kernel matrix + 2 x 3 cross-validation + the line in <name>.gkmqc.eval.out;
    the AUCs are those of sklearn's SVC on the same splits (what the reference's pool computes), host matrix or resident"""
    import types
    from sklearn.metrics import roc_auc_score
    from sklearn.model_selection import StratifiedKFold
    from sklearn.svm import SVC
    from gkmqc_b200 import driver
    rng = np.random.default_rng(11)
    n, npos = 240, 120
    acgt = np.frombuffer(b"ACGT", np.uint8)
    arr = acgt[rng.integers(0, 4, (n, 200))]
    for i in range(npos):   # a motif the positives share
        at = int(rng.integers(0, 190))
        arr[i, at:at + 10] = np.frombuffer(b"GATAAGGCAT", np.uint8)
    pos, neg = tmp_path / "p.fa", tmp_path / "n.fa"
    pos.write_text("".join(">p%d\n%s\n" % (i, arr[i].tobytes().decode()) for i in range(npos)))
    neg.write_text("".join(">n%d\n%s\n" % (i, arr[i].tobytes().decode()) for i in range(npos, n)))
    args = types.SimpleNamespace(kernel_type=4, full_word_length=10, non_gap_length=6, max_num_gaps=3, init_decay=50, half_life_decay=50.0,
                                 rbf_gamma=1.0, n_processes=1, verbosity=0, regularization=1.0, precision=1e-3, shrinking=0, cache_size=100,
                                 ncv=3, repeats=2, fast_estimation=0, random_seeds=5, name=str(tmp_path / "bin0"))
    auc, std = driver.init(str(pos), str(neg), args)
    auc_r, std_r = driver.init(str(pos), str(neg), args, resident=True)
    assert abs(auc - auc_r) < 1e-12 and abs(std - std_r) < 1e-12
    lines = open(args.name + ".gkmqc.eval.out").read().splitlines()
    assert len(lines) == 2 and lines[0].split("\t")[:3] == [str(pos), str(neg), str(npos)]
    assert abs(float(lines[0].split("\t")[3]) - auc) < 1e-15
    # the reference's own flow on the same matrix and splits (gkmsvm.py:104-160)
    kmat, _, _ = driver.computeGkmKernel([4, 10, 6, 3, 50, 50.0, 1.0, str(pos), str(neg), 1, 0], max_seqs=n)
    y = np.concatenate((np.repeat(1, npos), np.repeat(0, n - npos)))
    aucs = []
    for _ in range(2):
        for tr, te in StratifiedKFold(n_splits=3, shuffle=True, random_state=5).split(np.zeros(n), y):
            sv = SVC(kernel="precomputed", C=1.0, tol=1e-3, shrinking=False, gamma=1.0, cache_size=100)
            aucs.append(roc_auc_score(y[te], sv.fit(kmat[tr][:, tr], y[tr]).decision_function(kmat[te][:, tr])))
    assert abs(np.mean(aucs) - auc) < 1e-9 and abs(np.std(aucs) - std) < 1e-9 and auc > 0.7


def test_command_line_of_gkmsvm_on_the_engine(tmp_path):
    """python -m gkmqc_b200.driver: the option set and defaults of scripts/gkmsvm.py (gkmsvm.py:236-320) -- wgkm L=10 k=6 d=3 M=50
    H=50, C=1, eps=1e-3, 5-fold x 1 repeat -- and its output line"""
    from gkmqc_b200 import driver
    rng = np.random.default_rng(2)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    arr = acgt[rng.integers(0, 4, (160, 150))]
    for i in range(80):
        at = int(rng.integers(0, 140))
        arr[i, at:at + 10] = np.frombuffer(b"TTGACGTCAA", np.uint8)
    pos, neg = tmp_path / "p.fa", tmp_path / "n.fa"
    pos.write_text("".join(">p%d\n%s\n" % (i, arr[i].tobytes().decode()) for i in range(80)))
    neg.write_text("".join(">n%d\n%s\n" % (i, arr[i].tobytes().decode()) for i in range(80, 160)))
    name = str(tmp_path / "run")
    auc, std = driver.main(["-p", str(pos), "-n", str(neg), "-w", name, "-s", "3", "-v", "0"])
    auc2, std2 = driver.main(["-p", str(pos), "-n", str(neg), "-w", name, "-s", "3", "-v", "0", "--resident"])
    assert auc == auc2 and std == std2 and 0.5 < auc <= 1.0
    lines = [l.split("\t") for l in open(name + ".gkmqc.eval.out").read().splitlines()]
    assert len(lines) == 2 and lines[0][2] == "80" and float(lines[1][3]) == auc
    # the defaults are gkmQC's: the same call spelled out
    auc3, _ = driver.main(["-p", str(pos), "-n", str(neg), "-w", name, "-s", "3", "-v", "0", "-t", "4", "-L", "10", "-k", "6", "-d", "3",
                           "-M", "50", "-H", "50", "-C", "1.0", "-e", "0.001", "-x", "5", "-r", "1"])
    assert auc3 == auc


def _resident_against_host_matrix(ids):
    """the matrix kept on the device (gkmb200_resident_rows) = the caller-side matrix of the same problem, symmetrised like
    gkmsvm.py:96-97 -- on the GPUs `ids` of this process: with several, every GPU stores its chunks into the first one's
    memory (peer access), so the comparison covers rows written by each of them"""
    lib = capi.load()
    arr = (capi.ctypes.c_int * len(ids))(*ids)
    assert lib.gkmb200_set_devices(arr, len(ids)) == 0, capi.last_error()
    seqs = random_seqs(1300, 260, seed=41, ragged=True)
    try:
        for variant, kt, L, k, d in (("diag", 2, 11, 7, 3), ("index", 2, 11, 7, 3), ("index", 4, 10, 6, 3), ("mma", 4, 10, 6, 3)):
            capi.set_option("kernel", variant)
            with capi.Problem(kt, L, k, d, 50, 50.0, 1.0) as P:
                P.add_many(seqs)
                R = P.resident_matrix()
                st = P.stats()
                assert st["devices"] == len(ids), (variant, st)
                low = np.tril(P.kernel_lower(), -1)
                K = low + low.T + np.eye(P.n)
                assert np.array_equal(R, K), (variant, kt, len(ids))
                assert np.array_equal(P.resident_matrix(5, 7), K[5:12])          # a band of the matrix that is already there
                with pytest.raises(capi.GkmError):
                    P.resident_matrix(P.n - 3, 4)
            if variant == "index":
                assert st["kernel_variant"] == 4
    finally:
        capi.set_option("kernel", "auto")


def test_resident_matrix_read_back():
    _resident_against_host_matrix([0])


def test_resident_matrix_shared_by_the_gpus_of_one_process_if_available(problem):
    ndev = capi.device_count()
    if ndev < 2:
        pytest.skip("single GPU box")
    try:
        _resident_against_host_matrix(list(range(ndev)))
        # the consumer on a matrix every GPU wrote a share of: the same fits, bit for bit, as on the host matrix
        P0, K, y = problem
        splits = list(StratifiedKFold(n_splits=4, shuffle=True, random_state=11).split(np.zeros(len(y)), y))
        s_host, f_host, _ = capi.svm_cv(y, splits, kmat=K, C=1.0, eps=1e-3)
        with capi.Problem(4, 10, 6, 3, 50, 50.0, 1.0) as P:
            for i in range(P0.n):
                f, _ = P0.codes(i)
                P.add(bytes(b"ACGT"[c - 1] for c in f))
            s_dev, f_dev, _ = capi.svm_cv(y, splits, problem=P, C=1.0, eps=1e-3)
            assert P.stats()["devices"] == ndev
        for a, b in zip(s_host, s_dev):
            assert np.array_equal(a, b)
        assert [f["n_iter"] for f in f_host] == [f["n_iter"] for f in f_dev]
    finally:
        ids = (capi.ctypes.c_int * 1)(0)
        capi.load().gkmb200_set_devices(ids, 1)
