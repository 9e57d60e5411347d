"""CPU tier: the operator itself -- gkm_main_pywrapper(gkmOpt*, double **kmat, int *kmat_size), gkmkern_pylib.c:92-246 --
of the product's host code (gkm_capi.c over the host stand-in of the device layer, tests/test_abi_on_emulator.py) against
the unmodified reference's gkmkern_pylib.so, called the way scripts/gkmsvm.py:75-88 calls it, on generated options and
FASTA files.  What a caller can observe must be the same: the return code (the parameter gate of
gkmkern_pylib.c:38-64 included), kmat_size, and EVERY byte of the caller's matrix -- kernel values in the strict lower
triangle, 1.0 on the diagonal of the rows that exist, nothing written anywhere else (the matrix is pre-filled with a
sentinel instead of zeros to see that)."""
import ctypes
import os

import numpy as np
import pytest

import pyoracle
from gkmqc_b200 import capi
from test_fasta_fuzz import fasta_text

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

# derandomize: every run draws the same fixed sequence of examples (a test tier must not be a lottery); the wider sweeps
# named in DESIGN.md were run once by raising max_examples.

pytestmark = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref (the compiled reference) is not here")

NMAX = 16      # rows and columns of the caller's matrix: more than any generated problem holds
SENTINEL = -7.25


@pytest.fixture(scope="module")
def libs():
    import __graft_entry__ as ge
    ours = ctypes.CDLL(os.environ.get("GKM_ABI_EMU_LIB") or ge.build_abi_emulator())
    capi._declare(ours)
    ref = pyoracle.ref_pywrapper()
    ref.clog_free.argtypes = [ctypes.c_int]
    ref.clog_free.restype = None
    return ours, ref


def call(lib, pos, neg, opt, nthreads):
    kmat = np.full((NMAX, NMAX), SENTINEL)
    ret, kmat, npos, nneg = pyoracle.call_pywrapper(lib, pos, neg, nthreads=nthreads, verbosity=0, kmat=kmat, **opt)
    return ret, kmat.tobytes(), (npos, nneg)


@st.composite
def options(draw):
    kind = draw(st.integers(0, 9))
    if kind == 0:      # something the gate refuses (gkmkern_pylib.c:38-64)
        L = draw(st.sampled_from([0, 1, 5, 8, 13, 14, 40]))
        return dict(kernel_type=draw(st.sampled_from([-1, 2, 4, 6, 99])), L=L, k=draw(st.integers(0, 15)), d=draw(st.integers(0, 9)),
                    M=50, H=50.0, gamma=1.0)
    L = draw(st.integers(2, 8))
    k = draw(st.integers(1, L))
    d = draw(st.integers(0, min(L - k, 4)))
    return dict(kernel_type=draw(st.integers(0, 5)), L=L, k=k, d=d, M=draw(st.sampled_from([1, 50, 63, 64, 255])),
                H=draw(st.sampled_from([0.7, 50.0, 300.0])), gamma=draw(st.sampled_from([0.5, 1.0, 3.0])))


def gate(o):
    """the reference's own parameter gate, restated (gkmkern_pylib.c:38-64)"""
    return 0 <= o["kernel_type"] <= 5 and 2 <= o["L"] <= 12 and o["k"] <= o["L"] and o["d"] <= o["L"] - o["k"]


@settings(derandomize=True, max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(opt=options(), pos=fasta_text(min_len=8), neg=fasta_text(min_len=8), nthreads=st.integers(1, 3))
def test_same_return_code_matrix_and_sizes(libs, tmp_path, opt, pos, neg, nthreads):
    ours, ref = libs
    pf, nf = tmp_path / "pos.fa", tmp_path / "neg.fa"
    pf.write_bytes(pos.encode("ascii"))
    nf.write_bytes(neg.encode("ascii"))
    ok = gate(opt)
    if ok and (opt["k"] < 1 or opt["d"] < 0):
        return                                           # the gate lets these through and the reference then divides by nothing: not a case
    a = call(ours, str(pf), str(nf), opt, nthreads)
    b = call(ref, str(pf), str(nf), opt, nthreads)
    if b[0] != 0:
        ref.clog_free(0)                                 # the reference leaves its logger open on the error return (gkmkern_pylib.c:157-161)
    assert a[0] == b[0] == (0 if ok else 1), (opt, a[0], b[0])
    assert a[2] == b[2], (opt, a[2], b[2])
    assert a[1] == b[1], (opt, nthreads, np.frombuffer(a[1]).reshape(NMAX, NMAX), np.frombuffer(b[1]).reshape(NMAX, NMAX))
    if not ok:
        assert np.all(np.frombuffer(a[1]) == SENTINEL)
