"""CPU tier: the reference-ABI shim (gkmqc_b200/csrc/gkm_capi.c: gkmkernel_init, read_problems, new_object, build_tree,
kernelfunc_batch[_all], swap_index, update_index, delete_object -- libgkm.h:132-147) over a host stand-in of the device
layer (tests/emu/dev_stub.cc, build/libgkm_abi_emu.so).  The test BODIES are those of the GPU tier
(tests/test_gpu_abi.py), called here with the emulator library: every field of every object against the reference's
formulas, sqnorm and kernel rows bit-identical to the reference-generated golden values, kernelfunc_batch against the
oracle, objects longer than the engine holds.  tools/asan_host.sh runs this file under ASan + UBSan.

The stand-in is test infrastructure: the product library has no CPU path (test_host_logic.py::
test_compute_fails_loudly_without_gpu)."""
import ctypes
import os

import numpy as np
import pytest

import test_gpu_abi as T
from conftest import load_golden
from gkmqc_b200 import capi


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    lib = ctypes.CDLL(os.environ.get("GKM_ABI_EMU_LIB") or ge.build_abi_emulator())
    capi._declare(lib)
    assert ctypes.sizeof(T.gkm_data) == 88
    P, G = ctypes.c_void_p, ctypes.POINTER(T.gkm_data)
    lib.gkmkernel_init.restype = P
    lib.gkmkernel_init.argtypes = [ctypes.POINTER(capi.gkm_parameter)]
    lib.gkmkernel_read_problems.argtypes = [P, ctypes.POINTER(T.svm_problem), ctypes.c_char_p, ctypes.c_char_p]
    lib.gkmkernel_build_tree.argtypes = [P, ctypes.POINTER(G), ctypes.c_int]
    lib.gkmkernel_kernelfunc_batch_all.restype = capi.c_dbl_p
    lib.gkmkernel_kernelfunc_batch_all.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, capi.c_dbl_p]
    lib.gkmkernel_kernelfunc_batch.restype = capi.c_dbl_p
    lib.gkmkernel_kernelfunc_batch.argtypes = [P, ctypes.c_int, ctypes.POINTER(G), ctypes.c_int, capi.c_dbl_p]
    lib.gkmkernel_delete_object.argtypes = [G]
    lib.gkmkernel_free_object.argtypes = [G]
    lib.gkmkernel_destroy.argtypes = [P]
    lib.gkmkernel_new_object.restype = G
    lib.gkmkernel_new_object.argtypes = [P, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    lib.gkmkernel_swap_index.argtypes = [P, ctypes.c_int, ctypes.c_int]
    lib.gkmkernel_update_index.argtypes = [P]
    # the bodies ask capi.load() for the last error: that is this library for the duration of the module
    saved = capi._lib
    capi._lib = lib
    lib.gkmb200_set_verbosity(0)
    yield lib
    capi._lib = saved


@pytest.mark.parametrize("name", ["mix_t4_L11k7d3", "uni_t2_L11k7d3"])
def test_read_problems_fills_what_the_reference_fills(name, lib):
    T.test_read_problems_fills_what_the_reference_fills(name, lib)


@pytest.mark.parametrize("kernel_type", [2, 4, 5])
def test_kernelfunc_batch_against_oracle(kernel_type, lib, tmp_path):
    T.test_kernelfunc_batch_against_oracle(kernel_type, lib, tmp_path)


def test_new_object_cuts_what_the_engine_cannot_hold(lib):
    T.test_new_object_cuts_what_the_engine_cannot_hold(lib)


def test_index_permutation_bookkeeping(lib):
    """gkmkernel_swap_index / gkmkernel_update_index (libgkm.c:1071-1109): after the update id i means what libsvm calls i,
    and rows come back in the new order; a freed object (gkmkernel_free_object keeps the struct) is refused by name"""
    g, cfg, pos, neg = load_golden("uni_t2_L11k7d3")
    param = capi.make_param(**cfg)
    kern = lib.gkmkernel_init(ctypes.byref(param))
    prob = T.svm_problem()
    assert lib.gkmkernel_read_problems(kern, ctypes.byref(prob), os.fsencode(pos), os.fsencode(neg)) == int(g["npos"])
    n = prob.l
    lib.gkmkernel_build_tree(kern, prob.x, n)
    K = np.tril(g["kmat"], -1)
    K = K + K.T + np.eye(n)
    perm = list(range(n))
    for i, j in ((0, n - 1), (3, 7), (7, 2), (5, 5)):
        lib.gkmkernel_swap_index(kern, i, j)
        perm[i], perm[j] = perm[j], perm[i]
    lib.gkmkernel_swap_index(kern, -1, 2)        # out of range: ignored
    lib.gkmkernel_swap_index(kern, 1, n)
    lib.gkmkernel_update_index(kern)
    res = np.zeros(n)
    for a in (0, 2, 7, n - 1):
        lib.gkmkernel_kernelfunc_batch_all(kern, a, 0, n, res.ctypes.data_as(capi.c_dbl_p))
        want = K[perm[a]][perm]
        want[a] = res[a]                         # the diagonal of a full row is computed, not the constant 1
        assert np.array_equal(res, want), a
        assert abs(res[a] - 1.0) < 1e-12
    # start == end leaves res alone but for nothing; end < start returns without touching it
    res[:] = -3.0
    lib.gkmkernel_kernelfunc_batch_all(kern, 1, 4, 4, res.ctypes.data_as(capi.c_dbl_p))
    lib.gkmkernel_kernelfunc_batch_all(kern, 1, 5, 4, res.ctypes.data_as(capi.c_dbl_p))
    assert np.all(res == -3.0)
    # an object whose arrays were released but whose struct lives on: the next image build names it instead of reading freed memory
    lib.gkmkernel_free_object(prob.x[4])
    lib.gkmkernel_build_tree(kern, prob.x, n)
    res[:] = -3.0
    lib.gkmkernel_kernelfunc_batch_all(kern, 1, 0, 3, res.ctypes.data_as(capi.c_dbl_p))
    assert not res[:3].any() and b"object 4" in lib.gkmb200_last_error()
    for i in range(n):
        lib.gkmkernel_delete_object(prob.x[i])
    lib.gkmkernel_destroy(kern)
