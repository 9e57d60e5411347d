"""CPU tier: positional weights of the wgkm kernel types (libgkm.c:910-932) on generated (M, H, number of L-mers):
the product's per-position routine, the distance table the DEVICE packer indexes (gkm_posweight_table: exp() never runs
on the GPU) and the oracle's restatement must give the same bytes -- the uint8 wrap of M = 255 at the centre included."""
import ctypes

import numpy as np
import pytest

import pyoracle
from gkmqc_b200 import capi

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

# derandomize: every run draws the same fixed sequence of examples (a test tier must not be a lottery); the wider sweeps
# named in DESIGN.md were run once by raising max_examples.

MAX_BASES = 2047


def table(kernel_type, M, H):
    lib = capi.load()
    lib.gkm_posweight_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, capi.c_u8_p]
    lib.gkm_posweight_table.restype = None
    tab = np.zeros(MAX_BASES + 1, np.uint8)
    lib.gkm_posweight_table(kernel_type, M, H, tab.ctypes.data_as(capi.c_u8_p))
    return tab


@settings(derandomize=True, max_examples=150, deadline=None)
@given(M=st.integers(1, 255), H=st.one_of(st.floats(0.3, 3000.0, allow_nan=False), st.sampled_from([1.0, 50.0, 20.0, 0.5])),
       nk=st.integers(1, MAX_BASES - 1), kernel_type=st.sampled_from([4, 5]))
def test_three_routes_to_the_same_bytes(M, H, nk, kernel_type):
    wt, wt_rc = capi.posweights(nk, kernel_type, M, H)
    tab = table(kernel_type, M, H)
    dist = np.abs(nk // 2 - np.arange(nk))
    assert np.array_equal(wt, tab[dist]), "what the device packer looks up = what the host computes per position"
    assert np.array_equal(wt_rc, wt[::-1])
    assert wt.max() <= M
    L = 2
    o = pyoracle.Oracle(kernel_type, L, 1, 1, M, H)
    o.add("A" * (nk + L - 1))
    oa, ob = o.poswt(0)
    o.close()
    assert np.array_equal(wt, oa) and np.array_equal(wt_rc, ob)


def test_unit_weights_for_the_other_kernel_types():
    for kt in (0, 1, 2, 3):
        assert np.all(table(kt, 50, 50.0) == 1)
        a, b = capi.posweights(77, kt, 200, 3.0)
        assert np.all(a == 1) and np.all(b == 1)


def test_wrap_of_the_centre_weight_at_M_255():
    """floor(255 * 1 + 1) = 256 does not fit the reference's u_int8_t: the centre L-mer weighs 0 (golden set mix_t4_L10k6d3_M255)"""
    a, _ = capi.posweights(9, 4, 255, 20.0)
    assert a[4] == 0 and a[3] == a[5] and a[3] > 200
