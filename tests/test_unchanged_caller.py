"""The reference's own caller, unchanged: /root/reference/scripts/gkmsvm.py imported as it is, with its module
global `bin_dir` (gkmsvm.py:38, read at call time :85) pointed at gkmqc_b200/bin.  Runs only where the reference
tree exists (this container); the GPU box has no /root/reference, there gkmqc_b200/driver.py -- the line-for-line
mirror of that caller -- carries the same checks (tests/test_gpu_parity.py::test_driver_mirror_returns_what_gkmsvm_would).
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import load_golden
from gkmqc_b200 import capi

REF_SCRIPT = "/root/reference/scripts/gkmsvm.py"


def load_reference_caller():
    if not os.path.exists(REF_SCRIPT):
        pytest.skip("no reference tree on this box")
    spec = importlib.util.spec_from_file_location("gkmsvm_unchanged", REF_SCRIPT)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gkmsvm_unchanged"] = mod
    spec.loader.exec_module(mod)
    mod.bin_dir = capi.BIN_DIR
    return mod


def test_unchanged_gkmsvm_binds_to_the_library_and_fails_loudly_without_a_gpu():
    """no GPU here: the call must reach gkm_main_pywrapper of OUR library through gkmsvm.py's own ctypes code and come
    back as the error return that gkmsvm.py turns into sys.exit() (gkmsvm.py:90-92) -- never a CPU result"""
    capi.load()
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible: the GPU test below covers this box")
    mod = load_reference_caller()
    g, cfg, pos, neg = load_golden("uni_t4_L10k6d3")
    with pytest.raises(SystemExit):
        mod.computeGkmKernel([4, 10, 6, 3, 50, 50.0, 1.0, pos, neg, 1, 0])
    assert "no sm_100" in capi.last_error()


@pytest.mark.gpu
def test_unchanged_gkmsvm_end_to_end_with_fork_after_cuda():
    """computeGkmKernel + crossValidate of the unchanged script: the kernel matrix it symmetrises equals the
    reference's, and its multiprocessing.Pool forks AFTER the CUDA context is alive in the parent (gkmsvm.py:152-155;
    the children only run sklearn)"""
    mod = load_reference_caller()
    if capi.device_count() < 1:
        pytest.fail("no B200 visible; the product has no CPU fallback")
    g, cfg, pos, neg = load_golden("uni_t4_L10k6d3")
    kmat, npos, nneg = mod.computeGkmKernel([4, 10, 6, 3, 50, 50.0, 1.0, pos, neg, 1, 0])
    n = len(g["lens"])
    assert (npos, nneg) == (int(g["npos"]), n - int(g["npos"])) and kmat.shape == (n, n)
    assert np.array_equal(kmat, np.maximum(g["kmat"], g["kmat"].T))
    # regularization, precision, shrinking, cache_size, ncv, repeats, fast_estimation, random_seeds, processes
    auc, std = mod.crossValidate([1.0, 1e-3, 0, 100, 2, 2, 0, 1, 2], kmat, npos, nneg)
    assert 0.0 <= auc <= 1.0 and np.isfinite(std)
    # the parent's CUDA context is still usable after the pool's fork + join
    kmat2, _, _ = mod.computeGkmKernel([4, 10, 6, 3, 50, 50.0, 1.0, pos, neg, 1, 0])
    assert np.array_equal(kmat2, kmat)


def test_command_line_mirror_has_the_reference_options_and_defaults():
    """gkmqc_b200.driver.main mirrors the argparse of scripts/gkmsvm.py (gkmsvm.py:236-320): every option of the reference,
    with its short flag, type and default, read out of the reference's source text"""
    import argparse
    import re
    if not os.path.exists(REF_SCRIPT):
        pytest.skip("no reference tree on this box")
    text = open(REF_SCRIPT).read()
    ref = {}
    for m in re.finditer(r'add_argument\("(-[\w@])",\s*"--([\w-]+)",\s*type=(\w+),\s*(?:required=True|default=([^,]+)),', text):
        ref[m.group(2)] = (m.group(1), m.group(3), m.group(4))
    assert len(ref) >= 20 and "full-word-length" in ref and "pos-fa" in ref
    from gkmqc_b200 import driver
    captured = {}
    real = argparse.ArgumentParser.parse_args

    def spy(self, argv=None):
        captured["parser"] = self
        raise SystemExit(0)
    argparse.ArgumentParser.parse_args = spy
    try:
        with pytest.raises(SystemExit):
            driver.main([])
    finally:
        argparse.ArgumentParser.parse_args = real
    ours = {}
    for a in captured["parser"]._actions:
        longs = [o[2:] for o in a.option_strings if o.startswith("--")]
        shorts = [o for o in a.option_strings if not o.startswith("--")]
        if longs and shorts:
            ours[longs[0]] = (shorts[0], a.type.__name__ if a.type else None, a.default, a.required)
    for name, (short, typ, default) in ref.items():
        assert name in ours, "option --%s of the reference is missing" % name
        o_short, o_type, o_default, o_req = ours[name]
        assert o_short == short and o_type == typ, name
        if default is None:
            assert o_req, name
        else:
            assert o_default == eval(default), (name, o_default, default)


@pytest.mark.timeout(300)
def test_unchanged_gkmsvm_sees_no_difference_between_the_two_libraries(tmp_path):
    """CPU tier: scripts/gkmsvm.py, unchanged, run twice -- once on the reference's own gkmkern_pylib.so (oracle/_ref), once on
    the product's HOST code (gkm_capi.c: option handling, FASTA reading, the writes into the caller's rows) over the host
    stand-in of the device layer (test infrastructure, tests/emu/dev_stub.cc; the shipped library has no CPU path).
    computeGkmKernel must return the same matrix and sizes, crossValidate the same AUC."""
    import shutil
    import pyoracle
    import __graft_entry__ as ge
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref (the compiled reference) is not here")
    mod = load_reference_caller()
    emu_dir = tmp_path / "bin"
    emu_dir.mkdir()
    shutil.copy(ge.build_abi_emulator(), emu_dir / "gkmkern_pylib.so")     # the artefact name gkmsvm.py:85 loads
    pos, neg = tmp_path / "pos.fa", tmp_path / "neg.fa"
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    arr = acgt[rng.integers(0, 4, (40, 120))]
    for i in range(20):     # a motif the positives share, so that the cross-validation has something to find
        at = int(rng.integers(0, 110))
        arr[i, at:at + 10] = np.frombuffer(b"GATAAGGCAT", np.uint8)
    pos.write_text("".join(">p%d desc\n%s\n" % (i, arr[i].tobytes().decode()) for i in range(20)))
    neg.write_text("".join(">n%d\r\n%s\r\n" % (i, arr[i].tobytes().decode().lower()) for i in range(20, 40)))
    results = []
    try:
        for bin_dir in (pyoracle.REF_DIR, str(emu_dir)):
            mod.bin_dir = bin_dir
            kmat, npos, nneg = mod.computeGkmKernel([4, 8, 5, 3, 50, 50.0, 1.0, str(pos), str(neg), 2, 0])
            auc, std = mod.crossValidate([1.0, 1e-3, 0, 100, 4, 2, 0, 3, 2], kmat, npos, nneg)
            results.append((kmat, npos, nneg, auc, std))
    finally:
        mod.bin_dir = capi.BIN_DIR
    (k0, p0, n0, a0, s0), (k1, p1, n1, a1, s1) = results
    assert (p0, n0) == (p1, n1) == (20, 20) and k0.shape == k1.shape == (40, 40)
    assert np.array_equal(k0, k1)
    assert a0 == a1 and s0 == s1 and a0 > 0.6
    # gkmqc_b200/driver.py, the mirror the GPU box uses in place of this script: the same triple from the same library
    import ctypes
    from gkmqc_b200 import driver
    saved = capi._lib
    capi._lib = ctypes.CDLL(str(emu_dir / "gkmkern_pylib.so"))
    capi._declare(capi._lib)
    try:
        k2, p2, n2 = driver.computeGkmKernel([4, 8, 5, 3, 50, 50.0, 1.0, str(pos), str(neg), 2, 0], max_seqs=64)
    finally:
        capi._lib = saved
    assert (p2, n2) == (20, 20) and np.array_equal(k2, k0)
