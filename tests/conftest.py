import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "*_t*.npz")))


def load_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fx = name.split("_")[0]
    cfg = dict(kernel_type=int(g["kernel_type"]), L=int(g["L"]), k=int(g["k"]), d=int(g["d"]),
               M=int(g["M"]), H=float(g["H"]), gamma=float(g["gamma"]))
    return g, cfg, os.path.join(GOLD, fx + "_pos.fa"), os.path.join(GOLD, fx + "_neg.fa")


def random_seqs(n, length, seed, ragged=False):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(n):
        ln = int(rng.integers(max(20, length // 3), length + 1)) if ragged else length
        out.append(acgt[rng.integers(0, 4, ln)].tobytes().decode())
    return out


def write_fasta(path, seqs, prefix="s"):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n%s\n" % (prefix, i, s))


@pytest.fixture(scope="session")
def product_lib():
    from gkmqc_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def emu_lib():
    import ctypes
    import __graft_entry__ as ge
    from gkmqc_b200 import capi
    # GKM_EMU_LIB: another build of the emulator library (tools/asan_host.sh: the host C under ASan + UBSan)
    lib = ctypes.CDLL(os.environ.get("GKM_EMU_LIB") or ge.build_emulator())
    capi._declare(lib)
    lib.gkm_emu_hist_lower.argtypes = [ctypes.c_void_p, capi.c_i32_p]
    lib.gkm_emu_hist.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, capi.c_i32_p]
    return lib
