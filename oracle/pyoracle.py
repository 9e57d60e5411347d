"""ctypes access to the parity oracle (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -- oracle/gkm_oracle.c, the plain-C restatement (always available).
* ``RefHook`` -- oracle/_ref/gkmref_hook*.so, the UNMODIFIED reference compiled from
  /root/reference/src with a probe that exposes its integer mismatch profile
  (present only where oracle/Makefile target ``ref`` was built).
* ``ref_pywrapper`` -- the stock reference ``gkm_main_pywrapper`` called exactly like
  scripts/gkmsvm.py:75-88 does.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libgkm_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

c_int_p = ctypes.POINTER(ctypes.c_int)
c_dbl_p = ctypes.POINTER(ctypes.c_double)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "gkm_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref(ref_root="/root/reference"):
    """compile the reference where it lies; no-op (False) when it is not present"""
    if not os.path.isdir(os.path.join(ref_root, "src")):
        return have_ref()
    subprocess.check_call(["make", "-C", HERE, "ref", "REF=" + ref_root], stdout=subprocess.DEVNULL)
    return True


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("gkmkern_pylib.so", "gkmref_hook.so"))


class gkmOpt(ctypes.Structure):
    """mirror of struct _gkmOpt, libgkm.h:149-161 (same as scripts/gkmsvm.py:48-61)"""
    _fields_ = (
        ("kernel_type", ctypes.c_int), ("L", ctypes.c_int), ("k", ctypes.c_int), ("d", ctypes.c_int),
        ("M", ctypes.c_uint8), ("H", ctypes.c_double), ("gamma", ctypes.c_double),
        ("posfile", ctypes.c_char_p), ("negfile", ctypes.c_char_p),
        ("nthreads", ctypes.c_int), ("verbosity", ctypes.c_int),
    )


def call_pywrapper(lib, posfile, negfile, kernel_type=2, L=11, k=7, d=3, M=50, H=50.0, gamma=1.0,
                   nthreads=1, verbosity=0, nmax=None, kmat=None):
    """invoke a ``gkm_main_pywrapper`` (reference or product) the way gkmsvm.computeGkmKernel does.
    Returns (ret, kmat[nmax,nmax], npos, nneg); kmat holds the lower triangle + unit diagonal."""
    if kmat is None:
        kmat = np.zeros((nmax, nmax))
    rows = (kmat.ctypes.data + np.arange(kmat.shape[0]) * kmat.strides[0]).astype(np.uintp)
    narr = np.ones(2, dtype=np.int32)
    opts = gkmOpt(kernel_type, L, k, d, M, H, gamma, os.fsencode(posfile), os.fsencode(negfile), nthreads, verbosity)
    lib.gkm_main_pywrapper.restype = ctypes.c_int
    lib.gkm_main_pywrapper.argtypes = (ctypes.POINTER(gkmOpt), np.ctypeslib.ndpointer(dtype=np.uintp, ndim=1, flags="C"), c_int_p)
    ret = lib.gkm_main_pywrapper(ctypes.byref(opts), rows, narr.ctypes.data_as(c_int_p))
    return ret, kmat, int(narr[0]), int(narr[1])


def ref_pywrapper():
    return ctypes.CDLL(os.path.join(REF_DIR, "gkmkern_pylib.so"))


class Oracle:
    def __init__(self, kernel_type=2, L=11, k=7, d=3, M=50, H=50.0, gamma=1.0):
        lib = ctypes.CDLL(build_oracle())
        self.lib = lib
        lib.gkmo_set_new.restype = ctypes.c_void_p
        lib.gkmo_set_new.argtypes = [ctypes.c_int] * 5 + [ctypes.c_double] * 2
        lib.gkmo_set_free.argtypes = [ctypes.c_void_p]
        lib.gkmo_set_add.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        lib.gkmo_set_read_fasta.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        lib.gkmo_set_read_problem.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        lib.gkmo_set_size.argtypes = [ctypes.c_void_p]
        lib.gkmo_set_npos.argtypes = [ctypes.c_void_p]
        lib.gkmo_sqnorm.restype = ctypes.c_double
        lib.gkmo_sqnorm.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.gkmo_seqlen.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.gkmo_get_weights.argtypes = [ctypes.c_void_p, c_dbl_p]
        lib.gkmo_get_poswt.argtypes = [ctypes.c_void_p, ctypes.c_int, c_u8_p, c_u8_p]
        lib.gkmo_get_codes.argtypes = [ctypes.c_void_p, ctypes.c_int, c_u8_p, c_u8_p]
        lib.gkmo_hist.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_int_p]
        lib.gkmo_kernel.restype = ctypes.c_double
        lib.gkmo_kernel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.gkmo_matrix_lower.argtypes = [ctypes.c_void_p, c_dbl_p, ctypes.c_long, c_int_p]
        lib.gkmo_rect.argtypes = [ctypes.c_void_p, c_int_p, ctypes.c_int, ctypes.c_int, c_dbl_p, ctypes.c_long, c_int_p]
        lib.gkmo_weights.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p]
        self.L, self.d, self.kernel_type = L, d, kernel_type
        self.h = lib.gkmo_set_new(kernel_type, L, k, d, M, H, gamma)
        if not self.h:
            raise ValueError("bad oracle parameters")

    def close(self):
        if self.h:
            self.lib.gkmo_set_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add(self, seq, sid="s"):
        r = self.lib.gkmo_set_add(self.h, seq.encode("ascii"), sid.encode("ascii"))
        if r < 0:
            raise ValueError("sequence shorter than L")
        return r

    def read_problem(self, posfile, negfile):
        n = self.lib.gkmo_set_read_problem(self.h, os.fsencode(posfile), os.fsencode(negfile))
        if n < 0:
            raise IOError("oracle could not read problem")
        return n

    @property
    def n(self):
        return self.lib.gkmo_set_size(self.h)

    @property
    def npos(self):
        return self.lib.gkmo_set_npos(self.h)

    def weights(self):
        w = np.zeros(self.L + 1)
        self.lib.gkmo_get_weights(self.h, w.ctypes.data_as(c_dbl_p))
        return w

    def sqnorm(self):
        return np.array([self.lib.gkmo_sqnorm(self.h, i) for i in range(self.n)])

    def seqlen(self, i):
        return self.lib.gkmo_seqlen(self.h, i)

    def poswt(self, i):
        nk = self.seqlen(i) - self.L + 1
        a = np.zeros(nk, np.uint8)
        b = np.zeros(nk, np.uint8)
        self.lib.gkmo_get_poswt(self.h, i, a.ctypes.data_as(c_u8_p), b.ctypes.data_as(c_u8_p))
        return a, b

    def codes(self, i):
        n = self.seqlen(i)
        a = np.zeros(n, np.uint8)
        b = np.zeros(n, np.uint8)
        self.lib.gkmo_get_codes(self.h, i, a.ctypes.data_as(c_u8_p), b.ctypes.data_as(c_u8_p))
        return a, b

    def hist(self, a, b):
        h = np.zeros(self.d + 1, np.int32)
        self.lib.gkmo_hist(self.h, a, b, h.ctypes.data_as(c_int_p))
        return h

    def kernel(self, a, b):
        return self.lib.gkmo_kernel(self.h, a, b)

    def matrix_lower(self, with_hist=True):
        n = self.n
        K = np.zeros((n, n))
        H = np.zeros((n, n, self.d + 1), np.int32) if with_hist else None
        self.lib.gkmo_matrix_lower(self.h, K.ctypes.data_as(c_dbl_p), n,
                                   H.ctypes.data_as(c_int_p) if with_hist else None)
        return K, H

    def rect(self, rows, ncols, with_hist=True):
        rows = np.ascontiguousarray(rows, np.int32)
        K = np.zeros((len(rows), ncols))
        H = np.zeros((len(rows), ncols, self.d + 1), np.int32) if with_hist else None
        self.lib.gkmo_rect(self.h, rows.ctypes.data_as(c_int_p), len(rows), ncols, K.ctypes.data_as(c_dbl_p), ncols,
                           H.ctypes.data_as(c_int_p) if with_hist else None)
        return K, H


def oracle_weights(kernel_type, L, k):
    lib = ctypes.CDLL(build_oracle())
    lib.gkmo_weights.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p]
    w = np.zeros(L + 1)
    if lib.gkmo_weights(kernel_type, L, k, w.ctypes.data_as(c_dbl_p)):
        raise ValueError("bad parameters")
    return w


def ref_weights(kernel_type, L, k):
    """w[0..L] from the reference's own weight routines (no tree is allocated)"""
    lib = ctypes.CDLL(os.path.join(REF_DIR, "gkmref_hook_L16.so" if L > 12 else "gkmref_hook.so"))
    lib.gkmref_weights_only.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p, ctypes.c_int]
    w = np.zeros(L + 1)
    lib.gkmref_weights_only(kernel_type, L, k, w.ctypes.data_as(c_dbl_p), L + 1)
    return w


class RefHook:
    """one open problem inside the unmodified reference (process-global, like the reference itself)"""

    def __init__(self, posfile, negfile, kernel_type=2, L=11, k=7, d=3, M=50, H=50.0, gamma=1.0):
        name = "gkmref_hook_L16.so" if L > 12 else "gkmref_hook.so"
        lib = ctypes.CDLL(os.path.join(REF_DIR, name))
        self.lib = lib
        lib.gkmref_open.argtypes = [ctypes.c_int] * 5 + [ctypes.c_double] * 2 + [ctypes.c_char_p] * 2
        lib.gkmref_weights.argtypes = [c_dbl_p, ctypes.c_int]
        lib.gkmref_sqnorm.restype = ctypes.c_double
        lib.gkmref_sqnorm.argtypes = [ctypes.c_int]
        lib.gkmref_seqlen.argtypes = [ctypes.c_int]
        for f in ("gkmref_sid", "gkmref_seq_string"):   # absent from a hook built before these were added
            if hasattr(lib, f):
                getattr(lib, f).restype = ctypes.c_char_p
                getattr(lib, f).argtypes = [ctypes.c_int]
        lib.gkmref_poswt.argtypes = [ctypes.c_int, c_u8_p, c_u8_p]
        lib.gkmref_codes.argtypes = [ctypes.c_int, c_u8_p, c_u8_p]
        lib.gkmref_mmprofile.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p]
        lib.gkmref_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p]
        lib.gkmref_rows_timed.restype = ctypes.c_double
        lib.gkmref_rows_timed.argtypes = [c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p]
        lib.gkmref_rows_values.restype = ctypes.c_double
        lib.gkmref_rows_values.argtypes = [c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dbl_p, c_int_p, ctypes.c_long]
        self.L, self.d = L, d
        self.n = lib.gkmref_open(kernel_type, L, k, d, M, H, gamma, os.fsencode(posfile), os.fsencode(negfile))
        if self.n < 0:
            raise RuntimeError("reference probe: a problem is already open")
        self.npos = lib.gkmref_npos()

    def close(self):
        if self.lib is not None:
            self.lib.gkmref_close()
            self.lib = None

    def weights(self, n=None):
        n = self.L + 1 if n is None else n
        w = np.zeros(n)
        self.lib.gkmref_weights(w.ctypes.data_as(c_dbl_p), n)
        return w

    def sqnorm(self):
        return np.array([self.lib.gkmref_sqnorm(i) for i in range(self.n)])

    def seqlen(self, i):
        return self.lib.gkmref_seqlen(i)

    def sid(self, i):
        return self.lib.gkmref_sid(i)

    def seq_string(self, i):
        return self.lib.gkmref_seq_string(i)

    def poswt(self, i):
        nk = self.seqlen(i) - self.L + 1
        a = np.zeros(nk, np.uint8)
        b = np.zeros(nk, np.uint8)
        self.lib.gkmref_poswt(i, a.ctypes.data_as(c_u8_p), b.ctypes.data_as(c_u8_p))
        return a, b

    def codes(self, i):
        n = self.seqlen(i)
        a = np.zeros(n, np.uint8)
        b = np.zeros(n, np.uint8)
        self.lib.gkmref_codes(i, a.ctypes.data_as(c_u8_p), b.ctypes.data_as(c_u8_p))
        return a, b

    def mmprofile(self, a, end):
        """int32 [d+1, end]: H_m(a, j) for j < end, produced by the reference's own DFS"""
        out = np.zeros((self.d + 1, max(end, 1)), np.int32)
        if end > 0:
            self.lib.gkmref_mmprofile(a, end, out.ctypes.data_as(c_int_p))
        return out[:, :end]

    def row(self, a, start, end):
        res = np.zeros(max(end - start, 1))
        if end > start:
            self.lib.gkmref_row(a, start, end, res.ctypes.data_as(c_dbl_p))
        return res[: end - start]

    def rows_values(self, rows, nthreads, endcap=0, with_hist=True):
        """(K [len(rows), ld], H [len(rows), d+1, ld] or None, seconds): K[i, j] = K(rows[i], j) and H[i, m, j] = H_m(rows[i], j)
        for j < (endcap or rows[i]), every number produced by reference code, rows interleaved over nthreads"""
        rows = np.ascontiguousarray(rows, np.int32)
        ld = int(endcap if endcap > 0 else max(int(rows.max()), 1))
        K = np.zeros((len(rows), ld))
        H = np.zeros((len(rows), self.d + 1, ld), np.int32) if with_hist else None
        t = self.lib.gkmref_rows_values(rows.ctypes.data_as(c_int_p), len(rows), nthreads, endcap, K.ctypes.data_as(c_dbl_p),
                                        H.ctypes.data_as(c_int_p) if with_hist else None, ld)
        return K, H, t

    def rows_timed(self, rows, nthreads, endcap=0):
        rows = np.ascontiguousarray(rows, np.int32)
        cs = ctypes.c_double(0.0)
        t = self.lib.gkmref_rows_timed(rows.ctypes.data_as(c_int_p), len(rows), nthreads, endcap, ctypes.byref(cs))
        return t, cs.value
