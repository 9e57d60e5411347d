"""Host-side mirror of the reference's caller of the path (scripts/gkmsvm.py:67-99, ``computeGkmKernel``).

Same function name, same positional argument list and same return triple as the reference's driver, so the
parity tests read like its own code path, but bound to gkmqc_b200/bin/gkmkern_pylib.so.  The reference's
module itself works unchanged against this library; this mirror exists because the
GPU box has no copy of the reference tree.  Cross-validation with sklearn's SVC (gkmsvm.py:104-176) consumes
the path's output and is out of scope.
"""
import logging
import sys

import numpy as np

from . import capi

MAX_SEQS = 15000  # the reference allocates a fixed 15000 x 15000 matrix per call (gkmsvm.py:75)


def computeGkmKernel(args_gkm, max_seqs=MAX_SEQS):
    """args_gkm = [kernel_type, L, k, d, M, H, gamma, posfile, negfile, nthreads, verbosity]
    (the list gkmsvm.init builds, gkmsvm.py:183-195).  Returns (kmat, n_pseqs, n_nseqs): the symmetric
    n x n kernel matrix with unit diagonal, cut out of the padded matrix the C side filled."""
    kernel_type, L, k, d, M, H, gamma, posfile, negfile, nthreads, verbosity = args_gkm
    ret, padded, n_pseqs, n_nseqs = capi.main_pywrapper(
        posfile, negfile, kernel_type=kernel_type, L=L, k=k, d=d, M=M, H=H, gamma=gamma,
        nthreads=nthreads, verbosity=verbosity, nmax=max_seqs)
    if ret:
        logging.error("error on kernel construction")  # gkmsvm.py:90-92
        sys.exit()
    n = n_pseqs + n_nseqs
    kmat = padded[:n, :n]
    return np.maximum(kmat, kmat.T), n_pseqs, n_nseqs  # lower triangle -> symmetric (gkmsvm.py:96-97)
