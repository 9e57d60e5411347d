#!/usr/bin/env python
"""bring-up of the tcgen05 candidate: tiny problems first, compared with the diag kernel"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
capi.load()
from conftest_shim import random_seqs
def run(n, length, ktype=2, L=11, k=7, d=3, ragged=False):
    seqs = random_seqs(n, length, 5, ragged)
    out = {}
    for var in ("diag", "mma"):
        capi.set_option("kernel", var)
        with capi.Problem(ktype, L, k, d) as P:
            P.add_many(seqs)
            out[var] = (P.sqnorm(), P.hist_block(0, n, 0, n), P.kernel_lower(), P.stats()["kernel_variant"])
    ok = all(np.array_equal(out["diag"][i], out["mma"][i]) for i in range(3))
    print("n=%d len=%d type=%d L=%d d=%d ragged=%s: variant %d  %s" % (n, length, ktype, L, d, ragged, out["mma"][3], "OK" if ok else "MISMATCH"), flush=True)
    if not ok:
        H0, H1 = out["diag"][1], out["mma"][1]
        bad = np.argwhere(H0 != H1)
        print("  first diffs:", bad[:4].tolist(), H0[tuple(bad[0][:2])], H1[tuple(bad[0][:2])])
    return ok
ok = run(4, 40) and run(9, 300) and run(20, 300, ktype=4) and run(33, 500, ragged=True, L=10, k=6, d=4) and run(16, 2047, L=16, k=12, d=4)
if ok and len(sys.argv) > 1:
    n = int(sys.argv[1])
    tmp = tempfile.mkdtemp()
    pos, neg = bench.write_problem(tmp, n)
    capi.set_option("kernel", "mma")
    with capi.Problem(2, 11, 7, 3) as P:
        P.read(pos, neg)
        ms = P.bench_lower_resident(2, 1, True)
        print("mma kernel n=%d: %.1f ms/pass  %.1f M entries/s" % (n, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3))
