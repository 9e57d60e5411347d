#!/usr/bin/env python
"""wall-clock anatomy of gkm_main_pywrapper:  python tools/e2e_probe.py [n] [threads]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tmp = tempfile.mkdtemp(dir="/dev/shm")
pos, neg = bench.write_problem(tmp, n)
capi.load()
for it in range(4):
    kmat = np.zeros((n, n))
    t0 = time.perf_counter()
    ret, kmat, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=nt, verbosity=2 if it == 3 else 0, kmat=kmat)
    t1 = time.perf_counter()
    print("call %d: %.1f ms  (%.1f M entries/s)" % (it, 1e3 * (t1 - t0), n * (n - 1) / 2 / (t1 - t0) / 1e6), flush=True)
    del kmat
# same with a pre-touched matrix: isolates first-touch page faults of the caller's buffer
kmat = np.ones((n, n)); kmat[:] = 0
t0 = time.perf_counter()
capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=nt, verbosity=0, kmat=kmat)
print("pre-touched matrix: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
