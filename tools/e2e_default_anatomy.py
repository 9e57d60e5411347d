#!/usr/bin/env python
"""where the end-to-end call of gkmQC's default configuration (wgkm L=10 k=6 d=3, 5k + 5k x 600 bp) spends its time:
verbosity 3 prints the library's own anatomy of the last of four calls
    python tools/e2e_default_anatomy.py"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
n = 10000
pos, neg = bench.write_problem(tmp, n, seed=4321, seqlen=600, tag="_600")
for it in range(4):
    km = np.zeros((n, n))
    t0 = time.perf_counter()
    ret, km, a, b = capi.main_pywrapper(pos, neg, kernel_type=4, L=10, k=6, d=3, M=50, H=50.0, nthreads=1, verbosity=3 if it == 3 else 0, kmat=km)
    print("call %d: %.1f ms" % (it, 1e3 * (time.perf_counter() - t0)), flush=True)
    del km
