#!/usr/bin/env python
"""A/B of host-pipeline knobs on the end-to-end call (gkm_main_pywrapper, FASTA -> fresh numpy matrix), same process.
usage: e2e_ab.py ENVVAR valueA valueB [n] [calls]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
var, va, vb = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 8
tmp = tempfile.mkdtemp(dir="/dev/shm")
pos, neg = bench.write_problem(tmp, n)
capi.load()
res = {va: [], vb: []}
for it in range(2 * calls + 2):
    v = va if it % 2 == 0 else vb
    if v == "-": os.environ.pop(var, None)
    else: os.environ[var] = v
    kmat = np.zeros((n, n))
    t0 = time.perf_counter()
    ret, kmat, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=8, verbosity=0, kmat=kmat)
    t1 = time.perf_counter()
    st = capi.last_stats() if hasattr(capi, "last_stats") else None
    if it >= 2: res[v].append(1e3 * (t1 - t0))
    del kmat
for v in (va, vb):
    x = np.array(res[v]); print("%s=%s: mean %.2f ms  median %.2f  min %.2f  max %.2f  (%d calls)" % (var, v, x.mean(), np.median(x), x.min(), x.max(), len(x)), flush=True)
