#!/usr/bin/env python
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
n = 10000
tmp = tempfile.mkdtemp(dir="/dev/shm")
pos, neg = bench.write_problem(tmp, n)
capi.load()
for it in range(10):
    kmat = np.zeros((n, n))
    t0 = time.perf_counter()
    ret, kmat, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=8, verbosity=int(os.environ.get("V","2")), kmat=kmat)
    t1 = time.perf_counter()
    sys.stdout.flush()
    print("CALL %d: %.1f ms" % (it, 1e3 * (t1 - t0)), flush=True)
    t0 = time.perf_counter(); del kmat; print("  del %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip())
