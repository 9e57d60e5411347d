#!/usr/bin/env python
"""does the nvidia-smi clock sampler (bench.py) disturb the end-to-end call?"""
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
def run(tag, reps=3):
    for it in range(reps):
        kmat = np.zeros((n, n))
        t0 = time.perf_counter()
        ret, kmat, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=8, verbosity=0, kmat=kmat)
        print("%s call %d: %.1f ms" % (tag, it, 1e3 * (time.perf_counter() - t0)), flush=True)
        del kmat
run("plain")
P = capi.Problem(2, 11, 7, 3); P.read(pos, neg); P.upload()
ms = P.bench_lower_resident(3, 3, True); print("resident", ms)
run("after-resident")
s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
run("with-sampler")
print(s.finish())
run("sampler-stopped")
