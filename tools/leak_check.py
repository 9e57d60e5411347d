#!/usr/bin/env python
"""a long-lived host: 60 calls of gkm_main_pywrapper with changing parameters and sizes; device memory in use and the
process's resident set must level off (the per-GPU block pool is bounded, nothing of a call survives it)
    python tools/leak_check.py"""
import ctypes, os, sys, tempfile, time, resource
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
lib = capi.load()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
files = {n: bench.write_problem(tmp, n, tag="_%d" % n) for n in (3000, 6000, 10000)}
import subprocess
def gpu_used():
    out = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
    return int(out.strip().splitlines()[0])
hist = []
for it in range(60):
    n = (3000, 6000, 10000)[it % 3]
    kt, L, k, d = [(2, 11, 7, 3), (4, 10, 6, 3), (2, 12, 8, 2), (4, 11, 7, 3), (3, 9, 5, 2)][it % 5]
    km = np.zeros((n, n))
    ret, km, a, b = capi.main_pywrapper(files[n][0], files[n][1], kernel_type=kt, L=L, k=k, d=d, nthreads=1, kmat=km)
    assert ret == 0 and km[n - 1, n - 1] == 1.0, capi.last_error()
    del km
    if it % 5 == 4:
        hist.append((it + 1, gpu_used(), resource.getrusage(resource.RUSAGE_SELF).ru_maxrss // 1024))
        print("after %2d calls: GPU memory in use %d MiB, max RSS %d MiB" % hist[-1], flush=True)
assert hist[-1][1] <= hist[len(hist) // 2][1] + 64, "device memory keeps growing"
lib.gkmb200_trim()
print("after gkmb200_trim: GPU memory in use %d MiB" % gpu_used())
