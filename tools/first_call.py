#!/usr/bin/env python
"""what the FIRST call of a process costs (a slurm fan-out runs one bin per process): library load, CUDA context, first kernels
   python tools/first_call.py"""
import os, sys, time
t0 = time.perf_counter()
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
t1 = time.perf_counter()
lib = capi.load()
t2 = time.perf_counter()
nd = capi.device_count()
t3 = time.perf_counter()
ids = (capi.ctypes.c_int * 1)(0)
lib.gkmb200_set_devices(ids, 1)
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pos, neg = os.path.join(gold, "uni_pos.fa"), os.path.join(gold, "uni_neg.fa")
calls = []
for it in range(3):
    ta = time.perf_counter()
    ret, k, a, b = capi.main_pywrapper(pos, neg, kernel_type=4, L=10, k=6, d=3, nthreads=1, nmax=64)
    calls.append(time.perf_counter() - ta)
print("import numpy + capi %.2f s, dlopen %.3f s, device count (cuInit) %.2f s, first call (context, module, buffers) %.2f s, then %.4f / %.4f s"
      % (t1 - t0, t2 - t1, t3 - t2, calls[0], calls[1], calls[2]))
print("CUDA_MODULE_LOADING =", os.environ.get("CUDA_MODULE_LOADING"))
