#!/usr/bin/env python
"""BASELINE configs[0]: 1 000 synthetic 300-bp sequences, L=11 k=7 d=3, end to end through gkm_main_pywrapper, per kernel choice
   python tools/config0_time.py"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
lib = capi.load()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
n = 1000
pos, neg = bench.write_problem(tmp, n, tag="_1k")
for kern in ("auto", "index", "diag"):
    capi.set_option("kernel", kern)
    for kt in (2, 4):
        walls = []
        for it in range(6):
            km = np.zeros((15000, 15000))          # what gkmsvm.py:75 allocates
            t0 = time.perf_counter()
            ret, km, a, b = capi.main_pywrapper(pos, neg, kernel_type=kt, L=11, k=7, d=3, nthreads=1, kmat=km)
            walls.append(time.perf_counter() - t0)
            assert ret == 0
            del km
        st = capi.gkmb200_stats(); lib.gkmb200_get_stats(None, capi.ctypes.byref(st))
        print("kernel = %-5s type %d: %.2f ms end to end (best of 5 warm), device %.2f ms, variant %s" %
              (kern, kt, 1e3 * min(walls[1:]), st.kernel_ms, bench.VARIANTS.get(st.kernel_variant)), flush=True)
capi.set_option("kernel", "auto")
