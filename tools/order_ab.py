#!/usr/bin/env python
"""chunk order of the triangle (GKM_CHUNK_ORDER: widest rows first (default) or ascending): end-to-end wall and device span
   python tools/order_ab.py"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
lib = capi.load()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
cases = [("10k x 300 bp t2 L11", 10000, 300, dict(kernel_type=2, L=11, k=7, d=3)), ("10k x 300 bp t4 L11", 10000, 300, dict(kernel_type=4, L=11, k=7, d=3)),
         ("10k x 600 bp t4 L10", 10000, 600, dict(kernel_type=4, L=10, k=6, d=3)), ("20k x 300 bp t2 L11 d4", 20000, 300, dict(kernel_type=2, L=11, k=7, d=4)),
         ("50k x 300 bp t2 L11", 50000, 300, dict(kernel_type=2, L=11, k=7, d=3))]
for tag, n, sl, kw in cases:
    pos, neg = bench.write_problem(tmp, n, seqlen=sl, tag="_%d_%d" % (n, sl))
    for order in ("desc", "asc"):
        os.environ["GKM_CHUNK_ORDER"] = order
        walls = []
        for it in range(4):
            km = np.zeros((n, n))
            t0 = time.perf_counter()
            ret, km, a, b = capi.main_pywrapper(pos, neg, nthreads=1, verbosity=0, kmat=km, **kw)
            walls.append(time.perf_counter() - t0)
            assert ret == 0
            del km
        st = capi.gkmb200_stats(); lib.gkmb200_get_stats(None, capi.ctypes.byref(st))
        print("%-24s %-4s wall %.1f ms (best of 3 warm), device span %.1f ms, scatter %.1f ms" % (tag, order, 1e3 * min(walls[1:]), st.kernel_ms, st.scatter_ms), flush=True)
