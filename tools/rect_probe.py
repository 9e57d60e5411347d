"""dense test x SV block through the index kernel (RECT mode): timing per row, for ncu"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench
nsv = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
capi.load()
arr = bench.synth(nsv + nt, seed=99)
with capi.Problem(2, 11, 7, 3) as P:
    P.add_many([a.tobytes().decode() for a in arr])
    P.upload()
    for it in range(3):
        t0 = time.perf_counter()
        K = P.kernel_block(nsv, nt, 0, nsv)
        dt = time.perf_counter() - t0
        st = P.stats()
        print("rect %d x %d: wall %.1f ms, kernels %.2f ms, %.2f us/row, variant %d" % (nt, nsv, 1e3 * dt, st["kernel_ms"], 1e3 * st["kernel_ms"] / nt, st["kernel_variant"]), flush=True)
