"""fused decision values through the index kernel: time vs number of test rows"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench
nsv = 10000
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
capi.load()
arr = bench.synth(nsv + nt, seed=99)
alpha = np.random.default_rng(3).standard_normal(nsv)
with capi.Problem(2, 11, 7, 3) as P:
    P.add_many([a.tobytes().decode() for a in arr])
    P.upload()
    for rows in (2368, 9472, nt, nt):
        t0 = time.perf_counter()
        dv = P.decision_values(nsv, rows, 0, nsv, alpha, bias=-0.1)
        dt = time.perf_counter() - t0
        print("decision %d x %d: %.1f ms, %.2f us/row" % (rows, nsv, 1e3 * dt, 1e6 * dt / rows), flush=True)
