#!/usr/bin/env python
"""rectangles (scoring shape: every row probes every index block) with 20 000 columns: two equal blocks of 10 016 (the
former 16 384-column ceiling) against one block of 20 000 (ceiling 22 528); fused decision values, wall time of the call"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
res = {}
capi.set_option("kernel", "index")
os.makedirs("gpurun_out", exist_ok=True)
nsv, ntest = 20000, 11840
seqs = bench.synth(nsv + ntest)
alpha = np.random.default_rng(1).standard_normal(nsv)
for kt, L, k, d in ((2, 11, 7, 3), (4, 10, 6, 3)):
    ref = None
    for cols in (16384, 0):
        capi.set_option("index_cols", cols)
        with capi.Problem(kt, L, k, d) as P:
            P.add_block(seqs)
            P.upload()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                dv = P.decision_values(nsv, ntest, 0, nsv, alpha, bias=0.1)
                ts.append(time.perf_counter() - t0)
            key = "t%d_L%d_cols%d" % (kt, L, cols)
            res[key] = {"s": ts, "layout": P.index_layout(), "same_as_first": None if ref is None else bool(np.array_equal(dv, ref))}
            ref = dv if ref is None else ref
            print(key, res[key], flush=True)
        json.dump(res, open("gpurun_out/rect_ab.json", "w"), indent=1)
