#!/usr/bin/env python
"""second part of tools/split_ab.py: full blocks larger than the 16 384-column default (option index_cols), greedy split"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
res = {}
capi.set_option("kernel", "index")
capi.set_option("index_split", "greedy")
os.makedirs("gpurun_out", exist_ok=True)
for n, colsl in ((50000, (16384, 18432, 20480, 22528, 25024, 25600)), (20000, (16384, 20000)), (28288, (16384, 18880, 22528)), (35000, (16384, 17504, 18432, 25600))):
    seqs = bench.synth(n)
    for cols in colsl:
        capi.set_option("index_cols", cols)
        with capi.Problem(2, 11, 7, 3) as P:
            P.add_block(seqs)
            ms = P.bench_lower_resident(2, 1, True)
            res["n%d_cols%d" % (n, cols)] = {"ms": float(ms.mean()), "launches": P.stats()["launches"], "layout": P.index_layout()}
            print(n, cols, res["n%d_cols%d" % (n, cols)], flush=True)
        json.dump(res, open("gpurun_out/split_ab2.json", "w"), indent=1)
