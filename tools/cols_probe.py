import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from gkmqc_b200 import capi
import bench
capi.load()
n = 10000
seqs = [a.tobytes().decode() for a in bench.synth(n)]
capi.set_option("kernel", "index")
for cols in (0, 5024, 3360, 2528):
    capi.set_option("index_cols", str(cols))
    with capi.Problem(2, 11, 7, 3, 50, 50.0, 1.0) as P:
        P.add_many(seqs); P.upload()
        ms = P.bench_lower_resident(3, 2, flush_l2=True)
        print("index_cols", cols, "ms/pass %.2f" % ms.mean(), flush=True)
