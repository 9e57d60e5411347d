#!/usr/bin/env python
"""resident-pass timing for a parameter set:  python tools/time_cfg.py n ktype L k d [steps]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
n, ktype, L, k, d = [int(x) for x in sys.argv[1:6]]
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
tmp = tempfile.mkdtemp()
pos, neg = bench.write_problem(tmp, n)
capi.load()
with capi.Problem(ktype, L, k, d) as P:
    P.read(pos, neg)
    ms = P.bench_lower_resident(steps, 1, True)
    print("n %d type %d L %d k %d d %d: %.2f ms/pass  %.1f M entries/s" % (n, ktype, L, k, d, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
