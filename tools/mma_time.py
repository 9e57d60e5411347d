#!/usr/bin/env python
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi
n = int(sys.argv[1]); var = sys.argv[2]
capi.load(); capi.set_option("kernel", var)
tmp = tempfile.mkdtemp(); pos, neg = bench.write_problem(tmp, n)
with capi.Problem(2, 11, 7, 3) as P:
    P.read(pos, neg)
    ms = P.bench_lower_resident(1, 1, True)
    print("%s n=%d: %.1f ms/pass  %.1f M entries/s" % (var, n, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3))
