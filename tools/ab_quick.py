#!/usr/bin/env python
"""resident-pass timing of selected diag flavors:  python tools/ab_quick.py n steps flavor [flavor..]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gkmqc_b200 import capi  # noqa: E402

n, steps = int(sys.argv[1]), int(sys.argv[2])
flavors = [int(x) for x in sys.argv[3:]] or [-1]
ktype = int(os.environ.get("KTYPE", "2"))
tmp = tempfile.mkdtemp()
pos, neg = bench.write_problem(tmp, n)
capi.load()
entries = n * (n - 1) // 2
for flavor in flavors:
    capi.set_option("diag_flavor", flavor)
    with capi.Problem(ktype, 11, 7, 3) as P:
        P.read(pos, neg)
        ms = P.bench_lower_resident(steps, 1, True)
        print("n %d type %d flavor %2d: %8.2f ms/pass  %7.1f M entries/s  launches %d" % (n, ktype, flavor, ms.mean(), entries / ms.mean() / 1e3, P.stats()["launches"]), flush=True)
