#!/usr/bin/env python
"""A/B of the diag kernel's pipe-balancing flavors on the device (resident passes, CUDA events).
   python tools/ab_flavors.py [n] [steps]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gkmqc_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tmp = tempfile.mkdtemp()
pos, neg = bench.write_problem(tmp, n)
capi.load()
entries = n * (n - 1) // 2
for name in ("lop3", "shf", "popc", "iadd3", "imad", "imadhi", "lop3+imad", "lop3+popc", "lop3+imad+popc"):
    try:
        print("microbench %-16s %10.1f Gop/s" % (name, capi.microbench(name)))
    except capi.GkmError as e:
        print("microbench", name, "failed:", e)
for ktype in (2, 4):
    for flavor in ([-1] + list(range(8)) if ktype == 2 else [-1]):
        capi.set_option("diag_flavor", flavor)
        with capi.Problem(ktype, 11, 7, 3) as P:
            P.read(pos, neg)
            ms = P.bench_lower_resident(steps, 2, True)
            print("type %d flavor %2d: %8.2f ms/pass  %7.1f M entries/s" % (ktype, flavor, ms.mean(), entries / ms.mean() / 1e3))
capi.set_option("diag_flavor", -1)
capi.set_option("kernel", "lmer")
with capi.Problem(2, 11, 7, 3) as P:
    P.read(pos, neg)
    ms = P.bench_lower_resident(1, 1, True)
    print("lmer kernel      : %8.2f ms/pass  %7.1f M entries/s" % (ms.mean(), entries / ms.mean() / 1e3))
