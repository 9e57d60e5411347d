#!/usr/bin/env python
"""resident pass of the tcgen05 candidate ("mma" variant) on n x 300 bp, with whatever library GKM_PYLIB names
   python tools/mma_time.py [n] [kernel_type]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
kt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
capi.set_option("kernel", "mma")
with capi.Problem(kt, 11, 7, 3) as P:
    P.add_block(bench.synth(n))
    ms = P.bench_lower_resident(2, 1, True)
    print("%s: n = %d type %d: %.2f ms per pass, %.1f M entries/s" % (os.environ.get("GKM_PYLIB", "product"), n, kt, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3))
