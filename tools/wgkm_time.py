#!/usr/bin/env python
"""resident passes of the weighted kernel type (wgkm, type 4) with whatever library GKM_PYLIB names
   python tools/wgkm_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
for tag, arr, args in (("10k x 300 bp L=11", bench.synth(10000), (4, 11, 7, 3)), ("10k x 600 bp L=10", bench.synth(10000, seed=4321, seqlen=600), (4, 10, 6, 3)),
                       ("20k x 300 bp L=11", bench.synth(20000), (4, 11, 7, 3))):
    with capi.Problem(*args) as P:
        P.add_block(arr)
        ms = P.bench_lower_resident(3, 1, True)
        n = len(arr)
        print("%s: %s: %.2f ms per pass, %.1f M entries/s" % (os.environ.get("GKM_PYLIB", "product"), tag, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
