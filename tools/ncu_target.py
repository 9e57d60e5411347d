#!/usr/bin/env python
"""The short command that is profiled with `ncu --set full` (B200_PROFILING.md): the bench workload (BASELINE
configs[3], 50k x 300 bp, type 2, L=11 k=7 d=3) for one warm-up and one resident pass, nothing else.
    python tools/ncu_target.py [n] [kernel_type] [L] [k] [d] [seqlen]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
kt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L, k, d, seqlen = [int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((3, 11), (4, 7), (5, 3), (6, 300))]
capi.load()
with capi.Problem(kt, L, k, d) as P:
    P.add_block(bench.synth(n, seqlen=seqlen))
    ms = P.bench_lower_resident(1, 1, True)
    print("pass: %.2f ms, %d launches, variant %s" % (ms.mean(), P.stats()["launches"], bench.VARIANTS.get(P.stats()["kernel_variant"])))
