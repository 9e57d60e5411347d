#!/usr/bin/env python
"""50k x 50k resident pass under chunk-size / column-block settings (GKM_CHUNK_MB is read once per process: set it outside)
   GKM_CHUNK_MB=256 python tools/chunk50k.py [index_cols]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 0
capi.set_option("index_cols", cols)
with capi.Problem(2, 11, 7, 3) as P:
    P.add_block(bench.synth(50000))
    ms = P.bench_lower_resident(2, 1, True)
    print("GKM_CHUNK_MB=%s (default 128) index_cols=%d: %.1f ms per pass, %d launches, layout %s" % (os.environ.get("GKM_CHUNK_MB", "default"), cols, ms.mean(), P.stats()["launches"], P.index_layout()), flush=True)
