"""per-row time of the index kernel on both sides of the shared-memory / L1 split steps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gkmqc_b200 import capi
import bench
capi.load()
capi.set_option("kernel", "index")
for n in [int(x) for x in (sys.argv[1:] or ["6000", "6300", "6500", "8000", "8300", "8500", "10000", "10700", "10800", "12000"])]:
    seqs = [a.tobytes().decode() for a in bench.synth(n)]
    with capi.Problem(2, 11, 7, 3, 50, 50.0, 1.0) as P:
        P.add_many(seqs); P.upload()
        ms = P.bench_lower_resident(3, 2, flush_l2=True)
        hist_kb = 2 * 4 * ((n + 31) // 32 * 32) / 1024
        print("n %6d  hist %.1f KB/CTA  %.2f ms/pass  %.3f us/row  %.1f M entries/s" % (n, hist_kb, ms.mean(), 1e3 * ms.mean() / n, n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
