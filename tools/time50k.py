"""resident pass of the n x n lower triangle (default 50 000), one GPU"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
capi.load()
arr = bench.synth(n)
for kt in (2,):
    with capi.Problem(kt, 11, 7, 3, 50, 50.0, 1.0) as P:
        P.add_many([a.tobytes().decode() for a in arr])
        ms = P.bench_lower_resident(2, 1, flush_l2=True)
        st = P.stats()
        print(n, "type", kt, "variant", st["kernel_variant"], "launches", st["launches"], "ms/pass", ms, "us/row %.2f" % (1e3 * ms.mean() / n), "M entries/s %.0f" % (n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
