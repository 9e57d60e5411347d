import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench
capi.load()
capi.set_option("kernel", "index")
for n, cols in [(10000, 0), (12000, 0), (14144, 0), (14144, 7104), (20000, 0), (28288, 0), (28288, 9440)]:
    capi.set_option("index_cols", str(cols))
    arr = bench.synth(n)
    with capi.Problem(2, 11, 7, 3) as P:
        P.add_many([a.tobytes().decode() for a in arr])
        ms = P.bench_lower_resident(2, 1, flush_l2=True)
        print(os.environ.get("GKM_IDX_MINB"), n, cols, "ms/pass %.1f" % ms.mean(), "us/row %.2f" % (1e3 * ms.mean() / n), "M entries/s %.0f" % (n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
