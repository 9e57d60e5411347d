"""quick A/B of the kernel variants on the bench workload (resident pass), plus a histogram cross-check"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
kt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["index", "diag"]
L, KK, D = (int(x) for x in sys.argv[4].split(",")) if len(sys.argv) > 4 else (11, 7, 3)
capi.load()
arr = bench.synth(n)
seqs = [a.tobytes().decode() for a in arr]
H = {}
for v in variants:
    capi.set_option("kernel", v)
    with capi.Problem(kt, L, KK, D, 50, 50.0, 1.0) as P:
        P.add_many(seqs)
        t0 = time.time(); P.upload(); t1 = time.time()
        ms = P.bench_lower_resident(3, 2, flush_l2=True)
        st = P.stats()
        print(v, "variant", st["kernel_variant"], "upload %.1f ms" % (1e3 * (t1 - t0)), "ms/pass", ms, "-> %.1f M entries/s" % (n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
        H[v] = P.hist_block(n - 64, 64, 0, n - 64)
        K = P.kernel_block(n - 64, 64, 0, n - 64)
        H[v + "K"] = K
if len(variants) > 1:
    a, b = variants[0], variants[1]
    print("hist equal:", np.array_equal(H[a], H[b]), "kernel equal:", np.array_equal(H[a + "K"], H[b + "K"]))
