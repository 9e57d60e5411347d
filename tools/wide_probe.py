import os, sys
sys.path.insert(0, os.getcwd())
from gkmqc_b200 import capi
import bench
n = 10000
capi.load()
seqs = [a.tobytes().decode() for a in bench.synth(n)]
capi.set_option("kernel", "index")
for wide, minb in (("0", "1"), ("1", "1"), ("0", "2")):
    capi.set_option("index_wide", wide)
    os.environ["GKM_IDX_MINB"] = minb
    with capi.Problem(2, 11, 7, 3, 50, 50.0, 1.0) as P:
        P.add_many(seqs); P.upload()
        ms = P.bench_lower_resident(3, 2, flush_l2=True)
        print("type 2  index_wide=%s  MINB=%s: %.2f ms/pass" % (wide, minb, ms.mean()), flush=True)
