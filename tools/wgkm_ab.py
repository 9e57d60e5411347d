"""wgkm (kernel type 4) resident pass at 10k: slot formats and CTAs per SM"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
capi.load()
seqs = [a.tobytes().decode() for a in bench.synth(n)]
capi.set_option("kernel", "index")
for wide, minb in (("1", None), ("0", "1"), ("0", "2"), ("0", None)):
    capi.set_option("index_wide", wide)
    if minb is None: os.environ.pop("GKM_IDX_MINB", None)
    else: os.environ["GKM_IDX_MINB"] = minb
    with capi.Problem(4, 11, 7, 3, 50, 50.0, 1.0) as P:
        P.add_many(seqs); P.upload()
        ms = P.bench_lower_resident(3, 2, flush_l2=True)
        print("type 4  index_wide=%s  GKM_IDX_MINB=%s: %.2f ms/pass  %.1f M entries/s" % (wide, minb, ms.mean(), n * (n - 1) / 2 / ms.mean() / 1e3), flush=True)
capi.set_option("index_wide", "0")
