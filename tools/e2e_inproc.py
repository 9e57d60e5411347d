#!/usr/bin/env python
"""gkm_main_pywrapper in ONE process on 1, 2, ... all visible GPUs (the library's default: every sm_100 device, chunks from
one shared queue): wall time per call at 50k and 10k, the library's own anatomy (verbosity 3) for the last call
    python tools/e2e_inproc.py [n ...]"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gkmqc_b200 import capi
lib = capi.load()
ndev = capi.device_count()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
res = {"devices_visible": ndev, "cores": len(os.sched_getaffinity(0))}
for n in [int(a) for a in sys.argv[1:]] or [50000, 10000]:
    pos, neg = bench.write_problem(tmp, n, tag="_%d" % n)
    g = 1
    while g <= ndev:
        ids = (capi.ctypes.c_int * g)(*range(g))
        assert lib.gkmb200_set_devices(ids, g) == 0
        walls = []
        for it in range(4):
            km = np.zeros((n, n))
            t0 = time.perf_counter()
            ret, km, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=1, verbosity=3 if it == 3 else 0, kmat=km)
            walls.append(time.perf_counter() - t0)
            assert ret == 0, capi.last_error()
            if it == 3:
                ok = bool(km[n - 1, n - 1] == 1.0 and km[n - 1, : n - 1].min() > 0 and km[1, 0] > 0)
            del km
        st = capi.gkmb200_stats(); lib.gkmb200_get_stats(None, capi.ctypes.byref(st))
        res["n%d_g%d" % (n, g)] = {"walls_s": walls, "M_entries_s": n * (n - 1) / 2 / min(walls[1:]) / 1e6, "stats": st.as_dict(), "filled": ok}
        print(n, g, walls, st.as_dict(), flush=True)
        g *= 2
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "e2e_inproc.json"), "w"), indent=1)
