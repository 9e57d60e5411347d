#!/usr/bin/env python
"""BASELINE.json configs[3] and [4] on the GPUs that are visible (single process, all selected GPUs):
   config 4: 50 000 x 50 000 kernel matrix (resident pass + full host path into a numpy matrix if RAM allows)
   config 5: batch scoring, `ntest` synthetic test sequences x 10 000 support vectors, decision values fused on
             the device (no test x SV matrix leaves the GPU) and, for a slice, the dense block.
   python tools/run_config45.py [n4] [ntest]"""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkmqc_b200 import capi

n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ntest = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
nsv = 10000
capi.load()
res = {"devices": capi.device_count()}
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

# ---- config 4 ----
pos, neg = bench.write_problem(tmp, n4)
entries = n4 * (n4 - 1) // 2
with capi.Problem(2, 11, 7, 3) as P:
    P.read(pos, neg)
    ms = P.bench_lower_resident(1, 1, True)
    res["config4_resident"] = {"n": n4, "ms": float(ms.mean()), "M_entries_per_s": entries / ms.mean() / 1e3, "gpus": 1}
    print(res["config4_resident"], flush=True)
mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
need_gb = n4 * n4 * 8 / 2**30
if mem_gb > need_gb * 1.5:
    kmat = np.zeros((n4, n4))
    t0 = time.perf_counter()
    ret, kmat, a, b = capi.main_pywrapper(pos, neg, kernel_type=2, L=11, k=7, d=3, nthreads=16, verbosity=0, kmat=kmat)
    dt = time.perf_counter() - t0
    assert ret == 0
    low = kmat[n4 - 1, : n4 - 1]
    res["config4_host"] = {"n": n4, "s": dt, "M_entries_per_s": entries / dt / 1e6, "gpus": capi.device_count(),
                           "diag_ok": bool(kmat[12345 % n4, 12345 % n4] == 1.0), "last_row_min": float(low.min()), "last_row_max": float(low.max())}
    print(res["config4_host"], flush=True)
    del kmat
else:
    res["config4_host"] = "skipped: host RAM %.0f GB < 1.5 x %.0f GB" % (mem_gb, need_gb)
    print(res["config4_host"], flush=True)

# ---- config 5 ----
arr = bench.synth(nsv + ntest, seed=99)
with capi.Problem(2, 11, 7, 3) as P:
    t0 = time.perf_counter()
    for i in range(nsv + ntest):
        P.add(arr[i].tobytes())
    t_add = time.perf_counter() - t0
    P.upload()
    alpha = np.random.default_rng(3).standard_normal(nsv)
    P.decision_values(nsv, min(ntest, 2368), 0, nsv, alpha, bias=-0.1)   # first batch: builds the SV index, sizes the scratch
    t0 = time.perf_counter()
    dv = P.decision_values(nsv, ntest, 0, nsv, alpha, bias=-0.1)
    dt = time.perf_counter() - t0
    res["config5_decision"] = {"ntest": ntest, "nsv": nsv, "s": dt, "M_entries_per_s": ntest * nsv / dt / 1e6, "add_sequences_s": t_add}
    print(res["config5_decision"], flush=True)
    # dense block of the first 4096 tests, and agreement of the fused values with it
    t0 = time.perf_counter()
    K = P.kernel_block(nsv, 4096, 0, nsv)
    dtb = time.perf_counter() - t0
    err = float(np.max(np.abs(K @ alpha - 0.1 - dv[:4096])))
    res["config5_block"] = {"rows": 4096, "nsv": nsv, "s": dtb, "M_entries_per_s": 4096 * nsv / dtb / 1e6, "max_abs_diff_fused_vs_dense": err}
    print(res["config5_block"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config45.json"), "w"), indent=1)
