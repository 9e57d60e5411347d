#!/usr/bin/env python
"""The resident kernel matrix of the SVM consumer shared by the GPUs of one process: every GPU writes its
chunks into the first GPU's matrix over NVLink (peer stores from the epilogue, no gather step).
Runs the tests of tests/test_gpu_svm.py that cover it, then times the pass with one GPU and with all of them.
    python tools/resident_p2p.py [out.json]          (needs >= 2 B200s for the shared half)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gkmqc_b200 import capi  # noqa: E402

ACGT = np.frombuffer(b"ACGT", np.uint8)


def timed(ids, n, kt, L, k, d, seqlen, reps=3):
    lib = capi.load()
    arr = (capi.ctypes.c_int * len(ids))(*ids)
    assert lib.gkmb200_set_devices(arr, len(ids)) == 0, capi.last_error()
    x = ACGT[np.random.default_rng(1234).integers(0, 4, (n, seqlen))]
    out = []
    for rep in range(reps + 1):
        with capi.Problem(kt, L, k, d, 50, 50.0, 1.0) as P:
            P.add_block(x)
            P.upload()
            t0 = time.perf_counter()
            P.resident_matrix(0, 0)          # index build on every GPU, the pass, the mirror; waits for the device
            dt = time.perf_counter() - t0
            st = P.stats()
            if rep == reps:
                band = P.resident_matrix(n - 3, 3)
        if rep > 0:
            out.append(dt)
    return {"gpus": len(ids), "n": n, "kernel_type": kt, "L": L, "k": k, "d": d, "seqlen": seqlen, "ms": [1e3 * t for t in out],
            "ms_best": 1e3 * min(out), "devices_used": st["devices"], "variant": st["kernel_variant"], "launches": st["launches"],
            "checksum_last_rows": float(band.sum())}


def main():
    import pytest
    res = {"what": "resident symmetric kernel matrix (gkmb200_resident_rows, nrows = 0), wall time of the call on a fresh problem "
                   "whose sequences are already uploaded: index build + lower-triangle pass + mirror", "timing": []}
    ndev = capi.device_count()
    res["gpus_visible"] = ndev
    t0 = time.perf_counter()
    rc = pytest.main(["-q", "-x", "-m", "gpu", "-k", "resident_matrix", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_svm.py")])
    res["pytest_resident_matrix_tests"] = {"rc": int(rc), "seconds": time.perf_counter() - t0}
    for n, kt, L, k, d, ln in ((10000, 2, 11, 7, 3, 300), (10000, 4, 10, 6, 3, 600), (20000, 2, 11, 7, 3, 300)):
        for ids in ([0], list(range(ndev))) if ndev >= 2 else ([0],):
            r = timed(ids, n, kt, L, k, d, ln)
            res["timing"].append(r)
            print(json.dumps(r), flush=True)
    same = {}
    for r in res["timing"]:
        same.setdefault((r["n"], r["kernel_type"], r["seqlen"]), set()).add(r["checksum_last_rows"])
    res["same_last_rows_on_every_gpu_set"] = all(len(v) == 1 for v in same.values())
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "resident_p2p.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
    print("resident_p2p done: pytest rc", rc, "same rows:", res["same_last_rows_on_every_gpu_set"])
    return int(rc) or (0 if res["same_last_rows_on_every_gpu_set"] else 4)


if __name__ == "__main__":
    sys.exit(main())
