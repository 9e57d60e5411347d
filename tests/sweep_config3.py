#!/usr/bin/env python
"""BASELINE.json configs[2]: l/k/d sweep (l = 10..14, k = 6..8, d = 2..4) on 20 000 synthetic 300-bp
sequences -- integer mismatch histograms bit-exact against the REFERENCE, kernel doubles bit-identical.

For every word length L the unmodified reference (oracle/_ref, probe build for L > 12) is opened once on the
full 20k problem with d = 4; its DFS gives H_m(a, j), m = 0..4, for a stated subsample of rows a (all j <= a,
the diagonal included).  H for d = 2, 3 are prefixes of that.  The GPU histograms of every valid (L, k, d)
triple are compared with those integers; the GPU doubles are compared with
    K = (sum_m w[m] H_m) / (sqnorm_a sqnorm_j),   sqnorm = sqrt(sum_m w[m] H_m(x, x))
evaluated in the reference's operation order from the reference's integers and the reference's own w[m]
(weight routines of libgkm.c through the probe).  Three triples are additionally checked against the doubles
the reference itself returns (gkmkernel_kernelfunc_batch_all).  Also times one resident pass per (L, d).

    python tests/sweep_config3.py [n] [rows...]      (writes gpurun_out/sweep_config3.json)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import pyoracle  # noqa: E402
from gkmqc_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rows = [int(x) for x in sys.argv[2:]] or [1, 2, 777, n // 2 + 1, n - 2, n - 1]
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    pos, neg = bench.write_problem(tmp, n)
    capi.load()
    if not pyoracle.have_ref():
        raise SystemExit("oracle/_ref is not present: build it where /root/reference exists")
    out = {"n": n, "rows": rows, "triples": [], "timing": [], "GKM_KERNEL": os.environ.get("GKM_KERNEL", "auto")}
    nbad = 0
    for L in range(10, 15):
        t0 = time.time()
        h = pyoracle.RefHook(pos, neg, 2, L, L - 4, 4)
        Href = {}
        for a in rows:
            Href[a] = h.mmprofile(a, a + 1).T.copy()          # [a+1, 5]: columns j = 0..a, diagonal last
        spot = {}
        if L in (10, 11, 13):
            spot = {a: h.row(a, 0, a) for a in rows}          # the reference's own doubles at (L, L-4, 4)
        h.close()
        # the diagonal of every column j is needed for sqnorm_j: take it from the GPU histograms of the
        # d = 4 problem AFTER they have been checked on the sampled rows (same kernel, same integers)
        print("L=%d: reference opened and %d rows profiled in %.1f s" % (L, len(rows), time.time() - t0), flush=True)
        for k in (6, 7, 8):
            for d in (2, 3, 4):
                if d > L - k:
                    continue
                w = pyoracle.ref_weights(2, L, k)[: d + 1]
                with capi.Problem(2, L, k, d) as P:
                    P.read(pos, neg)
                    sq = P.sqnorm()
                    ok_h = ok_k = ok_sq = True
                    for a in rows:
                        Hg = P.hist_block(a, 1, 0, a)[0]                       # [a, d+1]
                        Hd = P.hist_block(a, 1, a, 1)[0, 0]                    # diagonal entry
                        ok_h &= bool(np.array_equal(Hg, Href[a][:a, : d + 1]) and np.array_equal(Hd, Href[a][a, : d + 1]))
                        s = 0.0
                        for m in range(d + 1):
                            s = s + w[m] * float(Href[a][a, m])
                        ok_sq &= bool(np.sqrt(s) == sq[a])
                        kraw = np.zeros(a)
                        for m in range(d + 1):
                            kraw = kraw + w[m] * Href[a][:a, m].astype(np.float64)
                        Kexp = kraw / (sq[a] * sq[:a])
                        Kg = P.kernel_block(a, 1, 0, a)[0]
                        ok_k &= bool(np.array_equal(Kg, Kexp))
                        if d == 4 and k == L - 4 and a in spot:
                            ok_k &= bool(np.array_equal(Kg, spot[a]))
                    rec = {"L": L, "k": k, "d": d, "hist_bit_exact": ok_h, "sqnorm_bit_identical": ok_sq,
                           "kernel_bit_identical": ok_k, "vs_reference_doubles": bool(d == 4 and k == L - 4 and bool(spot))}
                    if k == 6 or (L - k < 4 and d == L - k):
                        pass
                    out["triples"].append(rec)
                    nbad += not (ok_h and ok_k and ok_sq)
                    print(rec, flush=True)
        for d in (2, 3, 4):
            k = min(8, L - d)
            with capi.Problem(2, L, k, d) as P:
                P.read(pos, neg)
                ms = P.bench_lower_resident(1, 1, True)
                rate = n * (n - 1) / 2 / ms.mean() / 1e3
                variant = {1: "lmer", 2: "diag", 3: "mma", 4: "index"}.get(P.stats()["kernel_variant"], "?")
                out["timing"].append({"L": L, "d": d, "ms_per_pass": float(ms.mean()), "M_entries_per_s": rate, "kernel": variant})
                print("timing L=%d d=%d: %.1f ms  %.1f M entries/s (%s)" % (L, d, ms.mean(), rate, variant), flush=True)
    out["all_ok"] = nbad == 0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_config3_%s.json" % out["GKM_KERNEL"]), "w") as f:
        json.dump(out, f, indent=1)
    print("sweep: %d triples, %d failures" % (len(out["triples"]), nbad))
    sys.exit(1 if nbad else 0)


if __name__ == "__main__":
    main()
