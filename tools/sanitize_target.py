#!/usr/bin/env python
"""small, ragged problems through every kernel variant and entry point -- the command run under
`compute-sanitizer --tool memcheck` (B200_PROFILING.md: one tool per gpurun call)
    python tools/sanitize_target.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_seqs
from gkmqc_b200 import capi

seqs = random_seqs(70, 180, seed=3, ragged=True)
seqs = [s if len(s) >= 12 else s + "ACGTACGTACGT" for s in seqs] + ["A" * 40, "ACGT" * 200, "T" * 11 + "ACG"]
ref = {}
for kt in (2, 4):
    for variant in ("diag", "index", "mma", "lmer"):
        capi.set_option("kernel", variant)
        for cols in (0, 32):
            capi.set_option("index_cols", cols)
            with capi.Problem(kt, 11, 7, 3) as P:
                P.add_many(seqs)
                K = P.kernel_lower()
                H = P.hist_block(5, 40, 0, 30)
                B = P.kernel_block(40, 20, 0, 40)
                dv = P.decision_values(40, 20, 0, 40, np.linspace(-1, 1, 40), 0.25)
                if variant == "index":
                    P.image()
            key = kt
            if key in ref:
                assert np.array_equal(K, ref[key][0]) and np.array_equal(H, ref[key][1]) and np.array_equal(B, ref[key][2]), (kt, variant, cols)
                assert np.allclose(dv, ref[key][3], rtol=1e-12)
            else:
                ref[key] = (K, H, B, dv)
        print("type", kt, variant, "ok", flush=True)
capi.set_option("kernel", "auto"); capi.set_option("index_cols", 0)
print("sanitize target done")
