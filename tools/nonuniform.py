#!/usr/bin/env python
"""the same 10k x 10k problem on NON-uniform, genome-like inputs, to exercise the hit path.

For every workload the resident pass is timed with the index kernel, with kernel = auto and with the bit-sliced
kernel, and the integer histograms + kernel doubles of a band of rows are compared between the index and the
bit-sliced kernel (two independent constructions: neighbour enumeration against dense diagonals).

    python tools/nonuniform.py [n] [kernel_type] [out.json]

Workloads (300 bp, seeded):
  uniform     i.i.d. A,C,G,T (the bench workload)
  at_rich     i.i.d. with 41 % GC (human-genome composition)
  polyA_10    10 % of the sequences carry a poly-A or poly-T tract of 20..60 bp
  str_5       5 % carry a (CA)n / (GT)n / (AT)n ... dinucleotide repeat of 30..80 bp
  family_20   20 % carry a 150-bp member of one repeat family (consensus with 10 % divergence per copy, Alu-like)
  dup_5       5 % are exact copies of other sequences (overlapping peaks)
  mixed       all of the above on AT-rich background
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi

SEQLEN = 300
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def background(rng, n, gc=0.5):
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return rng.choice(4, size=(n, SEQLEN), p=p).astype(np.uint8)


def plant_homopolymer(rng, x, frac):
    idx = rng.choice(len(x), size=int(frac * len(x)), replace=False)
    for i in idx:
        ln = int(rng.integers(20, 61)); at = int(rng.integers(0, SEQLEN - ln))
        x[i, at:at + ln] = 0 if rng.random() < 0.5 else 3


def plant_str(rng, x, frac):
    units = [(1, 0), (2, 3), (0, 3), (0, 2), (3, 1)]
    idx = rng.choice(len(x), size=int(frac * len(x)), replace=False)
    for i in idx:
        ln = int(rng.integers(30, 81)) & ~1; at = int(rng.integers(0, SEQLEN - ln))
        u = units[int(rng.integers(0, len(units)))]
        x[i, at:at + ln] = np.tile(np.array(u, dtype=np.uint8), ln // 2)


def plant_family(rng, x, frac, flen=150, div=0.10):
    cons = rng.integers(0, 4, size=flen).astype(np.uint8)
    idx = rng.choice(len(x), size=int(frac * len(x)), replace=False)
    for i in idx:
        copy = cons.copy()
        mut = rng.random(flen) < div
        copy[mut] = (copy[mut] + rng.integers(1, 4, size=int(mut.sum()))) & 3
        if rng.random() < 0.5:
            copy = (3 - copy)[::-1]
        at = int(rng.integers(0, SEQLEN - flen))
        x[i, at:at + flen] = copy


def plant_dups(rng, x, frac):
    idx = rng.choice(len(x), size=int(frac * len(x)), replace=False)
    src = rng.integers(0, len(x), size=len(idx))
    x[idx] = x[src]


def workloads(n, seed=4321):
    out = {}
    rng = np.random.default_rng(seed)
    out["uniform"] = background(rng, n)
    out["at_rich"] = background(rng, n, gc=0.41)
    x = background(rng, n); plant_homopolymer(rng, x, 0.10); out["polyA_10"] = x
    x = background(rng, n); plant_str(rng, x, 0.05); out["str_5"] = x
    x = background(rng, n); plant_family(rng, x, 0.20); out["family_20"] = x
    x = background(rng, n); plant_dups(rng, x, 0.05); out["dup_5"] = x
    x = background(rng, n, gc=0.41)
    plant_family(rng, x, 0.20); plant_homopolymer(rng, x, 0.10); plant_str(rng, x, 0.05); plant_dups(rng, x, 0.05)
    out["mixed"] = x
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    kt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    outp = sys.argv[3] if len(sys.argv) > 3 else None
    only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    L, K, D = 11, 7, 3
    capi.load()
    res = {"n": n, "kernel_type": kt, "L": L, "k": K, "d": D, "workloads": {}}
    band = 64
    for name, x in workloads(n).items():
        if only and name not in only:
            continue
        seqs = [ACGT[r].tobytes().decode() for r in x]
        row = {}
        H = {}
        for v in ("index", "auto", "diag"):
            capi.set_option("kernel", v)
            with capi.Problem(kt, L, K, D, 50, 50.0, 1.0) as P:
                P.add_many(seqs)
                P.upload()
                ms = P.bench_lower_resident(3, 1, flush_l2=True)
                st = P.stats()
                row[v] = {"ms_per_pass": float(np.mean(ms)), "M_entries_per_s": n * (n - 1) / 2 / float(np.mean(ms)) / 1e3,
                          "variant_run": int(st["kernel_variant"])}
                if v != "auto":
                    # rows spread over the problem, so that planted sequences are among them
                    hs, ks = [], []
                    for r0 in (n - band, n // 2, n // 7):
                        hs.append(P.hist_block(r0, band, 0, r0).reshape(-1))
                        ks.append(P.kernel_block(r0, band, 0, r0).reshape(-1))
                    H[v] = (np.concatenate(hs), np.concatenate(ks))
        row["hist_equal_index_vs_diag"] = bool(np.array_equal(H["index"][0], H["diag"][0]))
        row["kernel_bits_equal_index_vs_diag"] = bool(np.array_equal(H["index"][1], H["diag"][1]))
        row["hist_sum_per_entry"] = float(H["diag"][0].sum()) / max(1, H["diag"][1].size)
        res["workloads"][name] = row
        print(name, json.dumps(row), flush=True)
    capi.set_option("kernel", "auto")
    if outp:
        with open(outp, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
