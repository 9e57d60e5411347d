#!/usr/bin/env python
"""bench.py -- gkm kernel entries/s (300 bp, L=11 k=7 d=3) on N B200s, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); torch.distributed is used for the
barrier and the max-over-ranks only -- the path has no collective.  A "step" is one full pass of the
hot path: the strict lower triangle of the kernel matrix of the workload below.

Workload (BASELINE.json configs[1]): 5 000 + 5 000 synthetic 300-bp sequences, full 10k x 10k kernel,
kernel type 2 (EST_TRUNC), L=11 k=7 d=3.  For N > 1 the matrix grows so that the work per GPU stays
fixed (n = 10 000 * sqrt(N): weak scaling) and its chunks of row tiles are sharded over the ranks.

  value  device-resident: packed sequences and the output matrix stay in HBM, CUDA-event time of the pass
  e2e    gkm_main_pywrapper(FASTA files -> caller's numpy rows), wall clock: file parse, H2D, kernels,
         D2H and the copy into the caller's matrix all inside the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, K, D, KTYPE, SEQLEN = 11, 7, 3, 2, 300
BASE_N = 10000
PAIRS_PER_ENTRY = (SEQLEN - L + 1) * 2 * (SEQLEN - L + 1)   # 168 200 L-mer pair comparisons
INT_OPS_PER_PAIR = 5                                        # canonical XOR/SHR/LOP3/POPC/ISETP count


def synth(n, seed=1234):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    return acgt[rng.integers(0, 4, size=(n, SEQLEN))]


def write_problem(tmp, n):
    arr = synth(n)
    half = n // 2
    paths = []
    for name, lo, hi in (("pos.fa", 0, half), ("neg.fa", half, n)):
        p = os.path.join(tmp, name)
        with open(p, "w") as f:
            for i in range(lo, hi):
                f.write(">s%d\n%s\n" % (i, arr[i].tobytes().decode()))
        paths.append(p)
    return paths


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm, reasons = [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if sm:
            top = sorted(sm)[len(sm) // 2:]   # samples under load: the upper half
            out["sm_mhz"] = float(np.median(top))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def dist_setup(world):
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    # NCCL prints its version banner on stdout when the first communicator comes up; stdout carries the
    # one JSON line only, so that banner goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group(backend=backend)
        dev = "cuda" if backend == "nccl" else "cpu"
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t)
        if backend == "nccl":
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return dist


def dist_max(dist, x):
    if dist is None:
        return x
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_sum(dist, x):
    if dist is None:
        return x
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def dist_barrier(dist):
    if dist is not None:
        dist.barrier()


def index_probe_counts(L, d, slot_bytes):
    """probes and distinct 32-byte sectors per query L-mer of the index variant, from the product's own mask list
    (slot_bytes = 8: compact slots of the unit-weight kernel types, 16: slots with weights)"""
    from gkmqc_b200 import capi
    import ctypes
    lib = capi.load()
    lib.gkm_idx_delta_count.restype = ctypes.c_longlong
    lib.gkm_idx_deltas.restype = ctypes.c_longlong
    lib.gkm_idx_deltas.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
    nd = lib.gkm_idx_delta_count(L, d)
    a = np.zeros(nd, dtype=np.uint32)
    assert lib.gkm_idx_deltas(L, d, a.ctypes.data, nd) == nd
    return int(nd), int(len(np.unique((a & 0x0FFFFFFF) >> (2 if slot_bytes == 8 else 1))))   # 32-byte sector = 4 or 2 slots


def measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        return None


def cpu_reference_rate(pos, neg, n, budget_s, threads):
    """the reference's own CPU path (oracle/_ref, unmodified sources) on a stated subsample of rows of the SAME
    problem (tree over all n sequences); falls back to the oracle port when the reference was not built here"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if pyoracle.have_ref():
        h = pyoracle.RefHook(pos, neg, KTYPE, L, K, D)
        try:
            probe = np.linspace(n // 4, n - 1, 2 * threads).astype(np.int32)
            t, _ = h.rows_timed(probe, threads)
            rate = float(probe.sum()) / max(t, 1e-9)
            want = max(rate * budget_s, float(probe.sum()))
            nrows = int(min(n - 1, max(threads, want / (n / 2))))
            rows = np.unique(np.linspace(1, n - 1, nrows).astype(np.int32))
            t, cs = h.rows_timed(rows, threads)
            entries = int(rows.sum())
        finally:
            h.close()
        return dict(value=entries / t, unit="entries/s", cores=threads, kind="reference",
                    sample="%d evenly spaced rows of the %d-sequence problem (%d entries, %.1f s), k-mer tree over all sequences"
                           % (len(rows), n, entries, t))
    o = pyoracle.Oracle(KTYPE, L, K, D)
    o.read_problem(pos, neg)
    rows = np.linspace(1, n - 1, 4).astype(np.int32)
    t0 = time.time()
    o.rect(rows, 256, with_hist=False)
    t = time.time() - t0
    return dict(value=len(rows) * 256 / t, unit="entries/s", cores=os.cpu_count(), kind="port",
                sample="%d rows x 256 columns by the brute-force oracle port (%.1f s)" % (len(rows), t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the number of sequences (debugging only)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n or int(round(BASE_N * np.sqrt(world) / 16.0)) * 16
    total_entries = n * (n - 1) // 2
    config = {"workload": "gkm kernel matrix, %d synthetic 300-bp seqs (%d pos + %d neg), strict lower triangle, "
                          "kernel type 2 (EST_TRUNC) L=11 k=7 d=3" % (n, n // 2, n - n // 2),
              "n_seqs": n, "entries_per_step": total_entries, "lmer_pairs_per_entry": PAIRS_PER_ENTRY,
              "sharding": "chunks of row tiles round-robin over %d rank(s), no collective" % world,
              "l2": "L2 flushed (256 MB memset) before every timed pass; %d MB of output written per pass"
                    % (n * n * 8 // (1 << 20))}

    tmp = tempfile.mkdtemp(prefix="gkmbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    pos, neg = write_problem(tmp, n)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        budget = max(2.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
        vals = []
        for it in range(args.warmup + args.steps):
            r = cpu_reference_rate(pos, neg, n, budget, threads)
            if it >= args.warmup:
                vals.append(r["value"])
        v = float(np.mean(vals))
        r["value"] = v
        per_step_ms = 1e3 * total_entries / v
        print(json.dumps({
            "impl": "reference", "metric": "gkm kernel entries/sec (300bp, l=11 k=7 d=3)", "value": v, "unit": "entries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": config, "cpu_baseline": r,
            "e2e": {"value": v, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm
    from gkmqc_b200 import capi
    lib = capi.load()
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no B200 visible and the product has no CPU fallback: " + capi.last_error())
    ids = (capi.ctypes.c_int * 1)(local_rank)
    if lib.gkmb200_set_devices(ids, 1) != 0:
        raise SystemExit(capi.last_error())
    os.environ["GKM_SHARD"] = "%d/%d" % (rank, world)   # read by gkm_main_pywrapper's problem
    dist = dist_setup(world)

    P = capi.Problem(KTYPE, L, K, D)
    P.read(pos, neg)
    P.set_shard(rank, world)
    P.upload()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # value: resident passes, CUDA events, L2 flushed between passes
    dist_barrier(dist)
    ms = P.bench_lower_resident(args.steps, max(args.warmup, 3), flush_l2=True)
    st = P.stats()
    my_ms = float(ms.sum())
    dist_barrier(dist)
    t_max = dist_max(dist, my_ms)
    launches = int(dist_sum(dist, st["launches"] * args.steps))
    value = total_entries * args.steps / (t_max * 1e-3)
    ms_per_step = t_max / args.steps

    # e2e: the drop-in call with host buffers (fresh, untouched output matrix every step, like gkmsvm.py:75)
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 3))
    walls = []
    h2d = d2h = 0
    for it in range(0 if args.no_e2e else 1 + e2e_steps):
        kmat = np.zeros((n, n))
        dist_barrier(dist)
        t0 = time.perf_counter()
        # nthreads is gkmQC's own default (bin/gkmqc.py:107,162: 1); the library sizes its copy-out threads by the host's cores
        ret, kmat, npos, nneg = capi.main_pywrapper(pos, neg, kernel_type=KTYPE, L=L, k=K, d=D, nthreads=1, verbosity=int(os.environ.get("GKM_BENCH_V", "0")), kmat=kmat)
        t1 = time.perf_counter()
        if ret != 0:
            raise SystemExit("gkm_main_pywrapper failed: " + capi.last_error())
        if it > 0:
            walls.append(dist_max(dist, t1 - t0))
        pst = capi.gkmb200_stats()
        lib.gkmb200_get_stats(None, capi.ctypes.byref(pst))
        h2d, d2h = int(pst.h2d_bytes), int(pst.d2h_bytes)
        del kmat
    if args.no_e2e:
        walls = [float("nan")]
    e2e_value = total_entries / float(np.mean(walls))
    h2d = int(dist_sum(dist, h2d))
    d2h = int(dist_sum(dist, d2h))
    clocks = sampler.finish() if rank == 0 else None
    if dist is not None:   # the last collective is behind us: what follows is rank 0's own reporting
        dist.barrier()
        dist.destroy_process_group()
        dist = None

    if rank != 0:
        return

    # roofline of the dominant kernel, by the variant that actually ran:
    #   index: random 16-byte slot probes of an L2-resident table -> L1/L2 sector rate, measured by an in-run gather
    #          micro-benchmark (neither HBM nor the tensor pipe binds: SURVEY.md 8d)
    #   diag / lmer / mma: integer-ALU issue rate (LOP3 lane-ops/s measured in-run)
    variant = {1: "lmer", 2: "diag", 3: "mma", 4: "index"}.get(st["kernel_variant"], "?")
    peak_lop3 = capi.microbench("lop3")       # 1e9 lane-ops/s on this GPU, measured now
    peak_popc = capi.microbench("popc")
    per_gpu_entries_s = value / world
    int_alu_equiv = per_gpu_entries_s * PAIRS_PER_ENTRY * INT_OPS_PER_PAIR / 1e9
    ncu = None
    tf = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")   # written from the committed ncu --set full capture
    if os.path.exists(tf):
        try:
            ncu = json.load(open(tf))
            if ncu.get("variant", "diag") != variant:
                ncu = None
        except Exception:
            ncu = None
    traffic = ncu.get("dram_bytes_per_launch") if ncu else None
    avg_launch_ms = ms_per_step / max(1, st["launches"])
    if variant == "index":
        peak_gather = capi.microbench("gather16")     # 1e9 random 16-byte gathers (one 32-byte sector each) per second
        slot_bytes = 16 if KTYPE in (4, 5) else 8
        probes, sectors = index_probe_counts(L, D, slot_bytes)
        nq = SEQLEN - L + 1
        my_rows = n / world                            # chunks of equal row counts, round-robin over the ranks
        sector_bytes = my_rows * nq * sectors * 32.0   # algorithmic: distinct sectors the probes of one pass must fetch
        achieved = sector_bytes / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "l2_sector", "achieved": achieved, "peak": peak_gather * 32.0, "unit": "GB/s",
                    "frac": achieved / (peak_gather * 32.0), "traffic": traffic,
                    "note": "index variant: every forward L-mer of a row probes %d %d-byte slots of the L2-resident table (%d distinct "
                            "32-byte sectors); achieved = rows x %d L-mers x sectors x 32 B per pass / time; peak = random 16-byte gathers "
                            "from a 64 MB table measured in this run (x 32 B per sector). MEASURED_PEAKS.json hbm_gbs = %s for scale: "
                            "the probes are served by L2, DRAM traffic is `traffic`" % (probes, slot_bytes, sectors, nq, measured_hbm()),
                    "probes_per_lmer": probes, "sectors_per_lmer": sectors, "peak_gather_gsectors": peak_gather,
                    "int_alu_equivalent": {"achieved_gops": int_alu_equiv, "peak_lop3_gops": peak_lop3, "ratio": int_alu_equiv / peak_lop3,
                                           "note": "what the canonical 5-op XOR/POPC form would need for the same entries/s (SURVEY.md 8d)"},
                    "ncu": ncu, "avg_launch_ms": avg_launch_ms, "kernel_variant": variant}
    else:
        roofline = {"bound": "int_alu", "achieved": int_alu_equiv, "peak": peak_lop3, "unit": "Gop/s", "frac": int_alu_equiv / peak_lop3,
                    "traffic": traffic,
                    "note": "achieved = entries/s x 168200 L-mer pairs x 5 integer ops of the canonical XOR/LOP3/POPC form "
                            "(SURVEY.md 8d); peak = LOP3 lane-ops/s measured in this run (MEASURED_PEAKS.json has no integer "
                            "peak). The bit-sliced kernel needs < 1 op per pair, so frac > 1 is possible; see DESIGN.md",
                    "peak_popc_gops": peak_popc, "ncu": ncu, "avg_launch_ms": avg_launch_ms, "kernel_variant": variant}

    # secondary figure: gkmQC's default weighted kernel, type 4 (wgkm, M=50 H=50), same sequences
    secondary = None
    if world == 1:
        with capi.Problem(4, L, K, D, 50, 50.0, 1.0) as P4:
            P4.read(pos, neg)
            ms4 = P4.bench_lower_resident(2, 1, flush_l2=True)
            secondary = {"kernel_type": 4, "value": total_entries / (float(ms4.mean()) * 1e-3), "unit": "entries/s",
                         "ms_per_step": float(ms4.mean()), "note": "wgkm (EST_TRUNC_PW) on the same workload, resident pass"}

    # second secondary figure: the same problem size on genome-like, NON-uniform input (AT-rich background,
    # a repeat family in 20 % of the sequences, poly-A tracts in 10 %, dinucleotide repeats in 5 %, 5 % exact duplicates)
    nonuniform = None
    if world == 1:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("nonuniform", os.path.join(ROOT, "tools", "nonuniform.py"))
            nu = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(nu)
            x = nu.workloads(n)["mixed"]
            with capi.Problem(KTYPE, L, K, D) as Pn:
                Pn.add_many([nu.ACGT[r].tobytes().decode() for r in x])
                msn = Pn.bench_lower_resident(2, 1, flush_l2=True)
                nonuniform = {"workload": "mixed (tools/nonuniform.py)", "value": total_entries / (float(msn.mean()) * 1e-3), "unit": "entries/s",
                              "ms_per_step": float(msn.mean()), "kernel_variant": {1: "lmer", 2: "diag", 3: "mma", 4: "index"}.get(Pn.stats()["kernel_variant"], "?"),
                              "note": "resident pass, kernel type 2; bit-exact against the bit-sliced kernel in tests/test_gpu_parity.py::test_nonuniform_index_vs_bitsliced"}
        except Exception as e:   # a reporting extra must not take the bench line down
            nonuniform = {"error": str(e)}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(pos, neg, n, 15.0, os.cpu_count() or 1)
        one = cpu_reference_rate(pos, neg, n, 5.0, 1)   # the 1-thread figure beside the all-cores one
        cpu["single_thread"] = {"value": one["value"], "unit": one["unit"], "kind": one["kind"], "sample": one["sample"]}

    print(json.dumps({
        "metric": "gkm kernel entries/sec (300bp, l=11 k=7 d=3)", "value": value, "unit": "entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": config, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "entries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * float(np.mean(walls)), "steps": e2e_steps,
                "call": "gkm_main_pywrapper(FASTA paths, double** rows of a fresh numpy matrix, int[2])"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary, "nonuniform": nonuniform}))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
