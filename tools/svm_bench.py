"""gkmQC's evaluate flow after the kernel matrix: 5-fold x 10-repeat C-SVC (scripts/gkmsvm.py:126-160).
GPU consumer (matrix resident on the device) against the reference flow (sklearn SVC fits in a process pool)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkmqc_b200 import capi, driver
import bench
from multiprocessing import Pool
from sklearn.svm import SVC
from sklearn.metrics import roc_auc_score
from sklearn.model_selection import StratifiedKFold

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cpu_fits = int(sys.argv[2]) if len(sys.argv) > 2 else 16
capi.load()
rng = np.random.default_rng(7)
arr = bench.synth(n)
npos = n // 2
motifs = [b"GATAAGGCAT", b"TTGACGTCAA", b"CCCGCCCCTA"]
for i in range(npos):   # plant degenerate motifs in the positives
    for m in motifs[: 1 + i % 3]:
        p = int(rng.integers(0, 290)); mm = bytearray(m)
        if rng.random() < 0.5: mm[int(rng.integers(0, 10))] = b"ACGT"[int(rng.integers(0, 4))]
        arr[i, p:p + 10] = np.frombuffer(bytes(mm), np.uint8)
y = np.concatenate((np.ones(npos, int), np.zeros(n - npos, int)))
splits = []
for _ in range(10):
    kf = StratifiedKFold(n_splits=5, shuffle=True, random_state=None)
    splits.extend(kf.split(np.zeros(n), y))
res = {"n": n, "fits": len(splits)}
with capi.Problem(4, 10, 6, 3, 50, 50.0, 1.0) as P:   # gkmQC's defaults: wgkm, L=10 k=6 d=3
    P.add_many([a.tobytes().decode() for a in arr])
    P.upload()
    for it in range(2):
        t0 = time.perf_counter()
        scores, fits, _ = capi.svm_cv(y, splits, problem=P, C=1.0, eps=1e-3)
        dt = time.perf_counter() - t0
        print("GPU: %d fits on the resident %d x %d matrix: %.3f s (first call computes the matrix)" % (len(splits), n, n, dt), flush=True)
    aucs = [roc_auc_score(y[te], s) for (_, te), s in zip(splits, scores)]
    iters = [f["n_iter"] for f in fits]
    res["gpu_s"] = dt
    res["gpu_auc_mean"] = float(np.mean(aucs)); res["iters_mean"] = float(np.mean(iters)); res["nsv_mean"] = float(np.mean([f["n_sv"] for f in fits]))
    print("AUC %.4f +- %.4f, iterations %.0f (max %d), SVs %.0f" % (np.mean(aucs), np.std(aucs), np.mean(iters), max(iters), res["nsv_mean"]), flush=True)
    K = P.kernel_lower(); K = np.maximum(K, K.T)

def fit(args):
    tr, te = args
    sv = SVC(kernel="precomputed", C=1.0, tol=1e-3, shrinking=False, gamma=1.0, cache_size=1000)
    s = sv.fit(K[tr][:, tr], y[tr]).decision_function(K[te][:, tr])
    return s

ncpu = os.cpu_count() or 1
t0 = time.perf_counter()
with Pool(min(ncpu, cpu_fits)) as pool:
    ref = pool.map(fit, splits[:cpu_fits])
dt_cpu = time.perf_counter() - t0
res["cpu_fits"] = cpu_fits; res["cpu_s"] = dt_cpu; res["cpu_cores"] = ncpu
res["cpu_s_all_fits_estimate"] = dt_cpu * len(splits) / cpu_fits * max(1.0, cpu_fits / ncpu) / max(1.0, cpu_fits / ncpu)
maxdiff = max(float(np.max(np.abs(a - b))) for a, b in zip(ref, scores[:cpu_fits]))
res["max_abs_diff_decision_values"] = maxdiff
print("CPU (sklearn, %d processes): %d fits in %.1f s -> %d fits ~ %.1f s; max |GPU - sklearn| decision value %.3g"
      % (min(ncpu, cpu_fits), cpu_fits, dt_cpu, len(splits), dt_cpu * len(splits) / cpu_fits, maxdiff), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/svm_bench_%d.json" % n, "w"), indent=1)
