import numpy as np
def random_seqs(n, length, seed, ragged=False):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(n):
        ln = int(rng.integers(max(20, length // 3), length + 1)) if ragged else length
        out.append(acgt[rng.integers(0, 4, ln)].tobytes().decode())
    return out
