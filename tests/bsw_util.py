"""Seeded synthetic extension jobs shared by the BSW tests."""
import numpy as np


def random_jobs(rng, n, qmax, tmax, sim=0.9, nrate=0.01, h0max=150):
    len2 = rng.integers(1, qmax + 1, n).astype(np.int32)
    len1 = np.minimum(np.maximum(len2 + rng.integers(-20, 120, n), 0), tmax).astype(np.int32)
    idq = np.concatenate([[0], np.cumsum(len2[:-1])]).astype(np.int32)
    idr = np.concatenate([[0], np.cumsum(len1[:-1])]).astype(np.int32)
    qer = rng.integers(0, 4, int(len2.sum()), dtype=np.uint8)
    ref = rng.integers(0, 4, int(len1.sum()) + 1, dtype=np.uint8)
    for i in range(n):  # make the target a noisy copy of the query so that extensions go somewhere
        m = min(len1[i], len2[i])
        t = qer[idq[i]:idq[i] + m].copy()
        mut = rng.random(m) > sim
        t[mut] = rng.integers(0, 4, int(mut.sum()))
        if m > 30 and rng.random() < 0.3:  # an indel
            k = int(rng.integers(5, m - 5)); d = int(rng.integers(1, 6))
            t = np.concatenate([t[:k], t[k + d:], rng.integers(0, 4, d, dtype=np.uint8)])
        ref[idr[i]:idr[i] + m] = t
    qer[rng.random(len(qer)) < nrate] = 4
    h0 = rng.integers(1, h0max, n).astype(np.int32)
    return len1, len2, h0, idr, idq, ref, qer
