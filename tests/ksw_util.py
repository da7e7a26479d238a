"""Shared helpers of the mate-rescue local-alignment tests (ksw_align2): request sets and the binding of `ref_driver ksw`,
which calls the UNMODIFIED reference's ksw_align2 (oracle/ref_driver.cpp)."""
import ctypes as C, os, struct, subprocess, tempfile
import numpy as np
import oracle_lib as ol
from cigar_util import refbin

KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000


def mate_xtra(l_ms, a=1, min_seed_len=19):
    """xtra of mem_matesw (src/bwamem_pair.cpp:186-188)."""
    return KSW_XSUBO | KSW_XSTART | (KSW_XBYTE if l_ms * a < 250 else 0) | (min_seed_len * a)


def make_requests(rng, n, qlens=(151,), nrate=0.003):
    """(query, target, xtra) triples shaped like mate rescue: the target is a reference window of 1-4 read lengths that holds a noisy
    copy of the query (substitutions, an indel, sometimes a second weaker copy = score2 / te2), or nothing related."""
    reqs = []
    for i in range(n):
        ql = int(rng.choice(qlens))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        tl = int(rng.integers(max(8, ql // 2), 4 * ql + 40))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        kind = rng.random()
        if kind < 0.8:
            c = q.copy()
            mut = rng.random(ql) < rng.choice([0.0, 0.02, 0.08, 0.2]); c[mut] = rng.integers(0, 4, int(mut.sum()))
            if rng.random() < 0.4 and ql > 30:
                p = int(rng.integers(5, ql - 5)); d = int(rng.integers(1, 9))
                c = np.concatenate([c[:p], c[p + d:]]) if rng.random() < 0.5 else np.concatenate([c[:p], rng.integers(0, 4, d).astype(np.uint8), c[p:]])
            if rng.random() < 0.3:                                   # only a part of the query is there
                a0 = int(rng.integers(0, len(c) // 2 + 1)); c = c[a0:a0 + int(rng.integers(min(20, len(c) - a0), len(c) - a0 + 1))]
            pos = int(rng.integers(-(len(c) // 3) - 1, max(tl - len(c) // 2, 1)))  # may hang over either end of the window
            lo, hi = max(pos, 0), min(pos + len(c), tl)
            if hi > lo:
                t[lo:hi] = c[lo - pos:hi - pos]
            if rng.random() < 0.3 and tl > 2 * ql:                   # a second, weaker copy
                c2 = q[: ql // 2 + int(rng.integers(0, ql // 2))].copy()
                m2 = rng.random(len(c2)) < 0.1; c2[m2] = rng.integers(0, 4, int(m2.sum()))
                p2 = int(rng.integers(0, max(tl - len(c2), 1)))
                t[p2:p2 + len(c2)] = c2[:tl - p2]
        q[rng.random(ql) < nrate] = 4
        reqs.append((q, t, mate_xtra(ql)))
    return reqs


def reference_ksw(reqs):
    exe = refbin()
    assert exe, "oracle/_ref is not built"
    work = tempfile.mkdtemp(prefix="bm2_ksw_")
    with open(work + "/req.bin", "wb") as f:
        f.write(struct.pack("<q", len(reqs)))
        for q, t, x in reqs:
            f.write(struct.pack("<iii", len(q), len(t), x)); f.write(q.tobytes()); f.write(t.tobytes())
    subprocess.check_call([exe, "ksw", work + "/req.bin", work + "/out.bin"], stderr=subprocess.DEVNULL)
    return np.fromfile(work + "/out.bin", "<i4").reshape(-1, 7)


def oracle_ksw(reqs, opt):
    L = ol.lib()
    out = np.zeros((len(reqs), 7), np.int32)
    mat = (C.c_int8 * 25)(*[opt.mat[i] for i in range(25)])
    for i, (q, t, x) in enumerate(reqs):
        q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
        L.bm2o_ksw_align2(C.c_int32(len(q)), q.ctypes.data_as(C.c_void_p), C.c_int32(len(t)), t.ctypes.data_as(C.c_void_p), mat,
                          C.c_int32(opt.o_del), C.c_int32(opt.e_del), C.c_int32(opt.o_ins), C.c_int32(opt.e_ins), C.c_int32(x),
                          out[i].ctypes.data_as(C.c_void_p))
    return out
