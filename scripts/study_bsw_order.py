#!/usr/bin/env python
"""Offline study (CPU only): lane utilisation of the one-job-per-thread extension kernel under different job orders.
Runs the host build of bsw_col2.cuh on the reference's golden extension jobs with a per-row band trace and evaluates, for
every candidate sort key, the SIMT cost model of a warp (32 consecutive jobs of the sorted class):
    warp cost = sum over rows i of max over live lanes (ROW + ceil(width_i / 2) * PAIR)     [lanes run their row loops in lock step]
    efficiency = sum over lanes of their own cost / (32 * warp cost)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_bsw_col2_cpu as t
import oracle_lib as ol

ROW, PAIR = 90, 25          # instructions per row outside the pair loop / per column pair (SASS of bsw_col2_kernel)
BOUNDS = [32, 64, 96, 128, 160, 256]


def main():
    if len(sys.argv) > 1:      # a ref_driver dump (BM2_DUMP_PREFIX) of any data set: the left-extension calls of the first band try
        import refdump
        groups = refdump.merge_bsw(refdump.read_bsw(sys.argv[1]))
        gd = max(groups, key=lambda x: len(x["h0"]))
        pr = gd["params"]
        g = dict(len1=gd["len1"], len2=gd["len2"], h0=gd["h0"], idr=gd["idr"], idq=gd["idq"], ref=gd["ref"], qer=gd["qer"], w=gd["w"])
        prm = ol.bsw_params(a=int(pr["a"]), b=int(pr["b"]), o_del=int(pr["o_del"]), e_del=int(pr["e_del"]), o_ins=int(pr["o_ins"]),
                            e_ins=int(pr["e_ins"]), zdrop=int(pr["zdrop"]), end_bonus=int(pr["end_bonus"]))
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "bsw_c0.npz"))
        prm = ol.bsw_params(a=int(g["p_a"]), b=int(g["p_b"]), o_del=int(g["p_o_del"]), e_del=int(g["p_e_del"]), o_ins=int(g["p_o_ins"]),
                            e_ins=int(g["p_e_ins"]), zdrop=int(g["p_zdrop"]), end_bonus=int(g["p_end_bonus"]))
    len1, len2, h0 = g["len1"], g["len2"], g["h0"]
    idx = np.nonzero(t._eligible(len1, len2, h0, prm.a))[0]
    n = len(idx)
    i64 = lambda x: np.ascontiguousarray(x, np.int64); i32 = lambda x: np.ascontiguousarray(x, np.int32)
    qoff = i64(g["idq"][idx]); toff = i64(g["idr"][idx]); ql = i32(len2[idx]); tl = i32(len1[idx]); hh = i32(h0[idx])
    ones = i32(np.ones(n))
    p = i32([prm.a, prm.b, prm.o_del, prm.e_del, prm.o_ins, prm.e_ins, prm.zdrop, prm.end_bonus, int(g["w"])])
    out = np.zeros((n, 6), np.int32); rows = np.zeros(n, np.int32)
    cap = int(tl.sum()) + 16
    widths = np.zeros(cap, np.int16)
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    L = t._lib(); L.col2_extend_trace.restype = C.c_longlong
    cells = L.col2_extend_trace(C.c_int(n), P(qoff), P(toff), P(ql), P(tl), P(hh), P(ones), P(ones), P(np.ascontiguousarray(g["qer"])),
                                P(np.ascontiguousarray(g["ref"])), P(p), P(out), P(rows), P(widths), C.c_longlong(cap))
    assert cells > 0
    roff = np.concatenate([[0], np.cumsum(rows)])
    maxr = int(rows.max())
    # per-job row cost matrix (n x maxr), 0 beyond the job's last row
    cost = np.zeros((n, maxr), np.int32)
    for k in range(n):
        w = widths[roff[k]:roff[k + 1]].astype(np.int32)
        cost[k, :len(w)] = ROW + ((w + 1) // 2) * PAIR
    own = cost.sum(1).astype(np.int64)
    cls = np.searchsorted(BOUNDS, ql)            # query-length class
    work = own
    score = out[:, 0]
    keys = {
        "input order": lambda: np.lexsort((np.arange(n), cls)),
        "tlen desc, qlen desc (current)": lambda: np.lexsort((-ql, -tl, cls)),
        "tlen desc, h0 desc": lambda: np.lexsort((-hh, -tl, cls)),
        "h0 desc, tlen desc": lambda: np.lexsort((-tl, -hh, cls)),
        "qlen desc, h0 desc": lambda: np.lexsort((-hh, -ql, cls)),
        "h0+qlen desc, tlen desc": lambda: np.lexsort((-tl, -(hh + ql), cls)),
        "tlen-qlen desc, tlen desc": lambda: np.lexsort((-tl, -(tl - ql), cls)),
        "rows (oracle: true row count) desc": lambda: np.lexsort((-ql, -rows, cls)),
        "work (oracle: true cost) desc": lambda: np.lexsort((-work, cls)),
    }
    print(f"{n} jobs, {cells} cells, mean rows {rows.mean():.1f}, mean width {widths[:roff[-1]].mean():.1f}")
    for name, f in keys.items():
        order = f()
        tot_warp = 0; tot_own = 0
        # warps never straddle a class: pad each class to a multiple of 32
        for c in np.unique(cls):
            o = order[cls[order] == c]
            for s in range(0, len(o), 32):
                blk = o[s:s + 32]
                tot_warp += int(cost[blk].max(0).sum()) * 32
                tot_own += int(own[blk].sum())
        print(f"{name:40s} efficiency {tot_own / tot_warp:.3f}")


if __name__ == "__main__":
    main()
