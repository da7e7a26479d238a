"""Parsers for the stage dumps written by oracle/ref_driver.cpp (unmodified reference)."""
from __future__ import annotations
import numpy as np

SMEM_DT = np.dtype([("rid", "<u4"), ("m", "<u4"), ("n", "<u4"), ("_pad", "<u4"), ("k", "<i8"), ("l", "<i8"), ("s", "<i8")])
PAIR_DT = np.dtype([("idr", "<i4"), ("idq", "<i4"), ("id", "<i4"), ("len1", "<i4"), ("len2", "<i4"), ("h0", "<i4"),
                    ("seqid", "<i4"), ("regid", "<i4"), ("score", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                    ("qle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
REG_FIELDS = ["qb", "qe", "rid", "score", "truesc", "sub", "alt_sc", "csub", "sub_n", "w", "seedcov", "secondary",
              "secondary_all", "seedlen0", "n_comp", "is_alt"]


def read_bsw(path):
    """-> list of dict(kind,w,params, len1,len2,h0, out[n,6](score,tle,gtle,qle,gscore,max_off), ref, qer, idr, idq)."""
    buf = np.fromfile(path, dtype=np.uint8)
    pos = 0
    recs = []
    while pos < len(buf):
        hdr = buf[pos:pos + 48].view("<i4"); pos += 48
        assert hdr[0] == 0x31575342
        kind, n, w = int(hdr[1]), int(hdr[2]), int(hdr[3])
        params = dict(end_bonus=int(hdr[4]), zdrop=int(hdr[5]), o_del=int(hdr[6]), e_del=int(hdr[7]),
                      o_ins=int(hdr[8]), e_ins=int(hdr[9]), a=int(hdr[10]), b=-int(hdr[11]))
        len1 = np.empty(n, np.int32); len2 = np.empty(n, np.int32); h0 = np.empty(n, np.int32)
        out = np.empty((n, 6), np.int32)
        refs = []; qers = []
        for i in range(n):
            v = buf[pos:pos + 36].view("<i4"); pos += 36
            len1[i], len2[i], h0[i] = v[0], v[1], v[2]
            out[i] = v[3:9]
            refs.append(buf[pos:pos + len1[i]]); pos += int(len1[i])
            qers.append(buf[pos:pos + len2[i]]); pos += int(len2[i])
        idr = np.concatenate([[0], np.cumsum(len1[:-1], dtype=np.int64)]).astype(np.int64) if n else np.zeros(0, np.int64)
        idq = np.concatenate([[0], np.cumsum(len2[:-1], dtype=np.int64)]).astype(np.int64) if n else np.zeros(0, np.int64)
        recs.append(dict(kind=kind, w=w, params=params, len1=len1, len2=len2, h0=h0, out=out,
                         ref=np.concatenate(refs) if n else np.zeros(0, np.uint8),
                         qer=np.concatenate(qers) if n else np.zeros(0, np.uint8), idr=idr, idq=idq))
    return recs


def merge_bsw(recs):
    """Merge records that share (w, params) into flat job arrays: dict keyed by (w, end_bonus)."""
    groups = {}
    for r in recs:
        key = (r["w"], tuple(sorted(r["params"].items())))
        groups.setdefault(key, []).append(r)
    out = []
    for (w, pk), rs in groups.items():
        len1 = np.concatenate([r["len1"] for r in rs]); len2 = np.concatenate([r["len2"] for r in rs])
        d = dict(w=w, params=dict(pk), len1=len1, len2=len2, h0=np.concatenate([r["h0"] for r in rs]),
                 out=np.concatenate([r["out"] for r in rs]), ref=np.concatenate([r["ref"] for r in rs]),
                 qer=np.concatenate([r["qer"] for r in rs]), kind=np.concatenate([np.full(len(r["h0"]), r["kind"], np.int32) for r in rs]))
        d["idr"] = np.concatenate([[0], np.cumsum(len1[:-1], dtype=np.int64)]).astype(np.int64)
        d["idq"] = np.concatenate([[0], np.cumsum(len2[:-1], dtype=np.int64)]).astype(np.int64)
        out.append(d)
    return out


def read_smems(path):
    """-> structured array (global rid) in the reference's post-sortSMEMs order."""
    buf = np.fromfile(path, dtype=np.uint8)
    pos = 0
    parts = []
    while pos < len(buf):
        base, nreads, num = buf[pos:pos + 24].view("<i8"); pos += 24
        a = buf[pos:pos + int(num) * SMEM_DT.itemsize].view(SMEM_DT).copy(); pos += int(num) * SMEM_DT.itemsize
        a["rid"] += np.uint32(base)
        parts.append(a)
    return np.concatenate(parts) if parts else np.zeros(0, SMEM_DT)


def read_chains(path):
    """-> list per read of list of chains: dict(n,rid,w,kept,first,is_alt,seqid,frac_rep,pos,seeds[n,4](rbeg,qbeg,len,score))."""
    buf = np.fromfile(path, dtype=np.uint8)
    pos = 0
    reads = []
    while pos < len(buf):
        nc = int(buf[pos:pos + 4].view("<i4")[0]); pos += 4
        chains = []
        for _ in range(nc):
            hdr = buf[pos:pos + 32].view("<i4"); pos += 32
            frac = float(buf[pos:pos + 4].view("<f4")[0]); pos += 4
            cpos = int(buf[pos:pos + 8].view("<i8")[0]); pos += 8
            n = int(hdr[0])
            seeds = np.empty((n, 4), np.int64)
            raw = buf[pos:pos + 20 * n]; pos += 20 * n
            for k in range(n):
                seeds[k, 0] = raw[20 * k:20 * k + 8].view("<i8")[0]
                seeds[k, 1:] = raw[20 * k + 8:20 * k + 20].view("<i4")
            chains.append(dict(n=n, rid=int(hdr[1]), w=int(hdr[2]), kept=int(hdr[3]), first=int(hdr[4]),
                               is_alt=int(hdr[5]), seqid=int(hdr[6]), frac_rep=frac, pos=cpos, seeds=seeds))
        reads.append(chains)
    return reads


REG_DT = np.dtype([("rb", "<i8"), ("re", "<i8")] + [(f, "<i4") for f in REG_FIELDS] + [("frac_rep", "<f4"), ("hash", "<u8")])


def read_regs(path):
    """-> (regs structured array, read_off[n_reads+1])."""
    buf = np.fromfile(path, dtype=np.uint8)
    pos = 0
    offs = [0]
    parts = []
    isz = REG_DT.itemsize
    while pos < len(buf):
        nr = int(buf[pos:pos + 4].view("<i4")[0]); pos += 4
        parts.append(buf[pos:pos + nr * isz]); pos += nr * isz
        offs.append(offs[-1] + nr)
    allb = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return allb.view(REG_DT).copy(), np.array(offs, np.int64)
