"""SAM text from the records of the seam-4 entry points (bm2_sam_rec / bm2_sam_xa, or the host emulation's records): what is left
to the caller of bm2_sam_pe / bm2_sam_se - mem_aln2sam's formatting (reference src/bwamem.cpp:1592-1730) without its arithmetic.  Used by the tests to show that
the records are sufficient to reproduce the reference's SAM byte for byte (every column, SEQ / QUAL with hard clips, NM MD MC AS XS SA pa XA)."""
import numpy as np


def format_lines(recs, cigar, md, xa_strings, names, codes, offs, qual_char="I", is_alt=None, n_mc=None):
    """One string per record: the SAM line from the FLAG column on.  recs: structured array with read, flag, rid, pos, mapq, n_cigar,
    cigar_off, rnext, pnext, tlen, nm, n_md, md_off, score, sub and alt_sc (or _pad carrying it); is_alt / n_mc: per-record arrays when
    the record type has no such fields (the emulation's)."""
    has = recs.dtype.names
    alt_sc = recs["alt_sc"] if "alt_sc" in has else recs["_pad"]
    is_alt = recs["is_alt"] if "is_alt" in has else is_alt
    n_mc = recs["n_mc"] if "n_mc" in has else n_mc
    by_read = {}
    for k, r in enumerate(recs):
        by_read.setdefault(int(r["read"]), []).append(k)
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    out = []
    for k, r in enumerate(recs):
        ops = cigar[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]
        mc = cigar[r["cigar_off"] + r["n_cigar"]:r["cigar_off"] + r["n_cigar"] + n_mc[k]]
        flag = int(r["flag"])
        f = [str(flag)]
        if r["rid"] >= 0:
            f += [names[r["rid"]], str(int(r["pos"])), str(int(r["mapq"])), "".join(f"{int(o >> 4)}{'MIDSH'[int(o & 15)]}" for o in ops) or "*"]
        else:
            f += ["*", "0", "0", "*"]
        if r["rnext"] >= 0:
            f += ["=" if r["rnext"] == r["rid"] else names[r["rnext"]], str(int(r["pnext"])), str(int(r["tlen"]))]
        else:
            f += ["*", "0", "0"]
        rd = int(r["read"]); seq = codes[offs[rd]:offs[rd + 1]]
        secondary = lambda q: bool(q["flag"] & 0x100) and q["sub"] < 0           # a true secondary (-a: sub = -1), not a -M supplementary (0x10000 printed as 0x100)
        if secondary(r):
            f += ["*", "*"]
        else:
            qb, qe = 0, len(seq)
            if len(ops) and (ops[0] & 15) == 4: (qb, qe) = (qb, qe - int(ops[0] >> 4)) if flag & 0x10 else (qb + int(ops[0] >> 4), qe)
            if len(ops) and (ops[-1] & 15) == 4: (qb, qe) = (qb + int(ops[-1] >> 4), qe) if flag & 0x10 else (qb, qe - int(ops[-1] >> 4))
            s = seq[qb:qe]
            f += ["".join("ACGTN"[c] for c in (comp[s[::-1]] if flag & 0x10 else s)), qual_char * (qe - qb)]
        if r["n_cigar"]:
            f += [f"NM:i:{int(r['nm'])}", "MD:Z:" + bytes(md[r["md_off"]:r["md_off"] + r["n_md"] - 1]).decode()]
        if len(mc):
            f.append("MC:Z:" + "".join(f"{int(o >> 4)}{'MIDSH'[int(o & 15)]}" for o in mc))
        if r["score"] >= 0: f.append(f"AS:i:{int(r['score'])}")
        if r["sub"] >= 0: f.append(f"XS:i:{int(r['sub'])}")
        if not secondary(r):
            others = [j for j in by_read[rd] if j != k and not secondary(recs[j])]
            if others:
                sa = ""
                for j in others:
                    q = recs[j]; qo = cigar[q["cigar_off"]:q["cigar_off"] + q["n_cigar"]]
                    sa += f"{names[q['rid']]},{int(q['pos'])},{'-' if q['flag'] & 0x10 else '+'}," + \
                          "".join(f"{int(o >> 4)}{'MIDSS'[int(o & 15)]}" for o in qo) + f",{int(q['mapq'])},{int(q['nm'])};"
                f.append("SA:Z:" + sa)
            if alt_sc[k] > 0: f.append("pa:f:%.3f" % (float(r["score"]) / float(alt_sc[k])))
        if xa_strings[k]: f.append("XA:Z:" + xa_strings[k])
        out.append("\t".join(f))
    return out
