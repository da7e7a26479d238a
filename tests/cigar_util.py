"""Shared helpers of the seam-3 (CIGAR / NM / MD) tests: request sets and the binding of `ref_driver cigar`, which calls the
UNMODIFIED reference's bwa_gen_cigar2 (oracle/ref_driver.cpp)."""
import os, struct, subprocess, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def refbin(name="ref_driver"):
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    p = os.path.join(ROOT, "oracle", "_ref", isa, name)
    return p if os.path.exists(p) else None


def make_requests(capi, rng, regs, reg_off, read_len, l_pac, n_extra=300):
    """Requests as mem_reg2aln would issue them for the final regs of a batch (several band limits per alignment), perturbed end
    points (indels at the ends, soft gaps), and the rejected / degenerate cases of bwa_gen_cigar2."""
    rd = np.searchsorted(reg_off, np.arange(len(regs)), side="right") - 1
    parts = []
    for w in (0, 3, 20, 100, 400):
        q = np.zeros(len(regs), capi.CIGAR_REQ_DT)
        q["rb"] = regs["rb"]; q["re"] = regs["re"]; q["read"] = rd; q["qb"] = regs["qb"]; q["qe"] = regs["qe"]; q["w"] = w
        parts.append(q)
    base = parts[3]
    k = rng.integers(0, len(base), n_extra)
    p = base[k].copy()
    p["rb"] += rng.integers(-6, 7, n_extra); p["re"] += rng.integers(-6, 7, n_extra)
    p["qb"] = np.clip(p["qb"] + rng.integers(-4, 5, n_extra), 0, read_len - 1)
    p["qe"] = np.clip(p["qe"] + rng.integers(-4, 5, n_extra), p["qb"] + 1, read_len)
    p["w"] = rng.choice([0, 1, 5, 50, 100, 200], n_extra)
    # keep perturbed intervals on one strand and inside the text unless they are the deliberate rejects below
    bad = (p["rb"] < l_pac) & (p["re"] > l_pac)
    p["re"][bad] = l_pac
    p["rb"] = np.clip(p["rb"], 0, 2 * l_pac - 2); p["re"] = np.clip(p["re"], p["rb"] + 1, 2 * l_pac)
    parts.append(p)
    e = base[:8].copy()
    e["rb"][0] = l_pac - 20; e["re"][0] = l_pac + 20                     # bridges the strands: rejected
    e["re"][1] = e["rb"][1]                                                # empty reference interval: rejected
    e["qe"][2] = e["qb"][2]                                                # empty query: rejected
    e["rb"][3] = 2 * l_pac - 30; e["re"][3] = 2 * l_pac + 10; e["qb"][3] = 0; e["qe"][3] = 40     # beyond the text: rejected
    e["rb"][4] = 0; e["re"][4] = 35; e["qb"][4] = 0; e["qe"][4] = 35; e["w"][4] = 0                # no-gap path at the text start
    e["rb"][5] = 2 * l_pac - 35; e["re"][5] = 2 * l_pac; e["qb"][5] = 5; e["qe"][5] = 40; e["w"][5] = 0   # ... at the text end (reverse)
    e["rb"][6] = 100; e["re"][6] = 101; e["qb"][6] = 0; e["qe"][6] = 60; e["w"][6] = 100           # one reference base
    e["rb"][7] = 200; e["re"][7] = 330; e["qb"][7] = 10; e["qe"][7] = 11; e["w"][7] = 100          # one query base
    parts.append(e)
    return np.concatenate(parts)


def reference_gen_cigar(capi, prefix, codes, offsets, reqs):
    """Runs the reference's bwa_gen_cigar2 on the requests -> (recs, cigar, md) in bm2_gen_cigar's layout."""
    exe = refbin()
    assert exe, "oracle/_ref is not built"
    work = tempfile.mkdtemp(prefix="bm2_cigar_")
    with open(work + "/req.bin", "wb") as f:
        f.write(struct.pack("<q", len(reqs)))
        for r in reqs:
            o = int(offsets[r["read"]])
            q = np.ascontiguousarray(codes[o + r["qb"]:o + r["qe"]], np.uint8) if r["qe"] > r["qb"] else np.zeros(0, np.uint8)
            f.write(struct.pack("<qqii", int(r["rb"]), int(r["re"]), int(r["w"]), int(r["qe"] - r["qb"])))
            f.write(q.tobytes())
    subprocess.check_call([exe, "cigar", prefix, work + "/req.bin", work + "/out.bin"], stderr=subprocess.DEVNULL)
    buf = open(work + "/out.bin", "rb").read()
    recs = np.zeros(len(reqs), capi.CIGAR_REC_DT); ops = []; md = []
    pos = 0; no = 0; nm_ = 0
    for i in range(len(reqs)):
        score, n_cigar, nm, n_md = struct.unpack_from("<iiii", buf, pos); pos += 16
        recs[i] = (score, n_cigar, nm, n_md, no, nm_)
        ops.append(np.frombuffer(buf, "<u4", n_cigar, pos)); pos += 4 * n_cigar; no += n_cigar
        md.append(np.frombuffer(buf, "u1", n_md, pos)); pos += n_md; nm_ += n_md
    assert pos == len(buf)
    return recs, (np.concatenate(ops) if ops else np.zeros(0, "<u4")), (np.concatenate(md) if md else np.zeros(0, "u1"))


def same(a, b):
    """(recs, cigar, md) triples equal?  -> list of the first differing request indices."""
    ra, ca, ma = a; rb_, cb, mb = b
    bad = []
    if len(ra) != len(rb_):
        return [-1]
    for i in range(len(ra)):
        x, y = ra[i], rb_[i]
        ok = all(x[f] == y[f] for f in ("score", "n_cigar", "nm", "n_md"))
        ok = ok and np.array_equal(ca[x["cigar_off"]:x["cigar_off"] + x["n_cigar"]], cb[y["cigar_off"]:y["cigar_off"] + y["n_cigar"]])
        ok = ok and np.array_equal(ma[x["md_off"]:x["md_off"] + x["n_md"]], mb[y["md_off"]:y["md_off"] + y["n_md"]])
        if not ok:
            bad.append(i)
            if len(bad) >= 5:
                break
    return bad
