"""Paired-end SAM stage of the oracle (groundwork for SURVEY 8f): mate rescue + mem_mark_primary_se + mem_pair + the paired / unpaired MAPQ
logic of mem_sam_pe + the columns of mem_aln2sam, against the committed SAM of the UNMODIFIED reference (tests/golden/c0.sam) and, when
oracle/_ref is built, against live runs with options: FLAG, RNAME, POS, MAPQ, CIGAR, RNEXT, PNEXT, TLEN, NM, MD, AS, XS of every line."""
import ctypes as C, os, struct, subprocess, tempfile
import numpy as np
import pytest
import oracle_lib as ol
import cigar_util as cu

REC_DT = np.dtype([("read", "<i4"), ("flag", "<i4"), ("rid", "<i4"), ("mapq", "<i4"), ("rnext", "<i4"), ("tlen_valid", "<i4"), ("nm", "<i4"), ("score", "<i4"),
                   ("sub", "<i4"), ("n_cigar", "<i4"), ("n_md", "<i4"), ("_pad", "<i4"), ("pos", "<i8"), ("pnext", "<i8"), ("tlen", "<i8"),
                   ("cigar_off", "<i8"), ("md_off", "<i8")])


def oracle_sam_pe(capi, idx, opt, codes, offs, regs, ro, lh, as_):
    codes = np.ascontiguousarray(codes, np.uint8); offs = np.ascontiguousarray(offs, np.int64)
    regs = np.ascontiguousarray(regs); ro = np.ascontiguousarray(ro, np.int64)
    lh = np.ascontiguousarray(lh, np.int32); as_ = np.ascontiguousarray(as_, np.float64)
    rb = capi.ReadBatch(len(offs) - 1, codes.ctypes.data, offs.ctypes.data)
    rc_ = C.c_void_p(); cg = C.c_void_p(); md = C.c_void_p(); nr = C.c_int64(); no = C.c_int64(); nm = C.c_int64()
    L = ol.lib()
    rc = L.bm2o_sam_pe(C.byref(idx.desc), C.byref(opt), C.byref(rb), regs.ctypes.data_as(C.c_void_p), ro.ctypes.data_as(C.c_void_p),
                       lh.ctypes.data_as(C.c_void_p), as_.ctypes.data_as(C.c_void_p), C.c_int64(0), C.byref(rc_), C.byref(nr), C.byref(cg), C.byref(no),
                       C.byref(md), C.byref(nm))
    assert rc == 0
    def arr(p, n, dt):
        dt = np.dtype(dt)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1) * dt.itemsize,))[:n * dt.itemsize].view(dt).copy()
    out = arr(rc_, nr.value, REC_DT), arr(cg, no.value, "<u4"), arr(md, nm.value, "u1")
    for p in (rc_, cg, md):
        L.bm2o_free(p)
    return out


def fields(recs, cigar, md, names):
    out = []
    for r in recs:
        rname = names[r["rid"]] if r["rid"] >= 0 else "*"
        cs = "".join(f"{int(o >> 4)}{'MIDSH'[int(o & 0xf)]}" for o in cigar[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]) or "*"
        rnext = "*" if r["rnext"] < 0 else ("=" if r["rnext"] == r["rid"] else names[r["rnext"]])
        tags = {}
        if r["n_cigar"]:
            tags["NM"] = str(int(r["nm"])); tags["MD"] = bytes(md[r["md_off"]:r["md_off"] + r["n_md"] - 1]).decode()
        if r["score"] >= 0: tags["AS"] = str(int(r["score"]))
        if r["sub"] >= 0: tags["XS"] = str(int(r["sub"]))
        out.append((int(r["read"]), int(r["flag"]), rname, int(r["pos"]), int(r["mapq"]), cs, rnext, int(r["pnext"]), int(r["tlen"]), tags))
    return out


def parse_sam(lines):
    out = []; idx = {}
    for ln in lines:
        if ln.startswith("@"):
            continue
        f = ln.rstrip("\n").split("\t")
        flag = int(f[1])
        read = 2 * int(f[0][1:]) + (1 if flag & 0x80 else 0)
        tags = {t[:2]: t[5:] for t in f[11:] if t[:2] in ("NM", "MD", "AS", "XS")}
        out.append((read, flag, f[2], int(f[3]), int(f[4]), f[5], f[6], int(f[7]), int(f[8]), tags))
    return out


def _compare(got, want):
    assert len(got) == len(want), (len(got), len(want))
    bad = [i for i in range(len(got)) if got[i] != want[i]]
    assert not bad, (len(bad), [(got[i], want[i]) for i in bad[:3]])


@pytest.fixture(scope="module")
def c0(pkg, golden_dir):
    capi = pkg.capi
    idx = capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    names = [l.split()[1] for i, l in enumerate(open(golden_dir + "/c0_index/ref.fa.ann")) if i % 2 == 1]
    yield capi, idx, reads, codes, offs, names
    idx.close()


def _pestat(capi, idx, opt, reads, regs, ro):
    lh = np.zeros(12, np.int32); as_ = np.zeros(8, np.float64)
    regs_c = np.ascontiguousarray(regs); ro_c = np.ascontiguousarray(ro, np.int64)
    ol.lib().bm2o_pestat(C.byref(opt), C.c_int64(idx.desc.l_pac), C.c_int32(len(reads)), regs_c.ctypes.data_as(C.c_void_p), ro_c.ctypes.data_as(C.c_void_p),
                         lh.ctypes.data_as(C.c_void_p), as_.ctypes.data_as(C.c_void_p))
    return lh, as_


def test_paired_end_sam_matches_reference_golden(c0, golden_dir):
    capi, idx, reads, codes, offs, names = c0
    opt = capi.default_opt()
    opt.flag |= 0x2                                   # MEM_F_PE (two input files)
    regs, ro, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    lh, as_ = _pestat(capi, idx, opt, reads, regs, ro)
    recs, cig, md = oracle_sam_pe(capi, idx, opt, codes, offs, regs, ro, lh, as_)
    want = parse_sam(open(golden_dir + "/c0.sam"))
    _compare(fields(recs, cig, md, names), want)
    flags = np.array([w[1] for w in want])
    assert (flags & 0x2).sum() > 800 and (flags & 0x800).sum() >= 3 and (flags & 0x4).sum() >= 5      # proper pairs, supplementary, unmapped


@pytest.mark.parametrize("args", [["-a"], ["-M"], ["-P"], ["-S"], ["-Y", "-T", "40"], ["-U", "9"]], ids=["all", "no_multi", "no_pairing", "no_rescue", "softclip_T40", "U9"])
def test_paired_end_sam_matches_the_live_reference(c0, golden_dir, args):
    if cu.refbin() is None:
        pytest.skip("oracle/_ref not built")
    capi, idx, reads, codes, offs, names = c0
    work = tempfile.mkdtemp(prefix="bm2_pe_")
    for k, name in ((0, "r1.fq"), (1, "r2.fq")):
        with open(os.path.join(work, name), "w") as f:
            for i, r in enumerate(reads[k::2]):
                f.write(f"@p{i}\n{''.join('ACGTN'[c] for c in r)}\n+\n{'I' * len(r)}\n")
    with open(os.path.join(work, "o.sam"), "w") as f:
        subprocess.check_call([cu.refbin(), "mem", "-t", "1", "-K", "100000000"] + args + [golden_dir + "/c0_index/ref.fa", os.path.join(work, "r1.fq"),
                               os.path.join(work, "r2.fq")], stdout=f, stderr=subprocess.DEVNULL)
    opt = capi.default_opt(); opt.flag |= 0x2
    for fl, bit in (("-a", 0x8), ("-M", 0x10), ("-P", 0x4), ("-S", 0x20), ("-Y", 0x200)):
        if fl in args: opt.flag |= bit
    if "-T" in args: opt.T = int(args[args.index("-T") + 1])
    if "-U" in args: opt.pen_unpaired = int(args[args.index("-U") + 1])
    regs, ro, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    lh, as_ = _pestat(capi, idx, opt, reads, regs, ro)
    recs, cig, md = oracle_sam_pe(capi, idx, opt, codes, offs, regs, ro, lh, as_)
    _compare(fields(recs, cig, md, names), parse_sam(open(os.path.join(work, "o.sam"))))
