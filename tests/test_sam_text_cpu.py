"""bm2_sam_format (csrc/sam_text.cpp, host code of the product: the formatting half of mem_aln2sam) on the records of the SAM stage's
device logic (host emulation), against the reference's own SAM text: tests/golden/c0.sam byte for byte, QNAME to the last tag, on 1 and 4
formatter threads.  (The GPU twin, with the records of bm2_sam_pe and reads parsed by bm2_fastq_encode: tests/test_zz_fastq_sam_gpu.py.)"""
import numpy as np
import pytest
import test_oracle_sam_pe as tp


def _to_product_records(capi, e_recs, e_cig, aux, xas, xops):
    recs = np.zeros(len(e_recs), capi.SAM_REC_DT)
    for f in ("read", "flag", "rid", "rnext", "mapq", "nm", "score", "sub", "n_cigar", "n_md", "pos", "pnext", "tlen", "cigar_off", "md_off"):
        recs[f] = e_recs[f]
    recs["alt_sc"] = e_recs["_pad"]; recs["reg"] = aux[:, 0]; recs["is_alt"] = aux[:, 1]; recs["n_mc"] = aux[:, 2]
    xa = np.zeros(len(xas), capi.SAM_XA_DT)
    for f in ("read", "reg", "rid", "is_rev", "nm", "n_cigar", "pos"):
        xa[f] = xas[f]
    xa["cigar_off"] = xas["cigar_off"] + len(e_cig)               # the product keeps the XA operations behind the records' in one array
    return recs, xa, np.concatenate([e_cig, xops]).astype(np.uint32)


def _emul_full(capi, idx, opt, codes, offs, regs, ro, lh, as_):
    """emul_sam_pe with the raw XA entries (test_oracle_sam_pe.emul_sam_pe turns them into strings)."""
    import ctypes as C
    import oracle_lib as ol
    codes = np.ascontiguousarray(codes, np.uint8); offs = np.ascontiguousarray(offs, np.int64)
    regs = np.ascontiguousarray(regs); ro = np.ascontiguousarray(ro, np.int64)
    lh = np.ascontiguousarray(lh, np.int32); as_ = np.ascontiguousarray(as_, np.float64)
    rb = capi.ReadBatch(len(offs) - 1, codes.ctypes.data, offs.ctypes.data)
    rc_ = C.c_void_p(); cg = C.c_void_p(); md = C.c_void_p(); nr = C.c_int64(); no = C.c_int64(); nm = C.c_int64()
    rr = C.c_void_p(); xa = C.c_void_p(); nxa = C.c_int64(); xc = C.c_void_p(); nxc = C.c_int64()
    rc = tp._emul().emul_sam_pe(C.byref(idx.desc), C.byref(opt), C.byref(rb), regs.ctypes.data_as(C.c_void_p), ro.ctypes.data_as(C.c_void_p),
                                lh.ctypes.data_as(C.c_void_p), as_.ctypes.data_as(C.c_void_p), C.c_int64(0), C.byref(rc_), C.byref(nr), C.byref(cg), C.byref(no),
                                C.byref(md), C.byref(nm), C.byref(rr), C.byref(xa), C.byref(nxa), C.byref(xc), C.byref(nxc))
    assert rc == 0

    def arr(p, n, dt):
        dt = np.dtype(dt)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1) * dt.itemsize,))[:n * dt.itemsize].view(dt).copy()
    out = (arr(rc_, nr.value, tp.REC_DT), arr(cg, no.value, "<u4"), arr(md, nm.value, "u1"), arr(rr, 3 * nr.value, "<i4").reshape(-1, 3),
           arr(xa, nxa.value, tp.XA_DT), arr(xc, nxc.value, "<u4"))
    for p in (rc_, cg, md, rr, xa, xc):
        ol.lib().bm2o_free(p)
    return out


@pytest.mark.parametrize("threads", [1, 4])
def test_sam_format_reproduces_the_reference_text(pkg, golden_dir, threads):
    capi = pkg.capi
    import oracle_lib as ol
    idx = capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    names = [l.split()[1] for i, l in enumerate(open(golden_dir + "/c0_index/ref.fa.ann")) if i % 2 == 1]
    opt = capi.default_opt(); opt.flag |= 0x2
    regs, ro, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    assert rc == 0
    pes = capi.pestat(opt, idx.desc.l_pac, regs, ro)
    lh = np.array([v for d in range(4) for v in (pes[d]["low"], pes[d]["high"], pes[d]["failed"])], np.int32)
    as_ = np.array([v for d in range(4) for v in (pes[d]["avg"], pes[d]["std"])], np.float64)
    e_recs, e_cig, e_md, aux, xas, xops = _emul_full(capi, idx, opt, codes, offs, regs, ro, lh, as_)
    recs, xa, cig = _to_product_records(capi, e_recs, e_cig, aux, xas, xops)
    want = [ln for ln in open(golden_dir + "/c0.sam") if not ln.startswith("@")]
    assert len(want) == len(recs)
    qn = {}
    for ln, r in zip(want, recs):
        qn[int(r["read"])] = ln.split("\t", 1)[0]
    read_names = [qn.get(i, "x") for i in range(len(reads))]
    quals = np.full(len(codes), ord(want[0].split("\t")[10][0]), np.uint8)          # the golden reads carry one constant quality character
    got = capi.sam_format(recs, xa, cig, e_md, codes, offs, names, read_names=read_names, quals=quals, n_threads=threads).decode()
    assert got == "".join(want)
    idx.close()


def test_sam_format_rejects_records_that_are_not_grouped_by_read(pkg):
    capi = pkg.capi
    recs = np.zeros(2, capi.SAM_REC_DT); recs["read"] = [1, 0]; recs["rid"] = -1; recs["rnext"] = -1; recs["reg"] = -1; recs["sub"] = -1; recs["score"] = -1
    with pytest.raises(capi.Bm2Error):
        capi.sam_format(recs, np.zeros(0, capi.SAM_XA_DT), np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(20, np.uint8),
                        np.array([0, 10, 20], np.int64), ["c"])
