"""ctypes binding of libbm2b200.so (include/bm2_b200.h).  Fails loudly when the library or a CUDA
device is missing: there is no CPU fallback."""
from __future__ import annotations
import ctypes as C, os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbm2b200.so")

PAIR_DT = np.dtype([("idr", "<i4"), ("idq", "<i4"), ("id", "<i4"), ("len1", "<i4"), ("len2", "<i4"), ("h0", "<i4"),
                    ("seqid", "<i4"), ("regid", "<i4"), ("score", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                    ("qle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
SMEM_DT = np.dtype([("rid", "<u4"), ("m", "<u4"), ("n", "<u4"), ("_pad", "<u4"), ("k", "<i8"), ("l", "<i8"), ("s", "<i8")])
SEED_DT = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("chain", "<i4")])
CHAIN_DT = np.dtype([("pos", "<i8"), ("seqid", "<i4"), ("rid", "<i4"), ("n_seeds", "<i4"), ("seed_off", "<i4"),
                     ("w", "<i4"), ("kept", "<i4"), ("first", "<i4"), ("is_alt", "<i4"), ("frac_rep", "<f4"), ("_pad", "<i4")])
REG_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("_p0", "<i4"), ("c", "<u8"),
                   ("score", "<i4"), ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"),
                   ("w", "<i4"), ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"),
                   ("n_comp_is_alt", "<i4"), ("frac_rep", "<f4"), ("_p1", "<i4"), ("hash", "<u8"), ("flg", "<i4"), ("_p2", "<i4")])


class MemOpt(C.Structure):
    _fields_ = [("a", C.c_int), ("b", C.c_int), ("o_del", C.c_int), ("e_del", C.c_int), ("o_ins", C.c_int), ("e_ins", C.c_int),
                ("pen_unpaired", C.c_int), ("pen_clip5", C.c_int), ("pen_clip3", C.c_int), ("w", C.c_int), ("zdrop", C.c_int),
                ("max_mem_intv", C.c_uint64), ("T", C.c_int), ("flag", C.c_int), ("min_seed_len", C.c_int),
                ("min_chain_weight", C.c_int), ("max_chain_extend", C.c_int), ("split_factor", C.c_float),
                ("split_width", C.c_int), ("max_occ", C.c_int), ("max_chain_gap", C.c_int), ("n_threads", C.c_int),
                ("chunk_size", C.c_int64), ("mask_level", C.c_float), ("drop_ratio", C.c_float), ("XA_drop_ratio", C.c_float),
                ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float), ("mapQ_coef_fac", C.c_int), ("max_ins", C.c_int),
                ("max_matesw", C.c_int), ("max_XA_hits", C.c_int), ("max_XA_hits_alt", C.c_int), ("mat", C.c_int8 * 25)]


class IndexDesc(C.Structure):
    _fields_ = [("reference_seq_len", C.c_int64), ("count", C.c_int64 * 5), ("sentinel_index", C.c_int64),
                ("cp_occ", C.c_void_p), ("sa_ms_byte", C.c_void_p), ("sa_ls_word", C.c_void_p), ("ref_string", C.c_void_p),
                ("l_pac", C.c_int64), ("n_seqs", C.c_int32), ("ann_offset", C.c_void_p), ("ann_len", C.c_void_p),
                ("ann_is_alt", C.c_void_p)]


# seam 3 (bm2_gen_cigar): request / record layouts of include/bm2_b200.h
CIGAR_REQ_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("read", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("w", "<i4")])
CIGAR_REC_DT = np.dtype([("score", "<i4"), ("n_cigar", "<i4"), ("nm", "<i4"), ("n_md", "<i4"), ("cigar_off", "<i8"), ("md_off", "<i8")])


# seam 4, first piece (bm2_pestat): mem_pestat_t of the four orientations FF, FR, RF, RR
PESTAT_DT = np.dtype([("low", "<i4"), ("high", "<i4"), ("failed", "<i4"), ("_pad", "<i4"), ("avg", "<f8"), ("std", "<f8")])


# seam 4 (bm2_sam_pe): one record per SAM line, XA entries; layouts of include/bm2_b200.h
SAM_REC_DT = np.dtype([("read", "<i4"), ("flag", "<i4"), ("rid", "<i4"), ("rnext", "<i4"), ("mapq", "<i4"), ("nm", "<i4"), ("score", "<i4"), ("sub", "<i4"),
                       ("alt_sc", "<i4"), ("reg", "<i4"), ("n_cigar", "<i4"), ("n_md", "<i4"), ("is_alt", "<i4"), ("n_mc", "<i4"),
                       ("pos", "<i8"), ("pnext", "<i8"), ("tlen", "<i8"), ("cigar_off", "<i8"), ("md_off", "<i8")])
SAM_XA_DT = np.dtype([("read", "<i4"), ("reg", "<i4"), ("rid", "<i4"), ("is_rev", "<i4"), ("nm", "<i4"), ("n_cigar", "<i4"), ("pos", "<i8"), ("cigar_off", "<i8")])


# finer seam under seam 4 (bm2_ksw_align2): request / result layouts of include/bm2_b200.h
KSW_REQ_DT = np.dtype([("qoff", "<i8"), ("toff", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("xtra", "<i4"), ("_pad", "<i4")])
KSW_RES_DT = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4"), ("_pad", "<i4")])


class SamResult(C.Structure):
    _fields_ = [("n_recs", C.c_int64), ("recs", C.c_void_p), ("n_xa", C.c_int64), ("xa", C.c_void_p), ("n_ops", C.c_int64), ("cigar", C.c_void_p),
                ("n_md", C.c_int64), ("md", C.c_void_p)]


class CigarResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("recs", C.c_void_p), ("n_ops", C.c_int64), ("cigar", C.c_void_p), ("n_md", C.c_int64), ("md", C.c_void_p)]


class SamTextIn(C.Structure):
    _fields_ = [("res", C.c_void_p), ("reads", C.c_void_p), ("names", C.c_void_p), ("quals", C.c_void_p), ("contig_names", C.c_void_p),
                ("name_buf", C.c_char_p * 2), ("name_beg", C.c_void_p), ("name_len", C.c_void_p)]


class FastqBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("d_codes", C.c_void_p), ("d_offsets", C.c_void_p), ("codes", C.c_void_p), ("offsets", C.c_void_p),
                ("quals", C.c_void_p), ("name_beg", C.c_void_p), ("name_len", C.c_void_p)]


def sam_format(recs, xa, cigar, md, codes, offsets, contig_names, read_names=None, quals=None, n_threads=1, name_spans=None) -> bytes:
    """bm2_sam_format: the SAM text of a batch from the records of bm2_sam_pe / bm2_sam_se (one line per record, QNAME to the last tag).
    read_names: list of names, or name_spans = (buf1, buf2 or None, name_beg int64[], name_len int32[]) as bm2_fastq_encode returns them."""
    recs = np.ascontiguousarray(recs, SAM_REC_DT); xa = np.ascontiguousarray(xa, SAM_XA_DT)
    cigar = np.ascontiguousarray(cigar, np.uint32); md = np.ascontiguousarray(md, np.uint8)
    codes = np.ascontiguousarray(codes, np.uint8); offsets = np.ascontiguousarray(offsets, np.int64)
    res = SamResult(len(recs), recs.ctypes.data, len(xa), xa.ctypes.data, len(cigar), cigar.ctypes.data, len(md), md.ctypes.data)
    rb = ReadBatch(len(offsets) - 1, codes.ctypes.data, offsets.ctypes.data)
    cn = (C.c_char_p * len(contig_names))(*[s.encode() for s in contig_names])
    rn = (C.c_char_p * len(read_names))(*[s.encode() if isinstance(s, str) else bytes(s) for s in read_names]) if read_names is not None else None
    q = np.ascontiguousarray(np.frombuffer(quals, np.uint8) if isinstance(quals, (bytes, bytearray)) else quals, np.uint8) if quals is not None else None
    tin = SamTextIn(C.cast(C.byref(res), C.c_void_p), C.cast(C.byref(rb), C.c_void_p), C.cast(rn, C.c_void_p) if rn is not None else None,
                    q.ctypes.data if q is not None else None, C.cast(cn, C.c_void_p))
    if name_spans is not None:
        b1, b2, nbeg, nlen = name_spans
        nbeg = np.ascontiguousarray(nbeg, np.int64); nlen = np.ascontiguousarray(nlen, np.int32)
        tin.name_buf[0] = b1; tin.name_buf[1] = b2
        tin.name_beg = nbeg.ctypes.data; tin.name_len = nlen.ctypes.data
    text = C.c_void_p(); n = C.c_int64()
    f = lib().bm2_sam_format
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rc = f(C.byref(tin), int(n_threads), C.byref(text), C.byref(n))
    if rc:
        raise Bm2Error(f"bm2_sam_format failed ({rc})")
    out = C.string_at(text, n.value)
    lib().bm2_free.argtypes = [C.c_void_p]
    lib().bm2_free(text)
    return out


class ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("codes", C.c_void_p), ("offsets", C.c_void_p)]


class SmemResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("smems", C.c_void_p), ("read_off", C.c_void_p)]


class ChainResult(C.Structure):
    _fields_ = [("n_chains", C.c_int64), ("n_seeds", C.c_int64), ("chains", C.c_void_p), ("seeds", C.c_void_p),
                ("read_off", C.c_void_p)]


class RegResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("regs", C.c_void_p), ("read_off", C.c_void_p)]


EXPORTS = ["bm2_create_sibling", "bm2_fastq_encode", "bm2_sam_format", "bm2_free", "bm2_create_resident", "bm2_gather_probe", "bm2_set_sam_staged", "bm2_last_sam_stats", "bm2_gather64_gbs", "bm2_set_sub_batches", "bm2_seed_chain_extend_resident", "bm2_last_counters", "bm2_set_stream", "bm2_int_pipe_gops", "bm2_abi_version", "bm2_opt_init", "bm2_index_load", "bm2_index_free", "bm2_create", "bm2_destroy",
           "bm2_last_error", "bm2_extend_pairs", "bm2_extend_pairs_device", "bm2_collect_smems", "bm2_seed_chain",
           "bm2_seed_chain_extend", "bm2_last_stage_ms", "bm2_gen_cigar", "bm2_pestat", "bm2_sam_pe", "bm2_sam_se", "bm2_ksw_align2"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.bm2_last_error.restype = C.c_char_p
        _lib.bm2_last_error.argtypes = [C.c_void_p]
        _lib.bm2_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]
        _lib.bm2_destroy.argtypes = [C.c_void_p]
        _lib.bm2_index_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(IndexDesc))]
        _lib.bm2_index_free.argtypes = [C.POINTER(IndexDesc)]
        _lib.bm2_extend_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        _lib.bm2_extend_pairs_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_void_p]
        for f in ("bm2_collect_smems", "bm2_seed_chain", "bm2_seed_chain_extend"):
            if hasattr(_lib, f):
                getattr(_lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def default_opt() -> MemOpt:
    o = MemOpt()
    lib().bm2_opt_init(C.byref(o))
    return o


class Bm2Error(RuntimeError):
    pass


def pestat(opt, l_pac, regs, read_off):
    """bm2_pestat: insert-size statistics of a chunk (reads 2i, 2i+1 are mates) from the regs of bm2_seed_chain_extend -> PESTAT_DT[4]."""
    regs = np.ascontiguousarray(regs, REG_DT); read_off = np.ascontiguousarray(read_off, np.int64)
    out = np.zeros(4, PESTAT_DT)
    f = lib().bm2_pestat
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = f(C.addressof(opt), int(l_pac), len(read_off) - 1, regs.ctypes.data, read_off.ctypes.data, out.ctypes.data)
    if rc:
        raise Bm2Error(f"bm2_pestat failed ({rc})")
    return out


class Index:
    """Host-resident index loaded by the native loader (bm2_index_load)."""

    def __init__(self, prefix: str):
        self._p = C.POINTER(IndexDesc)()
        rc = lib().bm2_index_load(prefix.encode(), C.byref(self._p))
        if rc:
            lib().bm2_index_io_error.restype = C.c_char_p
            raise Bm2Error(f"bm2_index_load({prefix}): {lib().bm2_index_io_error().decode()}")
        self.desc = self._p.contents

    def close(self):
        if self._p:
            lib().bm2_index_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """bm2_ctx wrapper; mirrors the reference seams (see include/bm2_b200.h)."""

    def __init__(self, device: int = 0, index=None, opt: MemOpt | None = None, resident: bool = False):
        """index: an Index (host arrays, uploaded by bm2_create) or an IndexDesc; resident=True: the four big arrays of the
        descriptor are device pointers already in `device`'s memory (bm2_create_resident; the caller keeps them alive)."""
        self._ctx = C.c_void_p()
        self.opt = opt if opt is not None else default_opt()
        self._index = index
        idx_ptr = None
        if index is not None:
            idx_ptr = C.cast(C.byref(index.desc if isinstance(index, Index) else index), C.c_void_p)
        f = lib().bm2_create_resident if resident else lib().bm2_create
        rc = f(C.byref(self._ctx), device, idx_ptr, C.cast(C.byref(self.opt), C.c_void_p))
        if rc:
            raise Bm2Error(("bm2_create_resident: " if resident else "bm2_create: ") + lib().bm2_last_error(None).decode())

    def close(self):
        if self._ctx:
            lib().bm2_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise Bm2Error(f"{what}: " + lib().bm2_last_error(self._ctx).decode())

    # seam 1 (BandedPairWiseSW::getScores16 & co.)
    def extend_pairs(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int, end_bonus: int):
        assert pairs.dtype == PAIR_DT and pairs.flags.c_contiguous
        ref = np.ascontiguousarray(ref, np.uint8); qer = np.ascontiguousarray(qer, np.uint8)
        self._check(lib().bm2_extend_pairs(self._ctx, pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data,
                                           len(pairs), w, end_bonus), "bm2_extend_pairs")
        return pairs

    def extend_pairs_device(self, d_pairs_ptr, d_ref_ptr, d_qer_ptr, n, w, end_bonus, d_cells_ptr=None):
        self._check(lib().bm2_extend_pairs_device(self._ctx, d_pairs_ptr, d_ref_ptr, d_qer_ptr, n, w, end_bonus,
                                                  d_cells_ptr), "bm2_extend_pairs_device")

    # seam 2
    @staticmethod
    def _batch(codes: np.ndarray, offsets: np.ndarray):
        codes = np.ascontiguousarray(codes, np.uint8); offsets = np.ascontiguousarray(offsets, np.int64)
        rb = ReadBatch(len(offsets) - 1, codes.ctypes.data, offsets.ctypes.data)
        return rb, (codes, offsets)

    def collect_smems(self, codes, offsets):
        rb, keep = self._batch(codes, offsets)
        res = SmemResult()
        self._check(lib().bm2_collect_smems(self._ctx, C.byref(rb), C.byref(res)), "bm2_collect_smems")
        n = res.n
        sm = np.ctypeslib.as_array(C.cast(res.smems, C.POINTER(C.c_uint8)), shape=(n * SMEM_DT.itemsize,)).view(SMEM_DT).copy() if n else np.zeros(0, SMEM_DT)
        off = np.ctypeslib.as_array(C.cast(res.read_off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,)).copy()
        return sm, off

    def seed_chain(self, codes, offsets):
        rb, keep = self._batch(codes, offsets)
        res = ChainResult()
        self._check(lib().bm2_seed_chain(self._ctx, C.byref(rb), C.byref(res)), "bm2_seed_chain")
        nc, ns = res.n_chains, res.n_seeds
        ch = np.ctypeslib.as_array(C.cast(res.chains, C.POINTER(C.c_uint8)), shape=(nc * CHAIN_DT.itemsize,)).view(CHAIN_DT).copy() if nc else np.zeros(0, CHAIN_DT)
        sd = np.ctypeslib.as_array(C.cast(res.seeds, C.POINTER(C.c_uint8)), shape=(ns * SEED_DT.itemsize,)).view(SEED_DT).copy() if ns else np.zeros(0, SEED_DT)
        off = np.ctypeslib.as_array(C.cast(res.read_off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,)).copy()
        return ch, sd, off

    def seed_chain_extend(self, codes, offsets, copy=True):
        rb, keep = self._batch(codes, offsets)
        res = RegResult()
        self._check(lib().bm2_seed_chain_extend(self._ctx, C.byref(rb), C.byref(res)), "bm2_seed_chain_extend")
        n = res.n
        regs = np.ctypeslib.as_array(C.cast(res.regs, C.POINTER(C.c_uint8)), shape=(n * REG_DT.itemsize,)).view(REG_DT) if n else np.zeros(0, REG_DT)
        off = np.ctypeslib.as_array(C.cast(res.read_off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,))
        return (regs.copy(), off.copy()) if copy else (regs, off)

    def gen_cigar(self, codes, offsets, reqs):
        """bm2_gen_cigar: reqs is a CIGAR_REQ_DT array -> (recs CIGAR_REC_DT, cigar uint32[], md bytes)."""
        rb, keep = self._batch(codes, offsets)
        reqs = np.ascontiguousarray(reqs, CIGAR_REQ_DT)
        res = CigarResult()
        lib().bm2_gen_cigar.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        self._check(lib().bm2_gen_cigar(self._ctx, C.byref(rb), reqs.ctypes.data_as(C.c_void_p), len(reqs), C.byref(res)), "bm2_gen_cigar")
        def arr(p, n, dt):
            dt = np.dtype(dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).view(dt).copy() if n else np.zeros(0, dt)
        return arr(res.recs, res.n, CIGAR_REC_DT), arr(res.cigar, res.n_ops, "<u4"), arr(res.md, res.n_md, "u1")

    def ksw_align2(self, reqs):
        """bm2_ksw_align2: reqs = [(query codes, window codes, xtra), ...] -> int32[n, 7] (score, te, qe, score2, te2, tb, qb)."""
        seqs = np.concatenate([np.concatenate([np.asarray(q, np.uint8), np.asarray(t, np.uint8)]) for q, t, _ in reqs]) if reqs else np.zeros(0, np.uint8)
        rq = np.zeros(len(reqs), KSW_REQ_DT); pos = 0
        for i, (q, t, x) in enumerate(reqs):
            rq[i] = (pos, pos + len(q), len(q), len(t), x, 0); pos += len(q) + len(t)
        out = np.zeros(len(reqs), KSW_RES_DT)
        f = lib().bm2_ksw_align2
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        self._check(f(self._ctx, seqs.ctypes.data_as(C.c_void_p), len(seqs), rq.ctypes.data_as(C.c_void_p), len(rq), out.ctypes.data_as(C.c_void_p)), "bm2_ksw_align2")
        return np.stack([out[k] for k in ("score", "te", "qe", "score2", "te2", "tb", "qb")], axis=1).astype(np.int32) if len(out) else np.zeros((0, 7), np.int32)

    def sam_se(self, codes, offsets, regs, read_off, id_base=0):
        """bm2_sam_se: the SAM stage of a batch of single-end reads -> (recs SAM_REC_DT, xa SAM_XA_DT, cigar uint32[], md bytes)."""
        rb, keep = self._batch(codes, offsets)
        regs = np.ascontiguousarray(regs, REG_DT); read_off = np.ascontiguousarray(read_off, np.int64)
        res = SamResult()
        f = lib().bm2_sam_se
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        self._check(f(self._ctx, C.byref(rb), regs.ctypes.data_as(C.c_void_p), read_off.ctypes.data_as(C.c_void_p), int(id_base), C.byref(res)), "bm2_sam_se")
        def arr(p, n, dt):
            dt = np.dtype(dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).view(dt).copy() if n else np.zeros(0, dt)
        return arr(res.recs, res.n_recs, SAM_REC_DT), arr(res.xa, res.n_xa, SAM_XA_DT), arr(res.cigar, res.n_ops, "<u4"), arr(res.md, res.n_md, "u1")

    def sam_pe(self, codes, offsets, regs, read_off, pes, id_base=0):
        """bm2_sam_pe: the SAM stage of a batch of pairs -> (recs SAM_REC_DT, xa SAM_XA_DT, cigar uint32[], md bytes)."""
        rb, keep = self._batch(codes, offsets)
        regs = np.ascontiguousarray(regs, REG_DT); read_off = np.ascontiguousarray(read_off, np.int64); pes = np.ascontiguousarray(pes, PESTAT_DT)
        res = SamResult()
        f = lib().bm2_sam_pe
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        self._check(f(self._ctx, C.byref(rb), regs.ctypes.data_as(C.c_void_p), read_off.ctypes.data_as(C.c_void_p), pes.ctypes.data_as(C.c_void_p),
                      int(id_base), C.byref(res)), "bm2_sam_pe")
        def arr(p, n, dt):
            dt = np.dtype(dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).view(dt).copy() if n else np.zeros(0, dt)
        return arr(res.recs, res.n_recs, SAM_REC_DT), arr(res.xa, res.n_xa, SAM_XA_DT), arr(res.cigar, res.n_ops, "<u4"), arr(res.md, res.n_md, "u1")

    def fastq_encode(self, buf1: bytes, buf2: bytes | None = None, want_names: bool = True):
        """bm2_fastq_encode: raw FASTQ bytes of a chunk (two files for pairs) -> dict(n_reads, codes, offsets, quals, names, d_codes, d_offsets);
        the device pointers stay valid until the context's next call."""
        b = FastqBatch()
        f = lib().bm2_fastq_encode
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p]
        self._check(f(self._ctx, buf1, len(buf1), buf2, len(buf2) if buf2 is not None else 0, C.byref(b)), "bm2_fastq_encode")
        n = b.n_reads
        offs = np.ctypeslib.as_array(C.cast(b.offsets, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        tot = int(offs[-1])
        codes = np.ctypeslib.as_array(C.cast(b.codes, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
        quals = np.ctypeslib.as_array(C.cast(b.quals, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
        nb = np.ctypeslib.as_array(C.cast(b.name_beg, C.POINTER(C.c_int64)), shape=(max(n, 1),))[:n].copy()
        nl = np.ctypeslib.as_array(C.cast(b.name_len, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        bufs = (buf1, buf2 if buf2 is not None else buf1)
        stride = 2 if buf2 is not None else 1
        names = [bufs[r % stride][nb[r]:nb[r] + nl[r]] for r in range(n)] if want_names else None
        return dict(n_reads=n, codes=codes, offsets=offs, quals=quals, names=names, d_codes=b.d_codes, d_offsets=b.d_offsets,
                    name_spans=(buf1, buf2, nb, nl))

    def set_sam_staged(self, on: int):
        """bm2_set_sam_staged: 1 / 2 = the rescue's local alignments as a batch (one window per warp / per thread) before the per-pair kernel, 0 = inside it."""
        lib().bm2_set_sam_staged.argtypes = [C.c_void_p, C.c_int]
        self._check(lib().bm2_set_sam_staged(self._ctx, int(on)), "bm2_set_sam_staged")

    def last_sam_stats(self):
        """bm2_last_sam_stats -> dict: device ms of the last bm2_sam_pe / bm2_sam_se call (jobs, ksw, pairs, gather) and the rescue counters."""
        ms = (C.c_double * 4)(); cnt = (C.c_ulonglong * 6)()
        lib().bm2_last_sam_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self._check(lib().bm2_last_sam_stats(self._ctx, ms, cnt, 4, 6), "bm2_last_sam_stats")
        return {"ms": {"jobs": ms[0], "ksw": ms[1], "pairs": ms[2], "gather_and_copies": ms[3]}, "staged": int(cnt[0]), "jobs": int(cnt[1]),
                "looked_up": int(cnt[2]), "in_place": int(cnt[3]), "window_moved": int(cnt[4]), "waves": int(cnt[5])}

    def set_stream(self, cuda_stream_handle):
        lib().bm2_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        self._check(lib().bm2_set_stream(self._ctx, cuda_stream_handle), "bm2_set_stream")

    def gather64_gbs(self, span_bytes: int = 0) -> float:
        v = C.c_double()
        lib().bm2_gather64_gbs.argtypes = [C.c_void_p, C.c_ulonglong, C.POINTER(C.c_double)]
        self._check(lib().bm2_gather64_gbs(self._ctx, int(span_bytes), C.byref(v)), "bm2_gather64_gbs")
        return v.value

    def gather_probe(self, span_bytes: int = 0, mlp: int = 4, shape: int = 0) -> float:
        """bm2_gather_probe: GB/s of random requests over the Occ table (shape 0: 64 B as 4 x 16 B, 1: 32 B as one 256-bit load, 2: 64 B as two)."""
        v = C.c_double()
        lib().bm2_gather_probe.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_int, C.POINTER(C.c_double)]
        self._check(lib().bm2_gather_probe(self._ctx, int(span_bytes), int(mlp), int(shape), C.byref(v)), "bm2_gather_probe")
        return v.value

    def set_sub_batches(self, k: int, min_reads: int = 16384):
        """Seam 2 runs a batch as k sub-batches in flight (bm2_set_sub_batches); k = 1 turns the split off."""
        lib().bm2_set_sub_batches.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self._check(lib().bm2_set_sub_batches(self._ctx, int(k), int(min_reads)), "bm2_set_sub_batches")

    def int_pipe_gops(self) -> float:
        v = C.c_double()
        lib().bm2_int_pipe_gops.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._check(lib().bm2_int_pipe_gops(self._ctx, C.byref(v)), "bm2_int_pipe_gops")
        return v.value

    def seed_chain_extend_resident(self, codes, offsets, d_codes_ptr, d_offsets_ptr, copy_out=False, return_arrays=False):
        rb, keep = self._batch(codes, offsets)
        res = RegResult()
        lib().bm2_seed_chain_extend_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self._check(lib().bm2_seed_chain_extend_resident(self._ctx, C.byref(rb), d_codes_ptr, d_offsets_ptr, int(copy_out), C.byref(res)),
                    "bm2_seed_chain_extend_resident")
        if not return_arrays:
            return res.n
        n = res.n
        regs = np.ctypeslib.as_array(C.cast(res.regs, C.POINTER(C.c_uint8)), shape=(n * REG_DT.itemsize,)).view(REG_DT).copy() if n else np.zeros(0, REG_DT)
        off = np.ctypeslib.as_array(C.cast(res.read_off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,)).copy()
        return regs, off

    def counters(self):
        v = (C.c_ulonglong * 5)()
        lib().bm2_last_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib().bm2_last_counters(self._ctx, v, 5)
        return dict(n_ext=v[0], n_lf=v[1], cells=v[2], retry_left=v[3], retry_right=v[4])

    def stage_ms(self):
        names = C.POINTER(C.c_char_p)(); ms = C.POINTER(C.c_float)(); n = C.c_int()
        lib().bm2_last_stage_ms(self._ctx, C.byref(names), C.byref(ms), C.byref(n))
        return {names[i].decode(): ms[i] for i in range(n.value)}
