/* bm2_b200.h — C ABI of libbm2b200.so: the B200-native seed-and-extend hot path of bwa-mem2.
 *
 * Plain C, pointers and sizes only.  Every entry point names the reference interface it replaces
 * (file:line under bwa-mem2 @ 97978f95).  The reference has no FFI; the seam is the C++ function
 * boundary below `mem_process_seqs` (src/bwamem.cpp:1338): `kt_for(worker_bwt)` + `kt_for(worker_aln)`
 * (src/bwamem.cpp:1359,1363), and the finer seams `BandedPairWiseSW::getScores16/getScores8/
 * scalarBandedSWAWrapper` (src/bandedSWA.h:130-297) and `FMI_search::getSMEMs*`/
 * `get_sa_entries_prefetch` (src/FMI_search.h:106-165).  INTEGRATION.md shows the reference-side
 * binding.  All functions return 0 on success, non-zero on error (`bm2_last_error`); there is no
 * CPU fallback: without a usable CUDA device every compute entry fails.
 */
#ifndef BM2_B200_H
#define BM2_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BM2_ABI_VERSION 1

typedef struct bm2_ctx bm2_ctx;

/* Field-for-field mirror of `mem_opt_t` (src/bwamem.h:76-108) so a `const mem_opt_t*` can be
 * passed as `const bm2_mem_opt_t*`; defaults: src/bwamem.cpp:107-143. */
typedef struct bm2_mem_opt_t {
    int a, b;
    int o_del, e_del;
    int o_ins, e_ins;
    int pen_unpaired;
    int pen_clip5, pen_clip3;
    int w;
    int zdrop;
    uint64_t max_mem_intv;
    int T;
    int flag;
    int min_seed_len;
    int min_chain_weight;
    int max_chain_extend;
    float split_factor;
    int split_width;
    int max_occ;
    int max_chain_gap;
    int n_threads;
    int64_t chunk_size;
    float mask_level;
    float drop_ratio;
    float XA_drop_ratio;
    float mask_level_redun;
    float mapQ_coef_len;
    int mapQ_coef_fac;
    int max_ins;
    int max_matesw;
    int max_XA_hits, max_XA_hits_alt;
    int8_t mat[25];
} bm2_mem_opt_t;

/* `mem_opt_init()` defaults (src/bwamem.cpp:107-143) incl. `bwa_fill_scmat` (src/bwa.cpp:248). */
void bm2_opt_init(bm2_mem_opt_t *opt);

/* One checkpoint of the 2bit.64 Occ table: `CP_OCC` (src/FMI_search.h:54-58). */
typedef struct bm2_cp_occ {
    int64_t  cp_count[4];
    uint64_t one_hot_bwt_str[4];
} bm2_cp_occ;

/* Host view of the index the hot path reads (src/FMI_search.h:167-177, src/bntseq.h:53-61,
 * `ref_string` src/fastmap.cpp:860-881).  All pointers are HOST memory owned by the caller;
 * `bm2_create` copies them to HBM. */
typedef struct bm2_index_desc {
    int64_t reference_seq_len;      /* N = 2*l_pac + 1 (BWT length incl. sentinel)              */
    int64_t count[5];               /* as held in memory after load: (#symbols < b) + 1         */
    int64_t sentinel_index;
    const bm2_cp_occ *cp_occ;       /* (N >> 6) + 1 entries                                     */
    const int8_t   *sa_ms_byte;     /* (N >> 3) + 1 entries: bits 32..39 of sampled SA          */
    const uint32_t *sa_ls_word;     /* (N >> 3) + 1 entries: bits 0..31                         */
    const uint8_t  *ref_string;     /* 2*l_pac codes 0..3: forward then reverse complement      */
    int64_t l_pac;
    int32_t n_seqs;                 /* contigs (bntseq_t::n_seqs)                               */
    const int64_t *ann_offset;      /* n_seqs (bntann1_t::offset)                               */
    const int32_t *ann_len;         /* n_seqs (bntann1_t::len)                                  */
    const int32_t *ann_is_alt;      /* n_seqs, may be NULL (bntann1_t::is_alt)                  */
} bm2_index_desc;

/* Native loader of `<prefix>.bwt.2bit.64`, `.0123`, `.ann` (+ `.alt`): replaces
 * FMI_search::load_index (src/FMI_search.cpp:384-494), bwa_idx_load_ele (read_index_ele.cpp:60)
 * and the ref_string read (fastmap.cpp:860-881).  The returned descriptor owns its host memory. */
int  bm2_index_load(const char *prefix, bm2_index_desc **out);
void bm2_index_free(bm2_index_desc *idx);

/* Create a device context on CUDA device `device`: uploads the index to HBM and fixes the
 * parameters.  `idx` may be NULL for a BSW-only context (bm2_extend_pairs). */
int  bm2_create(bm2_ctx **out, int device, const bm2_index_desc *idx, const bm2_mem_opt_t *opt);
/* The same for an index that is ALREADY in the memory of `device` (multi-GPU start-up, SURVEY 8e: one rank reads the index files, the
 * others receive the four big arrays by one NCCL broadcast over NVLink instead of 8 x 16 GB of disk + PCIe traffic): `cp_occ`,
 * `sa_ms_byte`, `sa_ls_word` and `ref_string` of `dev_idx` are DEVICE pointers (file layout, sizes as in bm2_index_desc), the small
 * `ann_*` arrays HOST pointers.  The context does not own the four arrays: the caller keeps them alive until bm2_destroy and must not
 * read `cp_occ` afterwards - it is permuted in place into the device layout (fm_device.cuh) unless BM2_OCC_LAYOUT=0. */
int  bm2_create_resident(bm2_ctx **out, int device, const bm2_index_desc *dev_idx, const bm2_mem_opt_t *opt);
/* A second context on the same device that uses `ctx`'s index in place (no second copy in HBM), with its own streams and buffers and a copy of
 * `ctx`'s parameters: one context per host worker thread, the way the reference runs two chunks at a time in kt_pipeline (src/fastmap.cpp:
 * 952-1003, src/kthread.cpp:122-176) - a context serves one call at a time, two contexts serve two.  `ctx` must outlive the sibling. */
int  bm2_create_sibling(bm2_ctx **out, bm2_ctx *ctx);
void bm2_destroy(bm2_ctx *ctx);
const char *bm2_last_error(const bm2_ctx *ctx);   /* ctx may be NULL: last create error */
/* Launch on a caller-owned CUDA stream (cudaStream_t as void*), e.g. the caller's framework stream,
 * so that the caller's events bracket the kernels.  NULL restores the context's own stream. */
int  bm2_set_stream(bm2_ctx *ctx, void *cuda_stream);
/* Seam 2 runs a batch as `k` sub-batches in flight (own CUDA streams and scratch, shared index): the SMEM stage is
 * bound by memory latency and the extension stage by the integer pipe, so sub-batches at different stages fill
 * each other's stalls (measured +17 % reads/s at k = 4 on B200).  A batch is only split when every sub-batch gets
 * at least `min_reads` reads; cuts are multiples of 512 reads so that results do not depend on k (the reference's
 * kt_for works in 512-read blocks, src/kthread.cpp:41-115).  Defaults: k = 4, min_reads = 16384; k = 1 turns it off.
 * The mem_collect_smem / mem_kernel1_core stage entries (bm2_collect_smems, bm2_seed_chain) always run unsplit. */
int  bm2_set_sub_batches(bm2_ctx *ctx, int k, int min_reads);
/* Measured integer-pipe throughput of this device (G lane-ops/s of dependent 32-bit add/max
 * chains over all SMs): the denominator of the BSW cell-update roofline. */
int  bm2_int_pipe_gops(bm2_ctx *ctx, double *gops_s32);
/* Measured throughput (GB/s) of independent random 64-byte reads over the first `span_bytes` (0 = all) of this context's
 * Occ checkpoint table: what the memory system delivers for the SMEM stage's access shape (two random 64-B checkpoints
 * per interval extension) when no dependent address chain limits it.  Reported next to the HBM copy peak in bench.py. */
int  bm2_gather64_gbs(bm2_ctx *ctx, unsigned long long span_bytes, double *gbs);
/* The same probe with a selectable request shape and memory-level parallelism: shape 0 = 64 B as four 16-B loads of one thread,
 * 1 = 32 B as ONE 256-bit load (the half-checkpoint of the device Occ layout), 2 = 64 B as two 256-bit loads, 3 = 32 B by cp.async.bulk
 * into shared memory behind an mbarrier (the TMA path; at most 4 in flight per thread), 4 = shape 3 and shape 1 together (mlp of each); `mlp` (1, 2, 4, 8)
 * independent requests in flight per thread.  GB/s of requested bytes.  Decides whether the SMEM stage is bound by DRAM, by the
 * load/store unit's request rate or by latency. */
int  bm2_gather_probe(bm2_ctx *ctx, unsigned long long span_bytes, int mlp, int shape, double *gbs);
int  bm2_abi_version(void);

/* ---- seam 1: batched banded-SW seed extension -------------------------------------------------
 * Layout-compatible with `SeqPair` (src/bandedSWA.h:90-99). */
typedef struct bm2_seqpair {
    int32_t idr, idq, id;
    int32_t len1, len2;
    int32_t h0;
    int32_t seqid, regid;
    int32_t score, tle, gtle, qle;
    int32_t gscore, max_off;
} bm2_seqpair;

/* Replaces BandedPairWiseSW::getScores16 / getScores8 / scalarBandedSWAWrapper
 * (src/bandedSWA.cpp:2664, :1970, :242; AVX2 twins :1117, :412): for pair i, target = seq_buf_ref + idr (len1 codes 0..4),
 * query = seq_buf_qer + idq (len2), start score h0, band w; writes score/tle/gtle/qle/gscore/
 * max_off in place.  `end_bonus` is the constructor's end_bonus (pen_clip5 for left, pen_clip3 for
 * right extensions, src/bwamem.cpp:2456-2462).  Host buffers; copies are done inside. */
int bm2_extend_pairs(bm2_ctx *ctx, bm2_seqpair *pairs, const uint8_t *seq_buf_ref,
                     const uint8_t *seq_buf_qer, int32_t n_pairs, int32_t w, int32_t end_bonus);

/* Device-resident variant used by the throughput bench: same contract, all pointers are DEVICE
 * memory, the launch goes to the context's stream, no host sync.  `cells_out` (device, may be
 * NULL) accumulates the banded DP cells actually computed (sum over rows of end-beg). */
int bm2_extend_pairs_device(bm2_ctx *ctx, bm2_seqpair *d_pairs, const uint8_t *d_ref,
                            const uint8_t *d_qer, int32_t n_pairs, int32_t w, int32_t end_bonus,
                            unsigned long long *d_cells_out);

/* ---- seam 2: the whole hot path over a chunk of reads -----------------------------------------
 * Output record: layout-compatible with `mem_alnreg_t` (src/bwamem.h:137-160); the pointer slot
 * `c` is always NULL on output. */
typedef struct bm2_alnreg_t {
    int64_t rb, re;
    int32_t qb, qe;
    int32_t rid;
    int32_t pad0_;                  /* the compiler's padding of mem_alnreg_t, named so that it is written (0) */
    void   *c;
    int32_t score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0;
    int32_t n_comp_is_alt;          /* bit-field word: n_comp:30, is_alt:2                      */
    float   frac_rep;
    int32_t pad1_;
    uint64_t hash;
    int32_t flg;
    int32_t pad2_;
} bm2_alnreg_t;

/* SMEM record: `SMEM` (src/FMI_search.h:75-83). */
typedef struct bm2_smem {
    uint32_t rid;
    uint32_t m, n;
    int64_t  k, l, s;
} bm2_smem;

/* Seed / chain records (src/bwamem.h:113-135) as flat arrays. */
typedef struct bm2_seed {
    int64_t rbeg;
    int32_t qbeg, len, score;
    int32_t chain;                  /* index into the chunk's chain array                       */
} bm2_seed;

typedef struct bm2_chain {
    int64_t pos;
    int32_t seqid, rid;
    int32_t n_seeds, seed_off;      /* seeds [seed_off, seed_off+n_seeds) of the chunk's seeds   */
    int32_t w, kept, first, is_alt;
    float   frac_rep;
    int32_t _pad;
} bm2_chain;

/* A chunk of reads: concatenated base codes (0..3 = ACGT, 4 = other; exactly what
 * src/bwamem.cpp:992-1000 leaves in bseq1_t::seq) and n_reads+1 offsets. */
typedef struct bm2_read_batch {
    int32_t n_reads;
    const uint8_t *codes;
    const int64_t *offsets;
} bm2_read_batch;

/* Result arrays are owned by the context (pinned host memory) and stay valid until the next call
 * on the same context. */
typedef struct bm2_smem_result  { int64_t n; const bm2_smem *smems; const int64_t *read_off; } bm2_smem_result;
typedef struct bm2_chain_result { int64_t n_chains, n_seeds; const bm2_chain *chains; const bm2_seed *seeds;
                                  const int64_t *read_off; /* n_reads+1 offsets into chains */ } bm2_chain_result;
typedef struct bm2_reg_result   { int64_t n; const bm2_alnreg_t *regs; const int64_t *read_off; } bm2_reg_result;

/* Replaces mem_collect_smem (src/bwamem.cpp:626-804): three SMEM passes + ordering. */
int bm2_collect_smems(bm2_ctx *ctx, const bm2_read_batch *reads, bm2_smem_result *out);
/* Replaces mem_kernel1_core (src/bwamem.cpp:976-1091): SMEMs + SA lookup + chaining + filters. */
int bm2_seed_chain(bm2_ctx *ctx, const bm2_read_batch *reads, bm2_chain_result *out);
/* Replaces kt_for(worker_bwt) + kt_for(worker_aln) (src/bwamem.cpp:1359-1363): regs per read as
 * left by mem_kernel2_core (src/bwamem.cpp:1093-1172). */
int bm2_seed_chain_extend(bm2_ctx *ctx, const bm2_read_batch *reads, bm2_reg_result *out);

/* Device-resident variant for throughput measurement: `reads` still carries the HOST offsets (sizes
 * are needed on the host) but codes/offsets are taken from DEVICE memory (`d_codes`, `d_offsets`,
 * already uploaded by the caller), and the final regs stay on the device unless `copy_out` != 0
 * (out->regs is then NULL; out->n and out->read_off are valid). */
int bm2_seed_chain_extend_resident(bm2_ctx *ctx, const bm2_read_batch *reads, const uint8_t *d_codes,
                                   const int64_t *d_offsets, int copy_out, bm2_reg_result *out);

/* Per-stage device times (ms, CUDA events) of the last seam-2 call; names in `names`. */
int bm2_last_stage_ms(const bm2_ctx *ctx, const char *const **names, const float **ms, int *n);
/* Work counters of the last seam-2 call: v[0] interval extensions (128 algorithmic bytes each),
 * v[1] LF steps of the SA walk (64 B each), v[2] banded DP cells, v[3]/v[4] left/right jobs re-run
 * with the doubled band.  n >= 5. */
int bm2_last_counters(const bm2_ctx *ctx, unsigned long long *v, int n);

/* ---------------------------------------------------------------------------------------------
 * Seam 3 (first widening step): CIGAR, NM and MD of alignments whose end points are
 * known.  Replaces bwa_gen_cigar2 (src/bwa.cpp:260-347) with its banded global alignment + backtrack
 * ksw_global2 (src/ksw.cpp:558-668), called per output alignment by mem_reg2aln (src/bwamem.cpp:1757-1768):
 *     cigar = bwa_gen_cigar2(opt->mat, o_del, e_del, o_ins, e_ins, w2, bns->l_pac, pac, qe - qb, &query[qb], rb, re,
 *                            &score, &n_cigar, &NM);
 * One request = one such call; the query is reads[read][qb, qe), the scoring comes from the context's mem_opt_t.
 * --------------------------------------------------------------------------------------------- */
typedef struct bm2_cigar_req {
    int64_t rb, re;            /* reference interval in the [0, 2*l_pac) coordinate (mem_alnreg_t / mem_aln_t) */
    int32_t read;              /* index of the read in the batch                                               */
    int32_t qb, qe;            /* query interval of the read                                                   */
    int32_t w;                 /* band limit (the w_ argument)                                                 */
} bm2_cigar_req;
typedef struct bm2_cigar_rec {
    int32_t score;             /* INT32_MIN when the reference returns without setting *score (rejected)      */
    int32_t n_cigar;           /* operations: len << 4 | op (0 M, 1 I, 2 D), as the reference's uint32_t cigar */
    int32_t nm;                /* NM (-1 when rejected)                                                        */
    int32_t n_md;              /* bytes of the MD string incl. its NUL (the block appended after the cigar)    */
    int64_t cigar_off, md_off; /* offsets into bm2_cigar_result::cigar / ::md                                  */
} bm2_cigar_rec;
typedef struct bm2_cigar_result {
    int64_t n; const bm2_cigar_rec *recs;
    int64_t n_ops; const uint32_t *cigar;
    int64_t n_md; const char *md;
} bm2_cigar_result;
/* Needs a context created with an index.  Result arrays are owned by the context (valid until its next call). */
int bm2_gen_cigar(bm2_ctx *ctx, const bm2_read_batch *reads, const bm2_cigar_req *reqs, int64_t n, bm2_cigar_result *out);

/* ---- seam 4, first piece: insert-size statistics -------------------------------------------------------------------------
 * Replaces mem_pestat (reference src/bwamem_pair.cpp:81-148), called once per chunk between the alignment regions (seam 2) and
 * the SAM stage (src/bwamem.cpp:1368-1378).  Host code, as in the reference: one pass over the best region of every read; no
 * context, no device.  regs / read_off: the output of bm2_seed_chain_extend for a chunk whose reads 2i, 2i+1 are mates.
 * pes[d], d = FF, FR, RF, RR: mem_pestat_t (src/bwamem.h:162-166).  Returns 0; 1 on bad arguments; 2 where the reference
 * asserts (no insert size inside the outlier bounds). */
typedef struct bm2_pestat_t {
    int32_t low, high;         /* proper-pair bounds of the insert size                      */
    int32_t failed;            /* too few pairs of this orientation                          */
    int32_t _pad;
    double avg, std;
} bm2_pestat_t;
int bm2_pestat(const bm2_mem_opt_t *opt, int64_t l_pac, int32_t n_reads, const bm2_alnreg_t *regs, const int64_t *read_off, bm2_pestat_t pes[4]);

/* ---- finer seam under seam 4: a batch of the local alignments of mate rescue -----------------------------------------------
 * Replaces ksw_align2 (reference src/ksw.cpp:324-381: ksw_u8 / ksw_i16 forward, then the reversed prefixes for the start) as
 * mem_matesw calls it (src/bwamem_pair.cpp:186-193).  seqs: codes 0-4; a request names its query and its reference window by
 * offsets into seqs; xtra as the reference's (KSW_XBYTE 0x10000, KSW_XSTOP 0x20000, KSW_XSUBO 0x40000, KSW_XSTART 0x80000 |
 * threshold).  Scoring from the context's mem_opt_t.  out: n results, caller's memory.  Queries up to 497 bases. */
typedef struct bm2_ksw_req { int64_t qoff, toff; int32_t qlen, tlen; int32_t xtra, _pad; } bm2_ksw_req;
typedef struct bm2_ksw_res { int32_t score, te, qe, score2, te2, tb, qb, _pad; } bm2_ksw_res;       /* kswr_t (src/ksw.h:45-50) */
int bm2_ksw_align2(bm2_ctx *ctx, const uint8_t *seqs, int64_t n_seq_bytes, const bm2_ksw_req *reqs, int64_t n, bm2_ksw_res *out);

/* ---- seam 4: the SAM stage of a chunk of read pairs ----------------------------------------------------------------------
 * Replaces, for all pairs of a chunk at once, what worker_sam does per pair through mem_sam_pe (reference
 * src/bwamem_pair.cpp:349-552, MATE_SORT == 0): mate rescue (mem_matesw :150-283 over ksw_align2, src/ksw.cpp:324-381),
 * mem_mark_primary_se (src/bwamem.cpp:1420-1468), -5 reordering, mem_pair (:285-346), the MAPQ logic, mem_reg2aln with its
 * CIGAR / NM / MD (src/bwamem.cpp:1732-1805), the record selection of mem_reg2sam (:1521-1577), the columns of mem_aln2sam
 * (:1592-1730) and the entries of the XA tags (mem_gen_alt, src/bwamem_extra.cpp:130-183).  What is left to the caller is text:
 * QNAME, SEQ / QUAL (trimmed by the hard clips of the record's CIGAR), the tag syntax, SA and MC (columns of the read's other
 * records / the mate's record), -C / -R / -V constants.
 * regs / read_off: the output of bm2_seed_chain_extend for the same batch (reads 2i, 2i+1 are mates); pes: bm2_pestat or the
 * -I values.  One bm2_sam_rec per SAM line, in output order (pair by pair, read 0 then read 1).
 * A record is a true secondary (SEQ / QUAL '*', no SA / pa tags) iff (flag & 0x100) && sub < 0; a -M supplementary has 0x100 and sub >= 0.
 * tests/sam_text.py formats the reference's text from these records byte for byte (SEQ / QUAL with hard clips, NM MD MC AS XS SA pa XA). */
typedef struct bm2_sam_rec {
    int32_t read;              /* read of the batch the line belongs to                                        */
    int32_t flag;              /* FLAG as printed                                                              */
    int32_t rid, rnext;        /* contig ids of RNAME / RNEXT, -1: '*'                                         */
    int32_t mapq, nm;          /* nm valid iff n_cigar > 0                                                     */
    int32_t score, sub;        /* AS (printed if >= 0), XS (printed if >= 0)                                   */
    int32_t alt_sc;            /* > 0 and not a 0x100 record: pa:f: = score / alt_sc                           */
    int32_t reg;               /* its XA tag = the bm2_sam_xa entries of this read with the same reg; -1: none */
    int32_t n_cigar, n_md;     /* printed operations (len << 4 | index into "MIDSH"); MD bytes incl. the NUL   */
    int32_t is_alt;            /* the hit lies on an ALT contig                                                */
    int32_t n_mc;              /* MC tag: the n_mc operations after the record's own (cigar_off + n_cigar)     */
    int64_t pos, pnext, tlen;  /* as printed (1-based; 0 where the column is 0)                                */
    int64_t cigar_off, md_off; /* into bm2_sam_result::cigar / ::md                                            */
} bm2_sam_rec;
typedef struct bm2_sam_xa {    /* one entry of an XA tag: name(rid),[+-]pos,CIGAR,nm;                          */
    int32_t read, reg;
    int32_t rid, is_rev, nm, n_cigar;      /* operations: len << 4 | index into "MIDSHN"                       */
    int64_t pos;               /* 0-based: printed as pos + 1                                                  */
    int64_t cigar_off;
} bm2_sam_xa;
typedef struct bm2_sam_result {
    int64_t n_recs; const bm2_sam_rec *recs;
    int64_t n_xa; const bm2_sam_xa *xa;        /* in the order the reference appends them                      */
    int64_t n_ops; const uint32_t *cigar;
    int64_t n_md; const char *md;
} bm2_sam_result;
/* Needs a context created with an index; uses the context's mem_opt_t (flag bits -a -M -P -S -Y -5 -q included).
 * Result arrays are owned by the context (valid until its next call).  id_base: number of pairs before this batch in the
 * run (the reference's `id`, which seeds the tie-breaking hashes). */
int bm2_sam_pe(bm2_ctx *ctx, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off, const bm2_pestat_t pes[4],
               int64_t id_base, bm2_sam_result *out);
/* The single-end branch of worker_sam (src/bwamem.cpp:1320-1334: mem_mark_primary_se, -5, mem_reg2sam without a mate) for a batch of reads;
 * same records (no RNEXT / PNEXT / TLEN).  id_base: number of reads before this batch in the run. */
int bm2_sam_se(bm2_ctx *ctx, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off, int64_t id_base, bm2_sam_result *out);
/* ---- seam 0 (host I/O on the fast side): FASTQ bytes -> read batch ----------------------------------------------
 * Replaces the parsing of bseq_read_orig (src/bwa.cpp:170-216 over kseq.h; name up to the first blank, trim_readno :62-66) and the base
 * encoding at the head of mem_kernel1_core (src/bwamem.cpp:992-1000, nst_nt4_table).  buf1 / buf2: the raw bytes of a chunk of the two
 * FASTQ files (buf2 NULL: single-end); the result interleaves them (reads 2i / 2i+1 from buf1 / buf2).  Parsing and encoding run on the
 * GPU; the batch is left on the device for bm2_seed_chain_extend_resident and copied to pinned host memory for the SAM stage.
 * Four-line records only; a chunk must stay below 2 GiB per buffer.  Arrays are owned by the context (valid until its next call). */
typedef struct bm2_fastq_batch {
    int32_t n_reads;
    const uint8_t *d_codes; const int64_t *d_offsets;     /* DEVICE: codes 0-4 ('-' = 5 as nst_nt4_table), n_reads + 1 offsets */
    const uint8_t *codes;   const int64_t *offsets;       /* HOST copies                                                      */
    const char *quals;                                    /* HOST: qualities laid out like codes                              */
    const int64_t *name_beg; const int32_t *name_len;     /* HOST: QNAME of read r = its buffer [name_beg[r], + name_len[r])  */
} bm2_fastq_batch;
int  bm2_fastq_encode(bm2_ctx *ctx, const char *buf1, int64_t n1, const char *buf2, int64_t n2, bm2_fastq_batch *out);

/* ---- seam 5 (host I/O on the fast side): SAM text of a chunk --------------------------------------------------------
 * The formatting half of mem_aln2sam (src/bwamem.cpp:1592-1730): QNAME, the tab-separated columns, SEQ / QUAL trimmed by the record's
 * hard clips and reverse-complemented on the reverse strand, tags NM MD MC AS XS SA pa XA in the reference's order, one line per
 * bm2_sam_rec, byte for byte what `bwa-mem2 mem` prints for the record (without the constant -C / -R / -V additions).  Host code on
 * n_threads threads (read ranges).  *text is malloc'd (release with bm2_free), NUL-terminated, *len bytes. */
typedef struct bm2_sam_text_in {
    const bm2_sam_result *res;          /* records of bm2_sam_pe / bm2_sam_se for this batch                       */
    const bm2_read_batch *reads;        /* the batch (codes 0-4, offsets): SEQ                                     */
    const char *const *names;           /* QNAME per read (mates carry the same name); NULL: "r<index>"            */
    const char *quals;                  /* qualities laid out like reads->codes (same offsets); NULL: '*'          */
    const char *const *contig_names;    /* RNAME by contig id (bntann1_t::name)                                    */
    /* names == NULL: QNAME of read r = name_buf[paired ? r & 1 : 0][name_beg[r], + name_len[r]) - the spans bm2_fastq_encode returns
     * (name_buf[1] NULL: single-end); all NULL: "r<index>" */
    const char *name_buf[2];
    const int64_t *name_beg; const int32_t *name_len;
} bm2_sam_text_in;
int  bm2_sam_format(const bm2_sam_text_in *in, int n_threads, char **text, int64_t *len);
void bm2_free(void *p);

/* Staged mate rescue inside bm2_sam_pe (same records, other kernels): the windows mem_matesw (src/bwamem_pair.cpp:150-283) can ask for are
 * listed for all pairs of a wave from the regions before any rescue, aligned as one batch with one window per warp (the job shape of
 * bm2_ksw_align2; the reference batches the same alignments across pairs in its kswv path, src/bwamem_pair.cpp:930-1248, src/kswv.cpp),
 * and the per-pair logic looks them up - it still computes an alignment itself when an earlier rescue of the pair moved the window.
 * on = 0: off; 1: one window per warp (the row of a window split over 32 lanes); 2: one window per thread (the one-thread sweep of the per-pair
 * logic, 32 windows per warp); -1 (default) leaves the choice to the BM2_SAM_STAGED environment variable (unset: off, until the path has GPU numbers). */
int bm2_set_sam_staged(bm2_ctx *ctx, int on);
/* Device times and counters of the last bm2_sam_pe / bm2_sam_se call.  ms[0..3] (CUDA events, summed over waves): job listing, window
 * alignments (both 0 when not staged), the per-pair kernel, the gather.  counts[0..5]: staged mode (0/1/2), jobs listed, alignments looked up,
 * alignments computed in place by the per-pair kernel (staged mode only), of those the ones whose window had moved, waves.  n_ms >= 4, n_counts >= 6. */
int bm2_last_sam_stats(const bm2_ctx *ctx, double *ms, unsigned long long *counts, int n_ms, int n_counts);

#ifdef __cplusplus
}
#endif
#endif /* BM2_B200_H */
