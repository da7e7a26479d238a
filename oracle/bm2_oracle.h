/* bm2_oracle.h — CPU restatement of the bwa-mem2 seed-and-extend hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this library; the product (libbm2b200.so) never links or calls it.
 * Every function cites the reference file:line (bwa-mem2 @ 97978f95) it restates.  Parity is
 * PINNED: tests/test_oracle_*.py check each stage against dumps produced by the unmodified
 * reference (oracle/_ref/<isa>/ref_driver) and against the committed fixtures in tests/golden/.
 */
#ifndef BM2_ORACLE_H
#define BM2_ORACLE_H
#include <stdint.h>
#include "../include/bm2_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bm2o_bsw_params {
    int32_t a, b;                 /* match score, mismatch penalty (positive)       */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t zdrop, end_bonus;
    int32_t vector_quirks;        /* 1: band/z-drop as the SIMD kernels compute them */
} bm2o_bsw_params;

/* out[6] = score, qle, tle, gtle, gscore, max_off ; returns banded cells computed */
int64_t bm2o_bsw_extend(const uint8_t *query, int32_t qlen, const uint8_t *target, int32_t tlen,
                        int32_t w, int32_t h0, const bm2o_bsw_params *p, int32_t *out);

/* SeqPair batch, same contract as bm2_extend_pairs; returns total banded cells */
int64_t bm2o_extend_pairs(bm2_seqpair *pairs, const uint8_t *seq_buf_ref, const uint8_t *seq_buf_qer,
                          int32_t n_pairs, int32_t w, const bm2o_bsw_params *p);

/* ---- FM-index stages (A1-A4).  Results are malloc'd arrays the caller frees with bm2o_free. ---- */
void bm2o_free(void *p);

/* three SMEM passes + ordering == mem_collect_smem (src/bwamem.cpp:626-804); reads are processed
 * in blocks of `block` reads (BATCH_SIZE = 512, src/macro.h:48) like kt_for does. */
int64_t bm2o_collect_smems(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                           bm2_smem **out);

/* SA lookup of rows == FMI_search::get_sa_entries_prefetch (src/FMI_search.cpp:1257-1375) */
void bm2o_sa_lookup(const bm2_index_desc *idx, const int64_t *rows, int64_t n, int64_t *out);

/* SMEM + SA + chaining + chain filter == mem_kernel1_core (src/bwamem.cpp:976-1091), without
 * mem_flt_chained_seeds.  read_off has n_reads+1 entries. */
int bm2o_seed_chain(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                    bm2_chain **chains, int64_t *n_chains, bm2_seed **seeds, int64_t *n_seeds, int64_t **read_off);

/* the whole hot path == worker_bwt + worker_aln (src/bwamem.cpp:1193-1214, :1175-1191): regs per
 * read as mem_kernel2_core leaves them (src/bwamem.cpp:1093-1172).  Returns non-zero if a read
 * needs mem_flt_chained_seeds' local SW (long reads; not restated yet). */
int bm2o_seed_chain_extend(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                           bm2_alnreg_t **regs, int64_t *n_regs, int64_t **read_off, int64_t *bsw_cells);

/* CIGAR / NM / MD of alignments with known end points == bwa_gen_cigar2 (src/bwa.cpp:260-347) + ksw_global2 with backtrack
 * (src/ksw.cpp:545-668); same request / record layout as bm2_gen_cigar.  Arrays are malloc'd (bm2o_free). */
int bm2o_gen_cigar(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_cigar_req *reqs,
                   int64_t n, bm2_cigar_rec **recs, uint32_t **cigar, int64_t *n_ops, char **md, int64_t *n_md);

/* Local alignment of mate rescue == ksw_align2 (src/ksw.cpp:324-381) over ksw_u8 / ksw_i16 (:111-323), the call of mem_matesw
 * (src/bwamem_pair.cpp:189; SURVEY 8f item 1, groundwork for the next widening step).  The reference's kernels are striped
 * (Farrar) with a lazy-F loop; the E of the next row is taken from the FIRST-pass H (F propagated inside a stripe lane only), so
 * the result depends on the segmentation slen = ceil(qlen / 16 or 8): restated in scalar code with the same segmentation.
 * out[7] = score, te, qe, score2, te2, tb, qb (kswr_t).  query/target are not modified. */
void bm2o_ksw_align2(int32_t qlen, const uint8_t *query, int32_t tlen, const uint8_t *target, const int8_t *mat /*5x5*/,
                     int32_t o_del, int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out);

/* mem_pestat (src/bwamem_pair.cpp:88-148): lh[12] = low, high, failed of the orientations FF, FR, RF, RR; as[8] = avg, std. */
void bm2o_pestat(const bm2_mem_opt_t *opt, int64_t l_pac, int32_t n_reads, const bm2_alnreg_t *regs, const int64_t *read_off, int32_t *lh, double *as);
/* mem_matesw (src/bwamem_pair.cpp:150-283, MATE_SORT == 0): ma has room for *n_ma + 4 records. */
int bm2o_matesw(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const int32_t *pes_lh, const bm2_alnreg_t *a, int32_t l_ms, const uint8_t *ms,
                bm2_alnreg_t *ma, int32_t *n_ma);

/* Single-end SAM stage (worker_sam without MEM_F_PE, src/bwamem.cpp:1330-1336): mem_mark_primary_se (:1420-1468) on the read's regs
 * (modified in place, id = id_base + read index), then the records mem_reg2sam (:1521-1577) writes - mem_reg2aln (:1732-1805) per kept
 * region: MAPQ (mem_approx_mapq_se :1470-1494), CIGAR with clipping, NM, MD, POS.  One bm2o_aln per output line (an unmapped read
 * gives one with rid = -1); cigar ops as len << 4 | op with op 3 = clip (printed S or H); md NUL-terminated. */
typedef struct bm2o_aln {
    int32_t read, flag, rid, mapq, nm, score, sub, is_rev, is_alt, alt_sc, n_cigar, n_md;
    int64_t pos, cigar_off, md_off;
} bm2o_aln;
int bm2o_sam_se(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, bm2_alnreg_t *regs, const int64_t *read_off,
                int64_t id_base, bm2o_aln **alns, int64_t *n_alns, uint32_t **cigar, int64_t *n_ops, char **md, int64_t *n_md);

/* Paired-end SAM stage == mem_sam_pe (src/bwamem_pair.cpp:353-552, MATE_SORT == 0) per read pair: mate rescue (mem_matesw), mem_mark_primary_se,
 * mem_pair (:285-346), the paired / unpaired MAPQ logic, and the columns mem_aln2sam (src/bwamem.cpp:1592-1730) prints.  pes_lh[12] = low, high,
 * failed; pes_as[8] = avg, std (mem_pestat's result for the chunk).  regs of pair p: reads 2p and 2p+1; modified in place is not visible to the
 * caller (copies).  One bm2o_samrec per SAM line in output order; cigar ops are len << 4 | op with op an index into "MIDSH" as printed.
 * Not restated: XA / SA / MC / pa tags, -5 (MEM_F_PRIMARY5), -C. */
typedef struct bm2o_samrec {
    int32_t read, flag, rid, mapq, rnext, tlen_valid, nm, score, sub, n_cigar, n_md, _pad;   /* rid / rnext: -1 = '*'; sub < 0: no XS; nm valid iff n_cigar */
    int64_t pos, pnext, tlen, cigar_off, md_off;                                           /* pos / pnext 1-based as printed */
} bm2o_samrec;
int bm2o_sam_pe(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                const int32_t *pes_lh, const double *pes_as, int64_t id_base, bm2o_samrec **recs, int64_t *n_recs, uint32_t **cigar, int64_t *n_ops,
                char **md, int64_t *n_md);

/* The same stage as SAM TEXT: every line of mem_sam_pe from the FLAG column on (tags NM MD MC AS XS SA pa XA as mem_aln2sam writes them, src/bwamem.cpp:1592-1730;
 * XA by mem_gen_alt, src/bwamem_extra.cpp:130-183).  names: contig names; quals: qualities laid out like reads->codes, or NULL ('*'). */
int bm2o_sam_pe_text(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const char *quals, const char *const *names,
                     const bm2_alnreg_t *regs, const int64_t *read_off, const int32_t *pes_lh, const double *pes_as, int64_t id_base, char **text, int64_t *len);

/* bench.py's cpu_baseline leg: aggregate iterations/s of a fixed integer loop on `nthreads` pthreads (effective cores = rate(n)/rate(1)) */
double bm2o_cpu_probe(int32_t nthreads, double seconds);

#ifdef __cplusplus
}
#endif
#endif
