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

#ifdef __cplusplus
}
#endif
#endif
