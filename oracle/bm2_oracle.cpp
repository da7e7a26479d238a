/* bm2_oracle.cpp — CPU restatement (oracle) of the bwa-mem2 hot path.  TEST INFRASTRUCTURE ONLY:
 * see bm2_oracle.h.  Plain sequential C++, written from the algorithm description in SURVEY.md
 * Appendix A and checked stage by stage against the unmodified reference. */
#include "bm2_oracle.h"
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>
#include <algorithm>

/* ------------------------------------------------------------------------------------------------
 * A6. Banded affine-gap extension DP.
 * Restates BandedPairWiseSW::scalarBandedSWA (src/bandedSWA.cpp:116-237) == ksw_extend2
 * (src/ksw.cpp:432-533).  With vector_quirks the band is derived as the SIMD wrappers do
 * (integer division, src/bandedSWA.cpp:2905-2926) and z-drop ignores the gap-extension
 * multiplier (ZSCORE16, src/bandedSWA.cpp:1868-1881); both coincide with the scalar code when
 * e_del == e_ins == 1.
 * ---------------------------------------------------------------------------------------------- */
static inline int sub_score(const bm2o_bsw_params *p, uint8_t t, uint8_t q) {
    if (t > 3 || q > 3) return -1;                     /* DEFAULT_AMBIG, bandedSWA.h:53; bwa.cpp:248 */
    return t == q ? p->a : -p->b;
}

extern "C" int64_t bm2o_bsw_extend(const uint8_t *query, int32_t qlen, const uint8_t *target, int32_t tlen,
                                   int32_t w, int32_t h0, const bm2o_bsw_params *p, int32_t *out)
{
    const int oe_del = p->o_del + p->e_del, oe_ins = p->o_ins + p->e_ins;
    std::vector<int32_t> H(qlen + 2, 0), E(qlen + 2, 0);   /* H[j] = H(i-1, j-1), E[j] = E(i, j) */
    int64_t cells = 0;

    /* first row: gap-open from h0 then extension (bandedSWA.cpp:141-144) */
    H[0] = h0;
    if (qlen >= 1) H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && H[j - 1] > p->e_ins; ++j) H[j] = H[j - 1] - p->e_ins;

    /* band clipping (bandedSWA.cpp:146-156) */
    int maxsc = p->a;                                   /* max of mat[] for bwa_fill_scmat matrices */
    int max_ins, max_del;
    if (p->vector_quirks) {
        uint16_t t1 = (uint16_t)((uint16_t)(qlen * maxsc) + (uint16_t)(int16_t)(p->end_bonus - p->o_ins));
        max_ins = (int)((double)(t1 / p->e_ins) + 1.0);
        uint16_t t2 = (uint16_t)((uint16_t)(qlen * maxsc) + (uint16_t)(int16_t)(p->end_bonus - p->o_del));
        max_del = (int)((double)(t2 / p->e_del) + 1.0);
    } else {
        max_ins = (int)((double)(qlen * maxsc + p->end_bonus - p->o_ins) / p->e_ins + 1.);
        max_del = (int)((double)(qlen * maxsc + p->end_bonus - p->o_del) / p->e_del + 1.);
    }
    if (max_ins < 1) max_ins = 1;
    if (max_del < 1) max_del = 1;
    if (w > max_ins) w = max_ins;
    if (w > max_del) w = max_del;

    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h_left;                                     /* H(i, beg-1) */
        if (beg == 0) { h_left = h0 - (p->o_del + p->e_del * (i + 1)); if (h_left < 0) h_left = 0; }
        else h_left = 0;
        int f = 0, row_max = 0, row_arg = -1;
        int j;
        for (j = beg; j < end; ++j) {
            int diag = H[j], e = E[j];
            H[j] = h_left;
            int M = diag ? diag + sub_score(p, target[i], query[j]) : 0;
            int h = M > e ? M : e;
            if (f > h) h = f;
            h_left = h;
            if (h >= row_max) row_arg = j;              /* last column attaining the max (:188) */
            if (h > row_max) row_max = h;
            int t = M - oe_del; if (t < 0) t = 0;
            e -= p->e_del; if (t > e) e = t;
            E[j] = e;
            t = M - oe_ins; if (t < 0) t = 0;
            f -= p->e_ins; if (t > f) f = t;
            ++cells;
        }
        H[end] = h_left; E[end] = 0;
        if (j == qlen) {                                /* row reached the query end (:202-205) */
            if (h_left >= gscore) best_ie = i;
            if (h_left > gscore) gscore = h_left;
        }
        if (row_max == 0) break;
        if (row_max > best) {
            best = row_max; best_i = i; best_j = row_arg;
            int d = row_arg - i; if (d < 0) d = -d;
            if (d > max_off) max_off = d;
        } else if (p->zdrop > 0) {
            int di = i - best_i, dj = row_arg - best_j;
            if (di > dj) {
                int pen = p->vector_quirks ? (di - dj) : (di - dj) * p->e_del;
                if (best - row_max - pen > p->zdrop) break;
            } else {
                int pen = p->vector_quirks ? (dj - di) : (dj - di) * p->e_ins;
                if (best - row_max - pen > p->zdrop) break;
            }
        }
        /* shrink the band to the non-zero support of the row just written (:218-221) */
        for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    out[0] = best; out[1] = best_j + 1; out[2] = best_i + 1; out[3] = best_ie + 1; out[4] = gscore; out[5] = max_off;
    return cells;
}

extern "C" int64_t bm2o_extend_pairs(bm2_seqpair *pairs, const uint8_t *seq_buf_ref, const uint8_t *seq_buf_qer,
                                     int32_t n_pairs, int32_t w, const bm2o_bsw_params *p)
{
    int64_t cells = 0;
    for (int i = 0; i < n_pairs; ++i) {
        bm2_seqpair *sp = &pairs[i];
        int32_t o[6];
        cells += bm2o_bsw_extend(seq_buf_qer + sp->idq, sp->len2, seq_buf_ref + sp->idr, sp->len1, w, sp->h0, p, o);
        sp->score = o[0]; sp->qle = o[1]; sp->tle = o[2]; sp->gtle = o[3]; sp->gscore = o[4]; sp->max_off = o[5];
    }
    return cells;
}
