// ksw_device.cuh — the local alignment of mate rescue for ONE (query, reference window) pair (device logic, groundwork for
// no kernel launches this yet).
//
// Replaces ksw_align2 (reference src/ksw.cpp:324-381) = ksw_u8 (:111-233) or ksw_i16 (:235-316) forward, then the same kernel on the
// reversed prefixes to find the start, as mem_matesw calls it (src/bwamem_pair.cpp:186-193).  The reference's kernels are striped
// (Farrar): vector j holds the query positions j + l*slen of the lanes l, i.e. lane l owns the CONTIGUOUS segment
// [l*slen, (l+1)*slen).  Its first pass carries F inside a lane only; the lazy-F loop then moves F across the lane boundaries, which
// are neighbouring query positions, so the completed H is the ordinary affine-gap H.  What is not ordinary, and is reproduced here:
//   * E of the next row and the row maximum (imax -> te, the score2 list) are taken from the FIRST-pass H, whose F restarts at 0
//     at every segment start l*slen  (slen = ceil(qlen / 16) for the 8-bit kernel, ceil(qlen / 8) for the 16-bit one);
//   * the padding positions qlen .. slen*p - 1 (substitution score 0) take part in the row maximum;
//   * 8-bit kernel: H + S saturates at 255 before the bias is removed; the search stops when gmax + shift >= 255.
// One sequential sweep per row with two F registers (segment-local and complete) gives exactly the reference's values
// (tests/host_emul/ksw_emul.cpp against the oracle's lane-by-lane restatement, which is pinned to the reference binary).
#pragma once
#include "hd.h"

#define BM2_KSW_XBYTE 0x10000
#define BM2_KSW_XSTOP 0x20000
#define BM2_KSW_XSUBO 0x40000
#define BM2_KSW_XSTART 0x80000

struct KswRes { int score, te, qe, score2, te2, tb, qb; };        // kswr_t (src/ksw.h:45-50)

// One pass (ksw_u8 / ksw_i16).  query[k * qstride], target[i * tstride]; mat 5x5; scratch: 3 * nlen ints (H of the previous row,
// E, H of the best row), nlen = slen * p <= qlen + 15.  The score2 list (`b` array, :188-196) is folded on the fly: it is only
// consumed through "best entry outside [te - d, te + d]", which needs the final te, so the entries are kept in bsc/bpos (cap entries;
// consecutive rows merge, so a few dozen suffice for a mate window; on overflow *overflow is set and score2 is unreliable).
BM2_HD KswRes ksw_pass_d(int size, int qlen, const uint8_t *query, int qstride, int tlen, const uint8_t *target, int tstride, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *scratch, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const int p = size == 1 ? 16 : 8;
    const int slen = (qlen + p - 1) / p, nlen = slen * p;
    int shift = 127, qmax = 0;
    for (int a = 0; a < 25; ++a) { if (mat[a] < shift) shift = mat[a]; if (mat[a] > qmax) qmax = mat[a]; }
    shift = (256 - shift) & 0xff;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int minsc = (xtra & BM2_KSW_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & BM2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
    int32_t *H = scratch, *E = scratch + nlen, *Hmax = scratch + 2 * nlen;
    for (int k = 0; k < nlen; ++k) { H[k] = 0; E[k] = 0; Hmax[k] = 0; }
    KswRes r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    int gmax = 0, te = -1, n_b = 0;
    bool hmax_is_h = false;                   // Hmax is copied lazily: H holds the best row until the next row overwrites it
    for (int i = 0; i < tlen; ++i) {
        const int8_t *ma = mat + (int) target[(long long) i * tstride] * 5;
        if (hmax_is_h) { for (int k = 0; k < nlen; ++k) Hmax[k] = H[k]; hmax_is_h = false; }
        int rowmax = 0, fseg = 0, ffull = 0, diag = 0, seg = 0;
        for (int k = 0; k < nlen; ++k) {
            if (seg == slen) { seg = 0; fseg = 0; }                     // a new lane: the first pass restarts F
            ++seg;
            const int sc = k >= qlen ? 0 : (int) ma[query[(long long) k * qstride]];
            int h = diag;                                                // H(i-1, k-1) of the completed previous row
            diag = H[k];
            if (size == 1) { h = h + sc + shift; if (h > 255) h = 255; h -= shift; if (h < 0) h = 0; }
            else { h = h + sc; if (h > 32767) h = 32767; }
            int e = E[k];
            if (e > h) h = e;
            if (fseg > h) h = fseg;                                      // first-pass H
            if (h > rowmax) rowmax = h;
            { int t = h - oe_del; if (t < 0) t = 0; e -= e_del; if (e < 0) e = 0; E[k] = e > t ? e : t; }
            const int open = h - oe_ins > 0 ? h - oe_ins : 0;
            fseg -= e_ins; if (fseg < 0) fseg = 0; if (open > fseg) fseg = open;
            const int hf = ffull > h ? ffull : h;                         // completed H
            H[k] = hf;
            ffull -= e_ins; if (ffull < 0) ffull = 0; if (open > ffull) ffull = open;
        }
        if (rowmax >= minsc) {
            if (n_b == 0 || bpos[n_b - 1] + 1 != i) {
                if (n_b < bcap) { bsc[n_b] = rowmax; bpos[n_b] = i; ++n_b; } else *overflow |= 32;            // BM2_OVF_KSW_LIST (sam_device.cuh)
            } else if (bsc[n_b - 1] < rowmax) { bsc[n_b - 1] = rowmax; bpos[n_b - 1] = i; }
        }
        if (rowmax > gmax) {
            gmax = rowmax; te = i; hmax_is_h = true;
            if (size == 1 ? (gmax + shift >= 255 || gmax >= endsc) : (gmax >= endsc)) break;
        }
    }
    const int32_t *best = hmax_is_h ? H : Hmax;
    r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (size == 2 || r.score != 255) {
        int mx = -1;
        for (int k = 0; k < nlen; ++k) if (best[k] > mx) { mx = best[k]; r.qe = k; }      // smallest position among the maxima
        if (n_b) {
            const int d = (r.score + qmax - 1) / qmax, low = te - d, high = te + d;
            for (int k = 0; k < n_b; ++k)
                if ((bpos[k] < low || bpos[k] > high) && bsc[k] > r.score2) { r.score2 = bsc[k]; r.te2 = bpos[k]; }
        }
    }
    return r;
}

// ksw_align2: forward pass, then the reversed prefixes query[qe..0], target[te..0] followed by the target's tail (the reference
// reverses the prefix in place and runs over all tlen rows, :366-371).  tmp: tlen bytes for that target order.
BM2_HD KswRes ksw_align2_d(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins,
                           int e_ins, int xtra, int32_t *scratch, int32_t *bsc, int32_t *bpos, int bcap, uint8_t *tmp, int *overflow)
{
    const int size = (xtra & BM2_KSW_XBYTE) ? 1 : 2;
    KswRes r = ksw_pass_d(size, qlen, query, 1, tlen, target, 1, mat, o_del, e_del, o_ins, e_ins, xtra, scratch, bsc, bpos, bcap, overflow);
    if ((xtra & BM2_KSW_XSTART) == 0 || ((xtra & BM2_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    for (int i = 0; i <= r.te; ++i) tmp[i] = target[r.te - i];
    for (int i = r.te + 1; i < tlen; ++i) tmp[i] = target[i];
    int ov2 = 0;
    const KswRes rr = ksw_pass_d(size, r.qe + 1, query + r.qe, -1, tlen, tmp, 1, mat, o_del, e_del, o_ins, e_ins, BM2_KSW_XSTOP | r.score, scratch, bsc, bpos,
                                 bcap, &ov2);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}
