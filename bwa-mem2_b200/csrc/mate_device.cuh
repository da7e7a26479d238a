// mate_device.cuh — mate rescue of ONE read pair (device logic of SURVEY §8(f) item 1; launched by sam.cu's per-pair kernel, staged form in mate_stage.cuh).
//
// Replaces the rescue block of mem_sam_pe (reference src/bwamem_pair.cpp:378-412, MATE_SORT == 0) and mem_matesw (:150-283):
// for the best alignments of each read, align the mate inside the window the insert-size statistics predict (ksw_align2,
// ksw_device.cuh), add the hit to the mate's regions kept sorted by score, then mem_sort_dedup_patch without patching.
// The block is sequential per pair (every call sees the regions the previous calls added), so the unit of parallelism is the pair;
// the local alignment inside is the part worth a warp.  tests/host_emul/mate_emul.cpp checks it against the oracle, which is
// pinned to the reference's own mem_matesw.
#pragma once
#include "ext_device.cuh"
#include "ksw_device.cuh"

struct MatePes { int low[4], high[4], failed[4]; };            // mem_pestat_t without the moments (src/bwamem.h:162-166)

BM2_HD int mate_infer_dir_d(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist) {      // mem_infer_dir (src/bwamem_pair.cpp:57-65)
    const int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
    const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
    *dist = p2 > b1 ? p2 - b1 : b1 - p2;
    return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

// scratch of one pair
struct MateScratch {
    uint8_t *rev;            // l_ms bytes: the reverse complement of the mate
    uint8_t *tmp; int tcap;  // tcap bytes, tcap >= mate_window_max_d(): ksw_align2's reversed target
    int32_t *ksw;            // 3 * (l_ms + 16) ints
    int32_t *bsc, *bpos; int bcap;       // score2 candidates: rows of one window, neighbours merged -> bcap >= tcap / 2 + 1 never overflows
    int32_t *idx;            // regions + 4 ints
    TailSortKey *keys;       // regions + 4 keys
};

// longest window mem_matesw can ask for with these statistics and a mate of l_ms bases (:173-181): high - low + l_ms
BM2_HD int mate_window_max_d(const MatePes &pes, int l_ms) {
    int w = 0;
    for (int r = 0; r < 4; ++r) if (!pes.failed[r] && pes.high[r] - pes.low[r] > w) w = pes.high[r] - pes.low[r];
    return w + l_ms;
}

// The window of orientation r around anchor a for a mate of l_ms bases (src/bwamem_pair.cpp:168-185, bns_fetch_seq's clipping to the
// contig of the window's middle, src/bntseq.cpp:453-482).  False: the reference does not align (other contig, window too short).
BM2_HD bool mate_window_d(const ContigView &cv, int min_seed_len, const MatePes &pes, const bm2_alnreg_t *a, int l_ms, int r, int64_t *rb_, int64_t *re_, int *is_rev_) {
    const int64_t l_pac = cv.l_pac;
    const int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1);
    int64_t rb, re;
    if (!is_rev) {
        rb = is_larger ? a->rb + pes.low[r] : a->rb - pes.high[r];
        re = (is_larger ? a->rb + pes.high[r] : a->rb - pes.low[r]) + l_ms;
    } else {
        rb = (is_larger ? a->rb + pes.low[r] : a->rb - pes.high[r]) - l_ms;
        re = is_larger ? a->rb + pes.high[r] : a->rb - pes.low[r];
    }
    if (rb < 0) rb = 0;
    if (re > l_pac << 1) re = l_pac << 1;
    int rid = -1;
    if (rb < re) {
        const int64_t mid = (rb + re) >> 1;
        rid = bns_pos2rid_d(cv, bns_depos_d(cv, mid));
        int64_t far_beg = cv.ann_off[rid], far_end = far_beg + cv.ann_len[rid];
        if (mid >= l_pac) { const int64_t t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
        rb = rb > far_beg ? rb : far_beg;
        re = re < far_end ? re : far_end;
    }
    *rb_ = rb; *re_ = re; *is_rev_ = is_rev;
    return a->rid == rid && re - rb >= min_seed_len;
}

// which orientations mem_matesw would align for anchor a given the mate's regions ma[0..nm) (:159-166)
BM2_HD int mate_skip_d(const ContigView &cv, const MatePes &pes, const bm2_alnreg_t *a, const bm2_alnreg_t *ma, int nm, int skip[4]) {
    for (int r = 0; r < 4; ++r) skip[r] = pes.failed[r] ? 1 : 0;
    for (int i = 0; i < nm; ++i) {
        int64_t dist;
        const int r = mate_infer_dir_d(cv.l_pac, a->rb, ma[i].rb, &dist);
        if (dist >= pes.low[r] && dist <= pes.high[r]) skip[r] = 1;
    }
    return skip[0] + skip[1] + skip[2] + skip[3];
}

// the local alignment computed in place (the provider of the one-thread version; a staged driver passes a table lookup instead)
struct MateKswDirect {
    const ExtParams *ep; const uint8_t *ref; const MateScratch *sc; int *overflow;
    BM2_HD KswRes operator()(int /*anchor_read*/, int /*anchor*/, int /*r*/, int l_ms, const uint8_t *seq, int64_t rb, int64_t re, int xtra) const {
        return ksw_align2_d(l_ms, seq, (int) (re - rb), ref + rb, ep->mat, ep->o_del, ep->e_del, ep->o_ins, ep->e_ins, xtra, sc->ksw, sc->bsc, sc->bpos, sc->bcap,
                            sc->tmp, overflow);
    }
};

// mem_matesw: a = the anchor (anchor j of read ai), ms = the mate's codes; ma[0..*n_ma) the mate's regions with room for 4 more.  Returns n.
// ksw(ai, j, r, l_ms, seq, rb, re, xtra) supplies the local alignment of orientation r (seq = the mate, reverse-complemented in sc.rev if needed).
template <class Ksw>
BM2_HD int matesw_d(const ContigView &cv, const ExtParams &ep, int min_seed_len, const MatePes &pes, const uint8_t *ref, const bm2_alnreg_t *a, int l_ms,
                    const uint8_t *ms, bm2_alnreg_t *ma, int *n_ma, const MateScratch &sc, int ai, int j, Ksw &ksw, int *overflow)
{
    const int64_t l_pac = cv.l_pac;
    int skip[4], n = 0, nm = *n_ma;
    if (mate_skip_d(cv, pes, a, ma, nm, skip) == 4) return 0;
    for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        int64_t rb, re; int is_rev;
        if (mate_window_d(cv, min_seed_len, pes, a, l_ms, r, &rb, &re, &is_rev)) {
            if (re - rb > sc.tcap) { *overflow |= 64; continue; }             // BM2_OVF_WINDOW (sam_device.cuh): scratch sized for other statistics
            const uint8_t *seq = ms;
            if (is_rev) { for (int i = 0; i < l_ms; ++i) sc.rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4; seq = sc.rev; }
            const int xtra = BM2_KSW_XSUBO | BM2_KSW_XSTART | (l_ms * ep.a < 250 ? BM2_KSW_XBYTE : 0) | (min_seed_len * ep.a);
            const KswRes al = ksw(ai, j, r, l_ms, seq, rb, re, xtra);
            if (al.score >= min_seed_len && al.qb >= 0) {
                alignas(16) bm2_alnreg_t b; memset(&b, 0, sizeof(b));          // reg_copy moves 16-byte words
                b.rid = a->rid;
                reg_set_is_alt_d(b, (a->n_comp_is_alt >> 30) & 3);
                b.qb = is_rev ? l_ms - (al.qe + 1) : al.qb;
                b.qe = is_rev ? l_ms - al.qb : al.qe + 1;
                b.rb = is_rev ? (l_pac << 1) - (rb + al.te + 1) : rb + al.tb;
                b.re = is_rev ? (l_pac << 1) - (rb + al.tb) : rb + al.te + 1;
                b.score = al.score; b.csub = al.score2; b.secondary = -1;
                b.seedcov = (int) ((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
                int i;
                for (i = 0; i < nm; ++i) if (ma[i].score < b.score) break;              // keep ma sorted by score (:233-238)
                for (int k = nm; k > i; --k) reg_copy(&ma[k], &ma[k - 1]);
                reg_copy(&ma[i], &b); ++nm;
            }
            ++n;
        }
        if (n) nm = sort_dedup_patch_d(cv, ep, ref, nullptr, nm, ma, nullptr, sc.idx, sc.keys);
    }
    *n_ma = nm;
    return n;
}

// The rescue block of mem_sam_pe for one pair.  a[i] / n[i]: the regions of read i (capacity n[i] + 4 * max_matesw... the caller sizes it:
// every mem_matesw call adds at most 4), seq / l_seq the reads; b0 / b1: scratch for the anchor copies (n[i] records each).
template <class Ksw>
BM2_HD int mate_rescue_pair_d(const ContigView &cv, const ExtParams &ep, int min_seed_len, int pen_unpaired, int max_matesw, const MatePes &pes,
                              const uint8_t *ref, const uint8_t *const seq[2], const int l_seq[2], bm2_alnreg_t *const a[2], int n[2],
                              bm2_alnreg_t *const b[2], const MateScratch &sc, Ksw &ksw, int *overflow)
{
    int nb[2] = { 0, 0 }, total = 0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < n[i]; ++j)
            if (a[i][j].score >= a[i][0].score - pen_unpaired) reg_copy(&b[i][nb[i]++], &a[i][j]);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < nb[i] && j < max_matesw; ++j)
            total += matesw_d(cv, ep, min_seed_len, pes, ref, &b[i][j], l_seq[!i], seq[!i], a[!i], &n[!i], sc, i, j, ksw, overflow);
    return total;
}

BM2_HD int mate_rescue_pair_d(const ContigView &cv, const ExtParams &ep, int min_seed_len, int pen_unpaired, int max_matesw, const MatePes &pes,
                              const uint8_t *ref, const uint8_t *const seq[2], const int l_seq[2], bm2_alnreg_t *const a[2], int n[2],
                              bm2_alnreg_t *const b[2], const MateScratch &sc, int *overflow)
{
    MateKswDirect direct = { &ep, ref, &sc, overflow };
    return mate_rescue_pair_d(cv, ep, min_seed_len, pen_unpaired, max_matesw, pes, ref, seq, l_seq, a, n, b, sc, direct, overflow);
}

// Staged form, first stage: the local alignments the rescue block of a pair CAN ask for, decided from the regions before any rescue.
// A later call of the block sees the regions earlier calls added, which only turns more orientations off (a hit inside the proper window
// is what switches an orientation off), except when the dedup that follows an insertion drops the hit that had switched it off - so the
// second stage (mate_rescue_pair_d with a lookup) must still be able to compute a missing alignment itself.
// emit(anchor_read, anchor, r, rb, re, is_rev): the mate is read !anchor_read.
template <class EmitJob>
BM2_HD void mate_jobs_pair_d(const ContigView &cv, int min_seed_len, int pen_unpaired, int max_matesw, const MatePes &pes, const int l_seq[2],
                             const bm2_alnreg_t *const a[2], const int n[2], EmitJob &emit)
{
    for (int i = 0; i < 2; ++i) {
        int nb = 0;
        for (int j = 0; j < n[i] && nb < max_matesw; ++j) {
            if (!(a[i][j].score >= a[i][0].score - pen_unpaired)) continue;
            int skip[4];
            if (mate_skip_d(cv, pes, &a[i][j], a[!i], n[!i], skip) < 4)
                for (int r = 0; r < 4; ++r) {
                    if (skip[r]) continue;
                    int64_t rb, re; int is_rev;
                    if (mate_window_d(cv, min_seed_len, pes, &a[i][j], l_seq[!i], r, &rb, &re, &is_rev)) emit(i, nb, r, rb, re, is_rev);
                }
            ++nb;
        }
    }
}
