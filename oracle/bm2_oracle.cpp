/* bm2_oracle.cpp — CPU restatement (oracle) of the bwa-mem2 hot path.  TEST INFRASTRUCTURE ONLY:
 * see bm2_oracle.h.  Plain sequential C++, written from the algorithm description in SURVEY.md
 * Appendix A and checked stage by stage against the unmodified reference. */
#include "bm2_oracle.h"
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cmath>
#include <string>
#include <climits>

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
    /* the SIMD class of the job (sortPairsLenExt, src/bwamem.cpp:1944-1952): the 8-bit kernel keeps the band operands, the
     * z-drop threshold and the z-drop arithmetic in 8 bits (src/bandedSWA.cpp:2195-2216, :1826-1839, :2347), the 16-bit one in 16 */
    const int minlen = qlen < tlen ? qlen : tlen;
    const bool k8 = p->vector_quirks && tlen < 128 && qlen < 128 && h0 + minlen * p->a < 128;
    const int zthr = !p->vector_quirks ? p->zdrop : (k8 ? (int) (int8_t) p->zdrop : (int) (int16_t) p->zdrop);
    if (p->vector_quirks) {
        const unsigned mask = k8 ? 0xFFu : 0xFFFFu;
        unsigned t1 = ((unsigned) (qlen * maxsc) + (unsigned) (p->end_bonus - p->o_ins)) & mask;
        max_ins = (int)((double)(t1 / p->e_ins) + 1.0);
        unsigned t2 = ((unsigned) (qlen * maxsc) + (unsigned) (p->end_bonus - p->o_del)) & mask;
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
            /* the SIMD kernels evaluate the z-drop test on every row (ZSCORE8/16 after the best-score update: 0 > zdrop),
             * which fires for a threshold that went negative in 8 / 16 bits (-d 128..255 in the 8-bit class) */
            if (p->vector_quirks && 0 > zthr) break;
        } else if (p->vector_quirks || p->zdrop > 0) {     /* no `zdrop > 0` guard in the SIMD kernels: -d 0 drops at once */
            int di = i - best_i, dj = row_arg - best_j;
            if (di > dj) {
                int pen = p->vector_quirks ? (di - dj) : (di - dj) * p->e_del;
                if (best - row_max - pen > zthr) break;
            } else {
                int pen = p->vector_quirks ? (dj - di) : (dj - di) * p->e_ins;
                if (best - row_max - pen > zthr) break;
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

/* ================================================================================================
 * FM-index stages
 * ============================================================================================== */
extern "C" void bm2o_free(void *p) { free(p); }

namespace {

struct Fm {
    const bm2_index_desc *x;
    /* A1. Occ(b, pp): checkpoint count + popcount of the one-hot word masked to the first pp&63
     * symbols (GET_OCC, src/FMI_search.h:66-73; mask table src/FMI_search.cpp:386-394). */
    inline int64_t occ(int b, int64_t pp) const {
        const bm2_cp_occ &c = x->cp_occ[pp >> 6];
        int y = (int)(pp & 63);
        uint64_t mask = y ? ~0ULL << (64 - y) : 0ULL;
        return c.cp_count[b] + __builtin_popcountll(c.one_hot_bwt_str[b] & mask);
    }
};

struct Iv { int64_t k, l, s; };

/* backwardExt (src/FMI_search.cpp:1025-1052) */
static Iv backward_ext(const Fm &fm, Iv in, int a) {
    int64_t kk[4], ss[4], ll[4];
    for (int b = 0; b < 4; ++b) {
        int64_t o1 = fm.occ(b, in.k), o2 = fm.occ(b, in.k + in.s);
        kk[b] = fm.x->count[b] + o1;
        ss[b] = o2 - o1;
    }
    int64_t sent = (in.k <= fm.x->sentinel_index && in.k + in.s > fm.x->sentinel_index) ? 1 : 0;
    ll[3] = in.l + sent; ll[2] = ll[3] + ss[3]; ll[1] = ll[2] + ss[2]; ll[0] = ll[1] + ss[1];
    Iv out = { kk[a], ll[a], ss[a] };
    return out;
}
/* forward extension = backward extension of the swapped interval with the complement base
 * (src/FMI_search.cpp:544-553) */
static Iv forward_ext(const Fm &fm, Iv in, int a) {
    Iv sw = { in.l, in.k, in.s };
    Iv r = backward_ext(fm, sw, 3 - a);
    Iv out = { r.l, r.k, r.s };
    return out;
}
static Iv init_iv(const Fm &fm, int a) {
    Iv v = { fm.x->count[a], fm.x->count[3 - a], fm.x->count[a + 1] - fm.x->count[a] };
    return v;
}

struct Sm { int32_t m, n; Iv v; };

/* A2. one (read, x, min_intv) SMEM search (getSMEMsOnePosOneThread, src/FMI_search.cpp:496-670);
 * appends to out, returns next_x. */
static int smem_one_pos(const Fm &fm, const uint8_t *q, int len, int x, int min_intv, int min_seed_len,
                        uint32_t rid, std::vector<bm2_smem> &out)
{
    int next_x = x + 1;
    if (q[x] > 3) return next_x;
    std::vector<Sm> prev;
    Sm cur = { x, x, init_iv(fm, q[x]) };
    int j;
    for (j = x + 1; j < len; ++j) {
        next_x = j + 1;
        if (q[j] > 3) break;
        Iv nv = forward_ext(fm, cur.v, q[j]);
        if (nv.s != cur.v.s) prev.push_back(cur);
        if (nv.s < min_intv) { next_x = j; break; }
        cur.v = nv; cur.n = j;
    }
    if (cur.v.s >= min_intv) prev.push_back(cur);
    std::reverse(prev.begin(), prev.end());
    auto emit = [&](const Sm &s) {
        bm2_smem o; o.rid = rid; o.m = s.m; o.n = s.n; o.k = s.v.k; o.l = s.v.l; o.s = s.v.s; out.push_back(o);
    };
    for (j = x - 1; j >= 0; --j) {
        if (q[j] > 3) break;
        std::vector<Sm> curr;
        int curr_s = -1;
        size_t p = 0;
        for (; p < prev.size(); ++p) {
            Iv nv = backward_ext(fm, prev[p].v, q[j]);
            if (nv.s < min_intv && prev[p].n - prev[p].m + 1 >= min_seed_len) { emit(prev[p]); break; }
            if (nv.s >= min_intv && nv.s != curr_s) {
                curr_s = (int) nv.s;
                Sm t = { j, prev[p].n, nv }; curr.push_back(t);
                break;
            }
        }
        for (++p; p < prev.size(); ++p) {
            Iv nv = backward_ext(fm, prev[p].v, q[j]);
            if (nv.s >= min_intv && nv.s != curr_s) {
                curr_s = (int) nv.s;
                Sm t = { j, prev[p].n, nv }; curr.push_back(t);
            }
        }
        prev.swap(curr);
        if (prev.empty()) break;
    }
    if (!prev.empty() && prev[0].n - prev[0].m + 1 >= min_seed_len) emit(prev[0]);
    return next_x;
}

/* pass 3 (bwtSeedStrategyAllPosOneThread, src/FMI_search.cpp:726-812) */
static void seed_strategy(const Fm &fm, const uint8_t *q, int len, int max_intv, int min_len, uint32_t rid,
                          std::vector<bm2_smem> &out)
{
    int x = 0;
    while (x < len) {
        int next_x = x + 1;
        if (q[x] < 4) {
            Iv v = init_iv(fm, q[x]);
            for (int j = x + 1; j < len; ++j) {
                next_x = j + 1;
                if (q[j] > 3) break;
                v = forward_ext(fm, v, q[j]);
                if (v.s < max_intv && j - x + 1 >= min_len) {
                    if (v.s > 0) { bm2_smem o; o.rid = rid; o.m = x; o.n = j; o.k = v.k; o.l = v.l; o.s = v.s; out.push_back(o); }
                    break;
                }
            }
        }
        x = next_x;
    }
}

/* all three passes for one read, ordered (m asc, n asc) (src/bwamem.cpp:626-804,
 * sortSMEMs src/FMI_search.cpp:987-1022) */
static void collect_read(const Fm &fm, const bm2_mem_opt_t *opt, const uint8_t *q, int len, uint32_t rid,
                         std::vector<bm2_smem> &out)
{
    size_t first = out.size();
    if (len <= 0) return;
    int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
    for (int x = 0; x < len;) x = smem_one_pos(fm, q, len, x, 1, opt->min_seed_len, rid, out);
    size_t n1 = out.size();
    for (size_t i = first; i < n1; ++i) {
        bm2_smem p = out[i];
        int start = p.m, end = p.n + 1;
        if (end - start < split_len || p.s > opt->split_width) continue;
        smem_one_pos(fm, q, len, (end + start) >> 1, (int)(p.s + 1), opt->min_seed_len, rid, out);
    }
    if (opt->max_mem_intv > 0) seed_strategy(fm, q, len, (int) opt->max_mem_intv, opt->min_seed_len + 1, rid, out);
    std::stable_sort(out.begin() + first, out.end(), [](const bm2_smem &a, const bm2_smem &b) {
        return (((uint64_t) a.m << 32) | a.n) < (((uint64_t) b.m << 32) | b.n);
    });
}

/* A3. SA of one BWT row: LF-walk to a sampled row (call_one_step, src/FMI_search.cpp:1202-1255);
 * the walk returns 0 when it meets the sentinel (:1230-1233). */
static int64_t sa_of_row(const Fm &fm, int64_t r) {
    int64_t steps = 0;
    while (r & 7) {
        const bm2_cp_occ &c = fm.x->cp_occ[r >> 6];
        int y = 63 - (int)(r & 63);
        int b = 4;
        for (int t = 0; t < 4; ++t) if ((c.one_hot_bwt_str[t] >> y) & 1) { b = t; break; }
        if (b == 4) return 0;
        r = fm.x->count[b] + fm.occ(b, r);
        ++steps;
    }
    int64_t sa = ((int64_t) fm.x->sa_ms_byte[r >> 3] << 32) + (int64_t) fm.x->sa_ls_word[r >> 3];
    return sa + steps;
}

/* bns_depos / bns_pos2rid / bns_intv2rid (src/bntseq.h:87-90, src/bntseq.cpp:378-402) */
static int pos2rid(const bm2_index_desc *x, int64_t pos_f) {
    if (pos_f >= x->l_pac) return -1;
    int left = 0, mid = 0, right = x->n_seqs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= x->ann_offset[mid]) {
            if (mid == x->n_seqs - 1) break;
            if (pos_f < x->ann_offset[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
static int64_t depos(const bm2_index_desc *x, int64_t pos) { return pos >= x->l_pac ? (x->l_pac << 1) - 1 - pos : pos; }
static int intv2rid(const bm2_index_desc *x, int64_t rb, int64_t re) {
    if (rb < x->l_pac && re > x->l_pac) return -2;
    int rid_b = pos2rid(x, depos(x, rb));
    int rid_e = rb < re ? pos2rid(x, depos(x, re - 1)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

/* ks_introsort (src/ksort.h:185-232) restated on indices: same comparisons, same swaps, hence the
 * same order among ties.  lt(a,b) is the reference's __sort_lt. */
template <class T, class LT> static void ks_insertsort(T *a, long s, long t, LT lt) {   /* [s,t) */
    for (long i = s + 1; i < t; ++i)
        for (long j = i; j > s && lt(a[j], a[j - 1]); --j) std::swap(a[j], a[j - 1]);
}
template <class T, class LT> static void ks_combsort(T *a, long n, LT lt) {
    const double shrink = 1.2473309501039786540366528676643;
    long gap = n; bool swapped;
    do {
        if (gap > 2) { gap = (long)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
        swapped = false;
        for (long i = 0; i + gap < n; ++i) if (lt(a[i + gap], a[i])) { std::swap(a[i], a[i + gap]); swapped = true; }
    } while (swapped || gap > 2);
    if (gap != 1) ks_insertsort(a, 0, n, lt);
}
template <class T, class LT> static void ks_introsort(T *a, long n, LT lt) {
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) std::swap(a[0], a[1]); return; }
    int d; for (d = 2; (1ul << d) < (unsigned long) n; ++d) {}
    struct Fr { long l, r; int d; };
    std::vector<Fr> stack;
    long s = 0, t = n - 1; d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { ks_combsort(a + s, t - s + 1, lt); t = s; continue; }
            long i = s, j = t, k = i + ((j - i) >> 1) + 1;
            if (lt(a[k], a[i])) { if (lt(a[k], a[j])) k = j; }
            else k = lt(a[j], a[i]) ? i : j;
            T rp = a[k];
            if (k != t) std::swap(a[k], a[t]);
            for (;;) {
                do ++i; while (lt(a[i], rp));
                do --j; while (i <= j && lt(rp, a[j]));
                if (j <= i) break;
                std::swap(a[i], a[j]);
            }
            std::swap(a[i], a[t]);
            if (i - s > t - i) {
                if (i - s > 16) { Fr f = { s, i - 1, d }; stack.push_back(f); }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { Fr f = { i + 1, t, d }; stack.push_back(f); }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (stack.empty()) { ks_insertsort(a, 0, n, lt); return; }
            Fr f = stack.back(); stack.pop_back(); s = f.l; t = f.r; d = f.d;
        }
    }
}

struct OSeed { int64_t rbeg; int32_t qbeg, len, score; };
struct OChain {
    int64_t pos; int32_t rid, seqid, is_alt; int32_t w, kept, first; float frac_rep;
    std::vector<OSeed> seeds;
};

/* test_and_merge (src/bwamem.cpp:357-399) */
static bool test_and_merge(const bm2_mem_opt_t *opt, int64_t l_pac, OChain &c, const OSeed &p, int seed_rid) {
    const OSeed &last = c.seeds.back(), &first = c.seeds.front();
    int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
    if (seed_rid != c.rid) return false;
    if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return true;
    if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return false;
    int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
    if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last.len < opt->max_chain_gap && y - last.len < opt->max_chain_gap) {
        c.seeds.push_back(p);
        return true;
    }
    return false;
}

/* mem_chain_weight (src/bwamem.cpp:429-448) */
static int chain_weight(const OChain &c) {
    int64_t end = 0; int w = 0, tmp;
    for (const OSeed &s : c.seeds) {
        if (s.qbeg >= end) w += s.len; else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
        end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
    }
    tmp = w; w = 0; end = 0;
    for (const OSeed &s : c.seeds) {
        if (s.rbeg >= end) w += s.len; else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
        end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
    }
    w = w < tmp ? w : tmp;
    return w < (1 << 30) ? w : (1 << 30) - 1;
}

/* mem_chain_flt (src/bwamem.cpp:506-624) for the chains of one read */
static void chain_filter(const bm2_mem_opt_t *opt, std::vector<OChain> &a) {
    if (a.empty()) return;
    std::vector<OChain> kept0;
    for (OChain &c : a) { c.first = -1; c.kept = 0; c.w = chain_weight(c); }
    for (OChain &c : a) if (c.w >= opt->min_chain_weight) kept0.push_back(c);
    if (kept0.empty()) kept0.push_back(a[0]);     /* reference quirk: range (0,1) is processed even when k == 0 */
    a.swap(kept0);
    int n = (int) a.size();
    ks_introsort(a.data(), n, [](const OChain &x, const OChain &y) { return x.w > y.w; });
    std::vector<int> chains;
    a[0].kept = 3; chains.push_back(0);
    auto cbeg = [](const OChain &c) { return c.seeds.front().qbeg; };
    auto cend = [](const OChain &c) { return c.seeds.back().qbeg + c.seeds.back().len; };
    for (int i = 1; i < n; ++i) {
        int large_ovlp = 0; size_t k;
        for (k = 0; k < chains.size(); ++k) {
            int j = chains[k];
            int b_max = cbeg(a[j]) > cbeg(a[i]) ? cbeg(a[j]) : cbeg(a[i]);
            int e_min = cend(a[j]) < cend(a[i]) ? cend(a[j]) : cend(a[i]);
            if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
                int li = cend(a[i]) - cbeg(a[i]), lj = cend(a[j]) - cbeg(a[j]);
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
                    large_ovlp = 1;
                    if (a[j].first < 0) a[j].first = i;
                    if (a[i].w < a[j].w * opt->drop_ratio && a[j].w - a[i].w >= opt->min_seed_len << 1) break;
                }
            }
        }
        if (k == chains.size()) { chains.push_back(i); a[i].kept = large_ovlp ? 2 : 3; }
    }
    for (int idx : chains) if (a[idx].first >= 0) a[a[idx].first].kept = 1;
    int i, k;
    for (i = k = 0; i < n; ++i) {
        if (a[i].kept == 0 || a[i].kept == 3) continue;
        if (++k >= opt->max_chain_extend) break;
    }
    for (; i < n; ++i) if (a[i].kept < 3) a[i].kept = 0;
    std::vector<OChain> out;
    for (OChain &c : a) if (c.kept) out.push_back(c);
    a.swap(out);
}

/* The chain tree (KBTREE_INIT(chn, mem_chain_t, chain_cmp), src/bwamem.cpp:40-41; src/kbtree.h).  Keys with equal pos are
 * legal in it, and which of them a lookup meets first -- and next to which of them a new one is placed -- depends on the
 * shape of the tree, so the tree is restated for real: order t from kb_init's arithmetic (src/kbtree.h:64) with the 48-byte
 * mem_chain_t (src/bwamem.h:126-133) and size KB_DEFAULT_SIZE + 8 = 520 (src/bwamem.cpp:845), top-down insertion that
 * splits every full node on the way (src/kbtree.h:181-235), lookups as kb_intervalp (src/kbtree.h:158-175).
 * Keys are indices into the chain pool. */
struct ChainTree {
    static const int T = (int) (((520 - 4 - 8) / (8 + 48) + 1) >> 1);          /* 5: at most 9 keys per node */
    struct Node { bool internal; std::vector<int> key; std::vector<int> child; };
    std::vector<Node> nodes; int root; long n_keys;
    const std::vector<OChain> &pool;
    explicit ChainTree(const std::vector<OChain> &pool_) : root(0), n_keys(0), pool(pool_) { nodes.push_back(Node{false, {}, {}}); }
    /* __kb_getp_aux (src/kbtree.h:125-139): index of the first key >= pos if it is equal (*r = 0), else the index before it
     * (*r = -1); the last index when every key is smaller (*r = 1); -1 for an empty node */
    int slot(const Node &x, int64_t pos, int *r) const {
        const int n = (int) x.key.size();
        if (n == 0) return -1;
        int b = 0, e = n;
        while (b < e) { const int mid = (b + e) >> 1; if (pool[x.key[mid]].pos < pos) b = mid + 1; else e = mid; }
        if (b == n) { *r = 1; return n - 1; }
        *r = pos < pool[x.key[b]].pos ? -1 : 0;
        return *r < 0 ? b - 1 : b;
    }
    /* kb_intervalp: the `lower` key (pool index) or -1 */
    int lower(int64_t pos) const {
        int low = -1, r = 0;
        for (int x = root;;) {
            const Node &nd = nodes[x];
            const int i = slot(nd, pos, &r);
            if (i >= 0 && r == 0) return nd.key[i];
            if (i >= 0) low = nd.key[i];
            if (!nd.internal) return low;
            x = nd.child[i + 1];
        }
    }
    /* __kb_split: child number i of x (full) gives its median key to x and its upper half to a new sibling */
    void split(int x, int i) {
        const int y = nodes[x].child[i];
        Node z; z.internal = nodes[y].internal;
        z.key.assign(nodes[y].key.begin() + T, nodes[y].key.end());
        if (z.internal) z.child.assign(nodes[y].child.begin() + T, nodes[y].child.end());
        const int median = nodes[y].key[T - 1];
        nodes[y].key.resize(T - 1);
        if (z.internal) nodes[y].child.resize(T);
        nodes.push_back(z);
        const int zi = (int) nodes.size() - 1;
        nodes[x].child.insert(nodes[x].child.begin() + i + 1, zi);
        nodes[x].key.insert(nodes[x].key.begin() + i, median);
    }
    /* kb_putp */
    void put(int id) {
        const int64_t pos = pool[id].pos;
        ++n_keys;
        if ((int) nodes[root].key.size() == 2 * T - 1) {
            Node s; s.internal = true; s.child.push_back(root);
            nodes.push_back(s);
            root = (int) nodes.size() - 1;
            split(root, 0);
        }
        int x = root, r = 0;
        while (nodes[x].internal) {
            int i = slot(nodes[x], pos, &r) + 1;
            if ((int) nodes[nodes[x].child[i]].key.size() == 2 * T - 1) {
                split(x, i);
                if (pos > pool[nodes[x].key[i]].pos) ++i;
            }
            x = nodes[x].child[i];
        }
        const int i = slot(nodes[x], pos, &r);
        nodes[x].key.insert(nodes[x].key.begin() + (i + 1), id);
    }
    void in_order(int x, std::vector<int> &out) const {
        const Node &nd = nodes[x];
        for (size_t i = 0; i < nd.key.size(); ++i) { if (nd.internal) in_order(nd.child[i], out); out.push_back(nd.key[i]); }
        if (nd.internal) in_order(nd.child.back(), out);
    }
};

/* A4. chains of one read from its ordered SMEMs (mem_chain_seeds, src/bwamem.cpp:806-974) */
static void chain_read(const Fm &fm, const bm2_mem_opt_t *opt, const bm2_smem *sm, int64_t nsm, int l_seq, int seqid,
                       std::vector<OChain> &chains)
{
    const bm2_index_desc *x = fm.x;
    int b = 0, e = 0, l_rep = 0;
    for (int64_t i = 0; i < nsm; ++i) {
        int sb = sm[i].m, se = sm[i].n + 1;
        if (sm[i].s <= opt->max_occ) continue;
        if (sb > e) { l_rep += e - b; b = sb; e = se; }
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    std::vector<OChain> pool;                    /* in creation order */
    ChainTree tree(pool);
    for (int64_t i = 0; i < nsm; ++i) {
        const bm2_smem &p = sm[i];
        int slen = p.n + 1 - p.m;
        int64_t step = p.s > opt->max_occ ? p.s / opt->max_occ : 1;
        int64_t k; int count;
        for (k = 0, count = 0; k < p.s && count < opt->max_occ; k += step, ++count) {
            OSeed s; s.rbeg = sa_of_row(fm, p.k + k); s.qbeg = p.m; s.len = s.score = slen;
            int rid = intv2rid(x, s.rbeg, s.rbeg + s.len);
            if (rid < 0) continue;
            bool to_add = true;
            if (tree.n_keys) {
                const int lower = tree.lower(s.rbeg);
                if (lower >= 0 && test_and_merge(opt, x->l_pac, pool[lower], s, rid)) to_add = false;
            }
            if (to_add) {
                OChain c; c.pos = s.rbeg; c.rid = rid; c.seqid = seqid; c.is_alt = x->ann_is_alt ? !!x->ann_is_alt[rid] : 0;
                c.w = 0; c.kept = 0; c.first = -1; c.frac_rep = 0; c.seeds.push_back(s);
                pool.push_back(c);
                tree.put((int) pool.size() - 1);
            }
        }
    }
    std::vector<int> order;
    tree.in_order(tree.root, order);
    chains.clear(); chains.reserve(order.size());
    for (int id : order) { chains.push_back(std::move(pool[id])); chains.back().frac_rep = (float) l_rep / l_seq; }
}

}  // namespace

namespace { void flt_chained_seeds(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int l_query, const uint8_t *query, std::vector<OChain> &chains); }

static const int BLOCK_READS = 512;   /* BATCH_SIZE, src/macro.h:48 */

static void collect_all(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                        std::vector<bm2_smem> &all)
{
    Fm fm = { idx };
    for (int b0 = 0; b0 < reads->n_reads; b0 += BLOCK_READS) {
        size_t first = all.size();
        int b1 = std::min(reads->n_reads, b0 + BLOCK_READS);
        for (int r = b0; r < b1; ++r)
            collect_read(fm, opt, reads->codes + reads->offsets[r], (int)(reads->offsets[r + 1] - reads->offsets[r]), r, all);
        (void) first;
    }
}

extern "C" int64_t bm2o_collect_smems(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                                      bm2_smem **out)
{
    std::vector<bm2_smem> all;
    collect_all(idx, opt, reads, all);
    *out = (bm2_smem *) malloc(sizeof(bm2_smem) * (all.size() + 1));
    memcpy(*out, all.data(), sizeof(bm2_smem) * all.size());
    return (int64_t) all.size();
}

extern "C" void bm2o_sa_lookup(const bm2_index_desc *idx, const int64_t *rows, int64_t n, int64_t *out) {
    Fm fm = { idx };
    for (int64_t i = 0; i < n; ++i) out[i] = sa_of_row(fm, rows[i]);
}

extern "C" int bm2o_seed_chain(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                               bm2_chain **chains, int64_t *n_chains, bm2_seed **seeds, int64_t *n_seeds, int64_t **read_off)
{
    Fm fm = { idx };
    std::vector<bm2_chain> oc; std::vector<bm2_seed> os;
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (reads->n_reads + 1));
    off[0] = 0;
    for (int b0 = 0; b0 < reads->n_reads; b0 += BLOCK_READS) {
        int b1 = std::min(reads->n_reads, b0 + BLOCK_READS);
        std::vector<bm2_smem> sm;
        for (int r = b0; r < b1; ++r)
            collect_read(fm, opt, reads->codes + reads->offsets[r], (int)(reads->offsets[r + 1] - reads->offsets[r]), r, sm);
        /* reference quirk: a 512-read block whose SMEM total is exactly 1 yields no chain
         * (loop guards `pos < num_smem - 1`, src/bwamem.cpp:835) */
        bool skip_block = sm.size() <= 1;
        size_t p = 0;
        for (int r = b0; r < b1; ++r) {
            size_t q = p;
            while (q < sm.size() && (int) sm[q].rid == r) ++q;
            std::vector<OChain> ch;
            int l_seq = (int)(reads->offsets[r + 1] - reads->offsets[r]);
            if (!skip_block && q > p && l_seq >= opt->min_seed_len) {
                chain_read(fm, opt, sm.data() + p, (int64_t)(q - p), l_seq, r, ch);
                chain_filter(opt, ch);
                flt_chained_seeds(idx, opt, l_seq, reads->codes + reads->offsets[r], ch);
            }
            for (OChain &c : ch) {
                bm2_chain o; memset(&o, 0, sizeof(o));
                o.pos = c.pos; o.seqid = c.seqid; o.rid = c.rid; o.n_seeds = (int32_t) c.seeds.size(); o.seed_off = (int32_t) os.size();
                o.w = c.w; o.kept = c.kept; o.first = c.first; o.is_alt = c.is_alt; o.frac_rep = c.frac_rep;
                for (OSeed &s : c.seeds) { bm2_seed t; t.rbeg = s.rbeg; t.qbeg = s.qbeg; t.len = s.len; t.score = s.score; t.chain = (int32_t) oc.size(); os.push_back(t); }
                oc.push_back(o);
            }
            off[r + 1] = (int64_t) oc.size();
            p = q;
        }
    }
    *chains = (bm2_chain *) malloc(sizeof(bm2_chain) * (oc.size() + 1)); memcpy(*chains, oc.data(), sizeof(bm2_chain) * oc.size());
    *seeds = (bm2_seed *) malloc(sizeof(bm2_seed) * (os.size() + 1)); memcpy(*seeds, os.data(), sizeof(bm2_seed) * os.size());
    *n_chains = (int64_t) oc.size(); *n_seeds = (int64_t) os.size(); *read_off = off;
    return 0;
}

/* ================================================================================================
 * Extension stage (A5-A8) and the tail of mem_kernel2_core
 * ============================================================================================== */
namespace {

/* cal_max_gap (src/bwamem.cpp:66-76) */
static int cal_max_gap(const bm2_mem_opt_t *opt, int qlen) {
    int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
    int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < opt->w << 1 ? l : opt->w << 1;
}

/* ksw_global2, score only (src/ksw.cpp:558-668) */
static int global_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                        int o_del, int e_del, int o_ins, int e_ins, int w)
{
    const int MINUS_INF = -0x40000000;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    std::vector<int32_t> H(qlen + 1), E(qlen + 1);
    H[0] = 0; E[0] = MINUS_INF;
    int j;
    for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = MINUS_INF; }
    for (; j <= qlen; ++j) H[j] = E[j] = MINUS_INF;
    for (int i = 0; i < tlen; ++i) {
        int32_t f = MINUS_INF, h1;
        int beg = i > w ? i - w : 0;
        int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : MINUS_INF;
        for (j = beg; j < end; ++j) {
            int32_t m = H[j], e = E[j];
            H[j] = h1;
            m += mat[target[i] * 5 + query[j]];
            int32_t h = m >= e ? m : e;
            h = h >= f ? h : f;
            h1 = h;
            int32_t t = m - oe_del;
            e -= e_del; e = e > t ? e : t;
            E[j] = e;
            t = m - oe_ins;
            f -= e_ins; f = f > t ? f : t;
        }
        H[end] = h1; E[end] = MINUS_INF;
    }
    return H[qlen];
}

/* bwa_gen_cigar2 with n_cigar == NM == NULL: score of the banded global alignment of
 * query[0..l_query) against T[rb..re) (src/bwa.cpp:260-347); returns false when rejected. */
static bool gen_score(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int w_, int l_query, const uint8_t *query,
                      int64_t rb, int64_t re, int *score)
{
    int64_t l_pac = x->l_pac;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
    if (re > (l_pac << 1)) return false;       /* bns_get_seq clips -> re-rb != rlen -> no score */
    if (rb < 0) return false;
    int64_t rlen = re - rb;
    std::vector<uint8_t> rs(x->ref_string + rb, x->ref_string + re), qs(query, query + l_query);
    if (rb >= l_pac) { std::reverse(rs.begin(), rs.end()); std::reverse(qs.begin(), qs.end()); }
    if (l_query == re - rb && w_ == 0) {
        int sc = 0;
        for (int i = 0; i < l_query; ++i) sc += opt->mat[rs[i] * 5 + qs[i]];
        *score = sc;
    } else {
        int max_ins = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_ins) / opt->e_ins + 1.);
        int max_del = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_del) / opt->e_del + 1.);
        int max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        int diff = (int)(rlen - l_query); if (diff < 0) diff = -diff;
        int w = (max_gap + diff + 1) >> 1;
        w = w < w_ ? w : w_;
        int min_w = diff + 3;
        w = w > min_w ? w : min_w;
        *score = global_score(l_query, qs.data(), (int) rlen, rs.data(), opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w);
    }
    return true;
}

/* mem_patch_reg (src/bwamem.cpp:175-234) */
static int patch_reg(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const uint8_t *query, const bm2_alnreg_t *a,
                     const bm2_alnreg_t *b, int *_w)
{
    int w, score = 0, q_s, r_s;
    double r;
    if (query == 0) return 0;                                  /* mem_patch_reg: bns == 0 || pac == 0 || query == 0 (src/bwamem.cpp:179) */
    if (a->rb < x->l_pac && b->rb >= x->l_pac) return 0;
    if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
    w = (int)((a->re - b->rb) - (a->qe - b->qb));
    w = w > 0 ? w : -w;
    r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
    r = r > 0. ? r : -r;
    if (a->re < b->rb || a->qe < b->qb) {
        if (w > opt->w << 1 || r >= 0.05f) return 0;
    } else if (w > opt->w << 2 || r >= 0.05f * 2) return 0;
    w += a->w + b->w;
    w = w < opt->w << 2 ? w : opt->w << 2;
    if (!gen_score(x, opt, w, b->qe - a->qb, query + a->qb, a->rb, b->re, &score)) score = 0;
    q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
    r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
    if ((double) score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
    *_w = w;
    return score;
}

static inline int reg_n_comp(const bm2_alnreg_t &a) { return (a.n_comp_is_alt << 2) >> 2; }
static inline void reg_set_n_comp(bm2_alnreg_t &a, int v) { a.n_comp_is_alt = (a.n_comp_is_alt & ~0x3FFFFFFF) | (v & 0x3FFFFFFF); }
static inline void reg_set_is_alt(bm2_alnreg_t &a, int v) { a.n_comp_is_alt = (a.n_comp_is_alt & 0x3FFFFFFF) | ((v & 3) << 30); }

/* mem_sort_dedup_patch (src/bwamem.cpp:292-353) */
static int sort_dedup_patch(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const uint8_t *query, int n, bm2_alnreg_t *a) {
    int m, i, j;
    if (n <= 1) return n;
    ks_introsort(a, n, [](const bm2_alnreg_t &p, const bm2_alnreg_t &q) { return p.re < q.re; });
    for (i = 0; i < n; ++i) reg_set_n_comp(a[i], 1);
    for (i = 1; i < n; ++i) {
        bm2_alnreg_t *p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt->max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
            bm2_alnreg_t *q = &a[j];
            int64_t or_, oq, mr, mq;
            int score, w;
            if (q->qe == q->qb) continue;
            or_ = q->re - p->rb;
            oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
            mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if (or_ > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (q->rb < p->rb && (score = patch_reg(x, opt, query, q, p, &w)) > 0) {
                reg_set_n_comp(*p, reg_n_comp(*p) + reg_n_comp(*q) + 1);
                p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
                p->sub = p->sub > q->sub ? p->sub : q->sub;
                p->csub = p->csub > q->csub ? p->csub : q->csub;
                p->qb = q->qb; p->rb = q->rb;
                p->truesc = p->score = score;
                p->w = w;
                q->qb = q->qe;
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    n = m;
    ks_introsort(a, n, [](const bm2_alnreg_t &p, const bm2_alnreg_t &q) {
        return p.score > q.score || (p.score == q.score && (p.rb < q.rb || (p.rb == q.rb && p.qb < q.qb)));
    });
    for (i = 1; i < n; ++i)
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    return m;
}

struct ExtJob { int reg; int qlen, tlen; int64_t toff; int qoff; int dir; };   /* dir -1: left (reversed), +1: right */

/* mem_chain2aln_across_reads_V2 for one read (src/bwamem.cpp:2069-2994) */
static void extend_read(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const uint8_t *query, int l_query,
                        std::vector<OChain> &chains, std::vector<bm2_alnreg_t> &av, int64_t *cells)
{
    const int64_t l_pac = x->l_pac;
    const int H0_ = -99;
    struct SeedRef { int chain, seed, aln; };
    std::vector<SeedRef> order;              /* replay order (A5) */
    std::vector<ExtJob> left, right;
    std::vector<int> reg_chain;
    for (size_t ci = 0; ci < chains.size(); ++ci) {
        OChain &c = chains[ci];
        if (c.seeds.empty()) continue;
        int64_t rmax0 = l_pac << 1, rmax1 = 0;
        for (const OSeed &t : c.seeds) {
            int64_t b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
            int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
            rmax0 = rmax0 < b ? rmax0 : b;
            rmax1 = rmax1 > e ? rmax1 : e;
        }
        rmax0 = rmax0 > 0 ? rmax0 : 0;
        rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
        if (rmax0 < l_pac && l_pac < rmax1) { if (c.seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
        {   /* bns_fetch_seq_v2: clip to the contig of seeds[0].rbeg (src/bwamem.cpp:1890-1924) */
            int64_t mid = c.seeds[0].rbeg;
            int is_rev = mid >= l_pac;
            int rid = pos2rid(x, depos(x, mid));
            int64_t far_beg = x->ann_offset[rid], far_end = far_beg + x->ann_len[rid];
            if (is_rev) { int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
            rmax0 = rmax0 > far_beg ? rmax0 : far_beg;
            rmax1 = rmax1 < far_end ? rmax1 : far_end;
        }
        int n = (int) c.seeds.size();
        std::vector<uint64_t> srt(n);
        for (int i = 0; i < n; ++i) srt[i] = (uint64_t) c.seeds[i].score << 32 | (uint32_t) i;
        std::sort(srt.begin(), srt.end());       /* keys are unique: ks_introsort_64 gives the same order */
        for (int k = n - 1; k >= 0; --k) {
            int si = (int)(uint32_t) srt[k];
            const OSeed &s = c.seeds[si];
            bm2_alnreg_t a; memset(&a, 0, sizeof(a));
            int ai = (int) av.size();
            a.w = opt->w; a.score = a.truesc = -1; a.rid = c.rid; a.frac_rep = c.frac_rep; a.seedlen0 = s.len;
            a.rb = a.re = H0_; a.qb = a.qe = H0_;
            if (s.qbeg) {
                ExtJob j; j.reg = ai; j.qlen = s.qbeg; j.tlen = (int)(s.rbeg - rmax0); j.toff = s.rbeg - 1; j.qoff = s.qbeg - 1; j.dir = -1;
                left.push_back(j);
                a.qb = s.qbeg; a.rb = s.rbeg;
            } else { a.score = a.truesc = s.len * opt->a; a.qb = 0; a.rb = s.rbeg; }
            bool has_right = false;
            if (s.qbeg + s.len != l_query) {
                int64_t qe = s.qbeg + s.len, re = s.rbeg + s.len - rmax0;
                ExtJob j; j.reg = ai; j.qlen = (int)(l_query - qe); j.tlen = (int)(rmax1 - rmax0 - re); j.toff = rmax0 + re; j.qoff = (int) qe; j.dir = 1;
                right.push_back(j);
                a.qe = (int) qe; a.re = rmax0 + re;
                has_right = true;
            } else { a.qe = l_query; a.re = s.rbeg + s.len; }
            av.push_back(a); reg_chain.push_back((int) ci);
            if (!has_right && av[ai].rb != H0_ && av[ai].qb != H0_) {
                int cov = 0;
                for (const OSeed &t : c.seeds)
                    if (t.qbeg >= av[ai].qb && t.qbeg + t.len <= av[ai].qe && t.rbeg >= av[ai].rb && t.rbeg + t.len <= av[ai].re) cov += t.len;
                av[ai].seedcov = cov;
            }
            SeedRef sr = { (int) ci, si, ai }; order.push_back(sr);
        }
    }
    auto seedcov = [&](bm2_alnreg_t &a, const OChain &c) {
        if (a.rb != H0_ && a.qb != H0_ && a.qe != H0_ && a.re != H0_) {
            int cov = 0;
            for (const OSeed &t : c.seeds)
                if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) cov += t.len;
            a.seedcov = cov;
        }
    };
    auto run = [&](const ExtJob &j, int h0, int w, int end_bonus, int32_t *o) {
        std::vector<uint8_t> q(j.qlen), t(j.tlen);
        for (int i = 0; i < j.qlen; ++i) q[i] = query[j.qoff + (int64_t) i * j.dir];
        for (int i = 0; i < j.tlen; ++i) t[i] = x->ref_string[j.toff + (int64_t) i * j.dir];
        bm2o_bsw_params p = { opt->a, opt->b, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, end_bonus, 1 };
        /* the reference routes jobs with 32-bit scores to the scalar kernel (src/bwamem.cpp:2304-2313) */
        int minlen = j.qlen < j.tlen ? j.qlen : j.tlen;
        if (!(j.tlen < 32768 && j.qlen < 32768 && h0 + minlen * opt->a < 32768)) p.vector_quirks = 0;
        *cells += bm2o_bsw_extend(q.data(), j.qlen, t.data(), j.tlen, w, h0, &p, o);
    };
    /* A7 left */
    for (const ExtJob &j : left) {
        bm2_alnreg_t &a = av[j.reg];
        const OChain &c = chains[reg_chain[j.reg]];
        int h0 = a.seedlen0 * opt->a;
        for (int i = 0; i < 2; ++i) {
            int w = opt->w << i; int32_t o[6];
            run(j, h0, w, opt->pen_clip5, o);
            int score = o[0], qle = o[1], tle = o[2], gtle = o[3], gscore = o[4], max_off = o[5];
            int prev = a.score; a.score = score;
            if (a.score == prev || max_off < (w >> 1) + (w >> 2) || i + 1 == 2) {
                if (gscore <= 0 || gscore <= a.score - opt->pen_clip5) { a.qb -= qle; a.rb -= tle; a.truesc = a.score; }
                else { a.qb = 0; a.rb -= gtle; a.truesc = gscore; }
                a.w = a.w > w ? a.w : w;
                seedcov(a, c);
                break;
            }
        }
    }
    /* A7 right */
    for (const ExtJob &j : right) {
        bm2_alnreg_t &a = av[j.reg];
        const OChain &c = chains[reg_chain[j.reg]];
        int h0 = a.score;
        for (int i = 0; i < 2; ++i) {
            int w = opt->w << i; int32_t o[6];
            run(j, h0, w, opt->pen_clip3, o);
            int score = o[0], qle = o[1], tle = o[2], gtle = o[3], gscore = o[4], max_off = o[5];
            int prev = a.score; a.score = score;
            if (a.score == prev || max_off < (w >> 1) + (w >> 2) || i + 1 == 2) {
                if (gscore <= 0 || gscore <= a.score - opt->pen_clip3) { a.qe += qle; a.re += tle; a.truesc += a.score - h0; }
                else { a.qe = l_query; a.re += gtle; a.truesc += gscore - h0; }
                a.w = a.w > w ? a.w : w;
                seedcov(a, c);
                break;
            }
        }
    }
    /* A8 post-filter (src/bwamem.cpp:2895-2989) */
    int lim = 0;
    size_t pos = 0;
    for (size_t ci = 0; ci < chains.size(); ++ci) {
        OChain &c = chains[ci];
        int n = (int) c.seeds.size();
        if (n == 0) continue;
        /* srt2[k] for k = n-1..0 is order[pos + (n-1-k)] */
        std::vector<int> srt2(n); std::vector<int> aln_of(n);
        for (int k = n - 1; k >= 0; --k) { srt2[k] = order[pos + (n - 1 - k)].seed; aln_of[k] = order[pos + (n - 1 - k)].aln; }
        pos += n;
        for (int k = n - 1; k >= 0; --k) {
            const OSeed &s = c.seeds[srt2[k]];
            int i, v = 0;
            for (i = 0; i < (int) av.size() && v < lim; ++i) {
                bm2_alnreg_t *p = &av[i];
                if (p->qb == -1 && p->qe == -1) continue;
                int64_t rd; int qd, w, max_gap;
                if (s.rbeg < p->rb || s.rbeg + s.len > p->re || s.qbeg < p->qb || s.qbeg + s.len > p->qe) { v++; continue; }
                if (s.len - p->seedlen0 > .1 * l_query) { v++; continue; }
                qd = s.qbeg - p->qb; rd = s.rbeg - p->rb;
                max_gap = cal_max_gap(opt, qd < rd ? qd : (int) rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
                qd = p->qe - (s.qbeg + s.len); rd = p->re - (s.rbeg + s.len);
                max_gap = cal_max_gap(opt, qd < rd ? qd : (int) rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
                v++;
            }
            if (v < lim) {
                int vv;
                for (vv = k + 1; vv < n; ++vv) {
                    if (srt2[vv] < 0) continue;
                    const OSeed &t = c.seeds[srt2[vv]];
                    if (t.len < s.len * .95) continue;
                    if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                    if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
                }
                if (vv == n) {
                    av[aln_of[k]].qb = av[aln_of[k]].qe = -1;
                    srt2[k] = -1;
                    continue;
                }
            }
            lim++;
        }
    }
}


/* local Smith-Waterman score == kswr_t::score of ksw_align2 / ksw_i16 (src/ksw.cpp:234-345): H floored at 0,
 * E(i+1,j) = max(E - e_del, H - oe_del), F(i,j+1) = max(F - e_ins, H - oe_ins) with unsigned saturation.
 * (Farrar's lazy-F pass does not feed F-corrected H back into E; that only matters for scoring systems
 * where an insertion next to a deletion beats a mismatch, which bwa's presets exclude.) */
static int local_sw_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                          int o_del, int e_del, int o_ins, int e_ins)
{
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    std::vector<int> H(qlen + 1, 0), E(qlen + 1, 0);
    int gmax = 0;
    for (int i = 0; i < tlen; ++i) {
        int f = 0, diag = 0;                       /* H(i-1, -1) = 0 */
        for (int j = 0; j < qlen; ++j) {
            int h = diag + mat[target[i] * 5 + query[j]];
            diag = H[j + 1];
            int e = E[j + 1];
            if (e > h) h = e;
            if (f > h) h = f;
            if (h < 0) h = 0;
            if (h > gmax) gmax = h;
            H[j + 1] = h;
            int t = h - oe_del; if (t < 0) t = 0;
            e -= e_del; if (e < 0) e = 0;
            E[j + 1] = e > t ? e : t;
            t = h - oe_ins; if (t < 0) t = 0;
            f -= e_ins; if (f < 0) f = 0;
            f = f > t ? f : t;
        }
    }
    return gmax;
}

/* mem_seed_sw (src/bwamem.cpp:401-427) */
static int seed_sw(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int l_query, const uint8_t *query, const OSeed &s) {
    const int64_t l_pac = x->l_pac;
    if (s.len >= 200) return -1;                               /* MEM_SHORT_LEN */
    int qb = s.qbeg, qe = s.qbeg + s.len;
    int64_t rb = s.rbeg, re = s.rbeg + s.len, mid = (rb + re) >> 1;
    qb -= 50; qb = qb > 0 ? qb : 0;                             /* MEM_SHORT_EXT */
    qe += 50; qe = qe < l_query ? qe : l_query;
    rb -= 50; rb = rb > 0 ? rb : 0;
    re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
    if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
    if (qe - qb >= 200 || re - rb >= 200) return -1;
    {   /* bns_fetch_seq (src/bntseq.cpp:453-482): clip to the contig of mid */
        int is_rev = mid >= l_pac;
        int rid = pos2rid(x, depos(x, mid));
        int64_t far_beg = x->ann_offset[rid], far_end = far_beg + x->ann_len[rid];
        if (is_rev) { int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
        rb = rb > far_beg ? rb : far_beg;
        re = re < far_end ? re : far_end;
    }
    return local_sw_score(qe - qb, query + qb, (int)(re - rb), x->ref_string + rb, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins);
}

/* mem_flt_chained_seeds (src/bwamem.cpp:472-504) for the chains of one read */
void flt_chained_seeds(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int l_query, const uint8_t *query, std::vector<OChain> &chains) {
    double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log((double) l_query);
    int min_HSP_score = (int)(opt->a * min_l + .499);
    if (min_l > 0.05f * l_query) return;
    for (OChain &c : chains) {
        std::vector<OSeed> kept;
        for (OSeed s : c.seeds) {
            s.score = seed_sw(x, opt, l_query, query, s);
            if (s.score < 0 || s.score >= min_HSP_score) {
                s.score = s.score < 0 ? s.len * opt->a : s.score;
                kept.push_back(s);
            }
        }
        c.seeds.swap(kept);
    }
}

}  // namespace

extern "C" int bm2o_seed_chain_extend(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads,
                                      bm2_alnreg_t **regs, int64_t *n_regs, int64_t **read_off, int64_t *bsw_cells)
{
    Fm fm = { idx };
    std::vector<bm2_alnreg_t> all;
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (reads->n_reads + 1));
    off[0] = 0;
    int64_t cells = 0;
    int rc = 0;
    for (int b0 = 0; b0 < reads->n_reads; b0 += BLOCK_READS) {
        int b1 = std::min(reads->n_reads, b0 + BLOCK_READS);
        std::vector<bm2_smem> sm;
        for (int r = b0; r < b1; ++r)
            collect_read(fm, opt, reads->codes + reads->offsets[r], (int)(reads->offsets[r + 1] - reads->offsets[r]), r, sm);
        bool skip_block = sm.size() <= 1;
        size_t p = 0;
        for (int r = b0; r < b1; ++r) {
            size_t q = p;
            while (q < sm.size() && (int) sm[q].rid == r) ++q;
            std::vector<OChain> ch;
            const uint8_t *query = reads->codes + reads->offsets[r];
            int l_seq = (int)(reads->offsets[r + 1] - reads->offsets[r]);
            if (!skip_block && q > p && l_seq >= opt->min_seed_len) {
                chain_read(fm, opt, sm.data() + p, (int64_t)(q - p), l_seq, r, ch);
                chain_filter(opt, ch);
                flt_chained_seeds(idx, opt, l_seq, query, ch);
            }
            std::vector<bm2_alnreg_t> av;
            extend_read(idx, opt, query, l_seq, ch, av, &cells);
            int m = 0;
            for (size_t i = 0; i < av.size(); ++i) if (av[i].qe > av[i].qb) av[m++] = av[i];
            m = sort_dedup_patch(idx, opt, query, m, av.data());
            for (int i = 0; i < m; ++i) {
                if (av[i].rid >= 0 && idx->ann_is_alt && idx->ann_is_alt[av[i].rid]) reg_set_is_alt(av[i], 1);
                all.push_back(av[i]);
            }
            off[r + 1] = (int64_t) all.size();
            p = q;
        }
    }
    *regs = (bm2_alnreg_t *) malloc(sizeof(bm2_alnreg_t) * (all.size() + 1));
    memcpy(*regs, all.data(), sizeof(bm2_alnreg_t) * all.size());
    *n_regs = (int64_t) all.size(); *read_off = off;
    if (bsw_cells) *bsw_cells = cells;
    return rc;
}


/* ================================================================================================
 * CIGAR / NM / MD: bwa_gen_cigar2 (src/bwa.cpp:260-347) over ksw_global2 with the
 * backtrack matrix (src/ksw.cpp:558-668) and push_cigar (:545-556).  Plain restatement: sequences are
 * copied (and reversed for reverse-strand hits), the matrix is a vector.
 * ============================================================================================== */
static void o_push_cigar(std::vector<uint32_t> &c, int op, int len) {
    if (c.empty() || op != (int) (c.back() & 0xf)) c.push_back((uint32_t) len << 4 | (uint32_t) op);
    else c.back() += (uint32_t) len << 4;
}

static int o_global_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                          int o_del, int e_del, int o_ins, int e_ins, int w, std::vector<uint32_t> &cigar)
{
    const int MINUS_INF = -0x40000000;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    std::vector<uint8_t> z((size_t) n_col * (size_t) tlen + 1);
    std::vector<int32_t> H(qlen + 1), E(qlen + 1);
    H[0] = 0; E[0] = MINUS_INF;
    int i, j, k;
    for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = MINUS_INF; }
    for (; j <= qlen; ++j) H[j] = E[j] = MINUS_INF;
    for (i = 0; i < tlen; ++i) {
        int32_t f = MINUS_INF, h1;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : MINUS_INF;
        uint8_t *zi = &z[(size_t) i * n_col];
        for (j = beg; j < end; ++j) {
            int32_t m = H[j], e = E[j], h, t;
            uint8_t d;
            H[j] = h1;
            m += mat[target[i] * 5 + query[j]];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del; e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            E[j] = e;
            t = m - oe_ins; f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        H[end] = h1; E[end] = MINUS_INF;
    }
    const int score = H[qlen];
    int which = 0;
    cigar.clear();
    i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
        which = z[(size_t) i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
        if (which == 0) { o_push_cigar(cigar, 0, 1); --i; --k; }
        else if (which == 1) { o_push_cigar(cigar, 2, 1); --i; }
        else { o_push_cigar(cigar, 1, 1); --k; }
    }
    if (i >= 0) o_push_cigar(cigar, 2, i + 1);
    if (k >= 0) o_push_cigar(cigar, 1, k + 1);
    std::reverse(cigar.begin(), cigar.end());
    return score;
}

/* bwa_gen_cigar2 for one alignment: returns false when rejected (score untouched, n_cigar = 0, NM = -1) */
static bool gen_cigar_one(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int w_, int l_query, const uint8_t *query, int64_t rb, int64_t re,
                          int *score, std::vector<uint32_t> &cig, int *NM, std::string &md)
{
    const int64_t l_pac = x->l_pac;
    cig.clear(); md.clear(); *NM = -1;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;     /* src/bwa.cpp:272 */
    if (re > (l_pac << 1) || rb < 0) return false;                                 /* bns_get_seq clips: rlen != re - rb, :274 */
    const int64_t rlen = re - rb;
    std::vector<uint8_t> rs(x->ref_string + rb, x->ref_string + re), qs(query, query + l_query);
    const bool rev = rb >= l_pac;
    if (rev) { std::reverse(rs.begin(), rs.end()); std::reverse(qs.begin(), qs.end()); }      /* :275-280 */
    if (l_query == rlen && w_ == 0) {                                                          /* :281-290 */
        cig.push_back((uint32_t) l_query << 4 | 0);
        int sc = 0;
        for (int i = 0; i < l_query; ++i) sc += opt->mat[rs[i] * 5 + qs[i]];
        *score = sc;
    } else {                                                                                    /* :291-304 */
        int max_ins = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_ins) / opt->e_ins + 1.);
        int max_del = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_del) / opt->e_del + 1.);
        int max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        int diff = (int)(rlen - l_query); if (diff < 0) diff = -diff;
        int w = (max_gap + diff + 1) >> 1;
        w = w < w_ ? w : w_;
        int min_w = diff + 3;
        w = w > min_w ? w : min_w;
        *score = o_global_align(l_query, qs.data(), (int) rlen, rs.data(), opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, cig);
    }
    /* NM and MD (:305-337) */
    int xq = 0, y = 0, u = 0, n_mm = 0, n_gap = 0;
    const char *int2base = rev ? "TGCAN" : "ACGTN";
    const int nc = (int) cig.size();
    for (int k = 0; k < nc; ++k) {
        const int op = (int) (cig[k] & 0xf), len = (int) (cig[k] >> 4);
        if (op == 0) {
            for (int i = 0; i < len; ++i) {
                if (qs[xq + i] != rs[y + i]) { md += std::to_string(u); md += int2base[rs[y + i]]; ++n_mm; u = 0; }
                else ++u;
            }
            xq += len; y += len;
        } else if (op == 2) {
            if (k > 0 && k < nc - 1) {
                md += std::to_string(u); md += '^';
                for (int i = 0; i < len; ++i) md += int2base[rs[y + i]];
                u = 0; n_gap += len;
            }
            y += len;
        } else if (op == 1) { xq += len; n_gap += len; }
    }
    md += std::to_string(u);
    *NM = n_mm + n_gap;
    return true;
}

extern "C" int bm2o_gen_cigar(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_cigar_req *reqs,
                              int64_t n, bm2_cigar_rec **recs_out, uint32_t **cigar_out, int64_t *n_ops_out, char **md_out, int64_t *n_md_out)
{
    std::vector<bm2_cigar_rec> recs((size_t) n);
    std::vector<uint32_t> all_ops; std::string all_md;
    for (int64_t r = 0; r < n; ++r) {
        const bm2_cigar_req &q = reqs[r];
        bm2_cigar_rec &o = recs[(size_t) r];
        o.score = INT32_MIN; o.n_cigar = 0; o.nm = -1; o.n_md = 0; o.cigar_off = (int64_t) all_ops.size(); o.md_off = (int64_t) all_md.size();
        if (q.read < 0 || q.read >= reads->n_reads) return 1;
        const int64_t ro = reads->offsets[q.read], rl = reads->offsets[q.read + 1] - ro;
        if (q.qb < 0 || q.qe > rl) return 1;
        std::vector<uint32_t> cig; std::string md; int score = INT32_MIN, nm = -1;
        if (!gen_cigar_one(x, opt, q.w, q.qe - q.qb, reads->codes + ro + q.qb, q.rb, q.re, &score, cig, &nm, md)) continue;
        o.score = score; o.n_cigar = (int32_t) cig.size(); o.nm = nm; o.n_md = (int32_t) md.size() + 1;
        all_ops.insert(all_ops.end(), cig.begin(), cig.end());
        all_md += md; all_md.push_back('\0');
    }
    *recs_out = (bm2_cigar_rec *) malloc(sizeof(bm2_cigar_rec) * (size_t) (n + 1)); memcpy(*recs_out, recs.data(), sizeof(bm2_cigar_rec) * (size_t) n);
    *cigar_out = (uint32_t *) malloc(4 * (all_ops.size() + 1)); memcpy(*cigar_out, all_ops.data(), 4 * all_ops.size());
    *md_out = (char *) malloc(all_md.size() + 1); memcpy(*md_out, all_md.data(), all_md.size());
    *n_ops_out = (int64_t) all_ops.size(); *n_md_out = (int64_t) all_md.size();
    return 0;
}


/* ================================================================================================
 * Mate-rescue local alignment: ksw_align2 (src/ksw.cpp:324-381) = ksw_u8 (:111-233) or
 * ksw_i16 (:235-316) forward, then the same kernel on the reversed prefixes to find the start.
 * The SSE2 kernels are restated lane by lane: a vector j of the striped layout holds the query
 * positions j + l * slen (l = lane), the first pass carries F inside a lane only, the lazy-F loop
 * (at most 16 sweeps, global early exit) completes H but not E.
 * ============================================================================================== */
namespace {
const int KX_BYTE = 0x10000, KX_STOP = 0x20000, KX_SUBO = 0x40000, KX_START = 0x80000;
struct KswR { int score, te, qe, score2, te2, tb, qb; };

KswR ksw_striped(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                 int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
    const int m = 5;
    const int p = size == 1 ? 16 : 8;                         /* values per vector (ksw_qinit, :64) */
    const int slen = (qlen + p - 1) / p, nlen = slen * p;
    int shift = 127, mdiff = 0;
    for (int a = 0; a < m * m; ++a) { if (mat[a] < shift) shift = mat[a]; if (mat[a] > mdiff) mdiff = mat[a]; }
    const int qmax = mdiff;
    shift = (256 - shift) & 0xff;                             /* uint8_t (:82) */
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int minsc = (xtra & KX_SUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & KX_STOP) ? xtra & 0xffff : 0x10000;
    const int cap = size == 1 ? 255 : 32767;
    auto sat0 = [](int v) { return v < 0 ? 0 : v; };
    std::vector<int> H0(nlen + 1, 0), H1(nlen + 1, 0), E(nlen + 1, 0), Hmax(nlen + 1, 0), fl(p, 0), hh(p, 0);
    std::vector<uint64_t> b;
    KswR r = { 0, -1, -1, -1, -1, -1, -1 };
    int gmax = 0, te = -1;
    for (int i = 0; i < tlen; ++i) {
        const int8_t *ma = mat + target[i] * m;
        int rowmax = 0;
        /* first pass: lane l walks its segment l*slen .. l*slen+slen-1 with its own F, starting at 0 */
        for (int l = 0; l < p; ++l) {
            int f = 0;
            for (int j = 0; j < slen; ++j) {
                const int k = l * slen + j;
                const int sc = k >= qlen ? 0 : ma[query[k]];
                int h = k == 0 ? 0 : H0[k - 1];
                if (size == 1) { h = h + sc + shift; if (h > 255) h = 255; h = sat0(h - shift); }        /* adds_epu8 + subs_epu8 */
                else { h = h + sc; if (h > cap) h = cap; if (h < -32768) h = -32768; }                  /* adds_epi16 */
                int e = E[k];
                if (e > h) h = e;
                if (f > h) h = f;
                if (h > rowmax) rowmax = h;
                H1[k] = h;
                e = sat0(e - e_del); { const int t = sat0(h - oe_del); if (t > e) e = t; }
                E[k] = e;
                f = sat0(f - e_ins); { const int t = sat0(h - oe_ins); if (t > f) f = t; }
            }
            fl[l] = f;
        }
        /* lazy F: shift F one lane up, sweep the vectors; stop as soon as no lane's F beats H - oe_ins (:172-186, :277-289) */
        {
            bool done = false;
            for (int kk = 0; kk < 16 && !done; ++kk) {
                for (int l = p - 1; l > 0; --l) fl[l] = fl[l - 1];
                fl[0] = 0;
                for (int j = 0; j < slen && !done; ++j) {
                    bool any = false;
                    for (int l = 0; l < p; ++l) {
                        const int k = l * slen + j;
                        int h = H1[k];
                        if (fl[l] > h) h = fl[l];
                        H1[k] = h;
                        h = sat0(h - oe_ins);
                        fl[l] = sat0(fl[l] - e_ins);
                        if (fl[l] > h) any = true;
                    }
                    if (!any) done = true;
                }
            }
        }
        const int imax = rowmax;
        if (imax >= minsc) {
            if (b.empty() || (int32_t) b.back() + 1 != i) b.push_back((uint64_t) imax << 32 | (uint32_t) i);
            else if ((int) (b.back() >> 32) < imax) b.back() = (uint64_t) imax << 32 | (uint32_t) i;
        }
        if (imax > gmax) {
            gmax = imax; te = i;
            for (int k = 0; k < nlen; ++k) Hmax[k] = H1[k];
            if (size == 1 ? (gmax + shift >= 255 || gmax >= endsc) : (gmax >= endsc)) break;
        }
        H0.swap(H1);
    }
    r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (size == 2 || r.score != 255) {
        int mx = -1;
        if (size == 2) r.qe = -1;
        /* memory order of the vectors: element i is lane i % p of vector i / p -> query position i / p + (i % p) * slen */
        for (int i = 0; i < nlen; ++i) {
            const int pos = i / p + i % p * slen, v = Hmax[pos];
            if (v > mx) { mx = v; r.qe = pos; }
            else if (v == mx && pos < r.qe) r.qe = pos;
        }
        if (!b.empty()) {
            const int d = (r.score + qmax - 1) / qmax, low = te - d, high = te + d;
            for (size_t k = 0; k < b.size(); ++k) {
                const int e = (int32_t) b[k];
                if ((e < low || e > high) && (int) (b[k] >> 32) > r.score2) { r.score2 = (int) (b[k] >> 32); r.te2 = e; }
            }
        }
    }
    return r;
}
}  // namespace

extern "C" void bm2o_ksw_align2(int32_t qlen, const uint8_t *query, int32_t tlen, const uint8_t *target, const int8_t *mat,
                                int32_t o_del, int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out)
{
    const int size = (xtra & KX_BYTE) ? 1 : 2;
    KswR r = ksw_striped(size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra);
    if (!((xtra & KX_START) == 0 || ((xtra & KX_SUBO) && r.score < (xtra & 0xffff)))) {
        /* the reversed prefixes (+1: qe / te point at the end itself); the target keeps its tail (:366-371) */
        std::vector<uint8_t> q(query, query + r.qe + 1), t(target, target + tlen);
        std::reverse(q.begin(), q.end()); std::reverse(t.begin(), t.begin() + r.te + 1);
        KswR rr = ksw_striped(size, r.qe + 1, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, KX_STOP | r.score);
        if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    }
    out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
}


/* ================================================================================================
 * Mate rescue around the local alignment: mem_pestat (src/bwamem_pair.cpp:88-148) and
 * mem_matesw (:150-283, MATE_SORT == 0).
 * ============================================================================================== */
static int o_infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist) {      /* mem_infer_dir (:57-65) */
    const int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
    const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
    *dist = p2 > b1 ? p2 - b1 : b1 - p2;
    return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

static int o_cal_sub(const bm2_mem_opt_t *opt, const bm2_alnreg_t *a, int n) {       /* cal_sub (:67-79) */
    int j;
    for (j = 1; j < n; ++j) {
        const int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
        const int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
        if (e_min > b_max) {
            const int min_l = a[j].qe - a[j].qb < a[0].qe - a[0].qb ? a[j].qe - a[j].qb : a[0].qe - a[0].qb;
            if (e_min - b_max >= min_l * opt->mask_level) break;
        }
    }
    return j < n ? a[j].score : opt->min_seed_len * opt->a;
}

/* lh[12] = low, high, failed of the four orientations; as[8] = avg, std */
extern "C" void bm2o_pestat(const bm2_mem_opt_t *opt, int64_t l_pac, int32_t n_reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                            int32_t *lh, double *as)
{
    std::vector<uint64_t> isize[4];
    for (int i = 0; i < n_reads >> 1; ++i) {
        const bm2_alnreg_t *r0 = regs + read_off[i << 1 | 0], *r1 = regs + read_off[i << 1 | 1];
        const int n0 = (int) (read_off[(i << 1 | 0) + 1] - read_off[i << 1 | 0]), n1 = (int) (read_off[(i << 1 | 1) + 1] - read_off[i << 1 | 1]);
        if (n0 == 0 || n1 == 0) continue;
        if (o_cal_sub(opt, r0, n0) > 0.8 * r0[0].score) continue;            /* MIN_RATIO */
        if (o_cal_sub(opt, r1, n1) > 0.8 * r1[0].score) continue;
        if (r0[0].rid != r1[0].rid) continue;
        int64_t is;
        const int dir = o_infer_dir(l_pac, r0[0].rb, r1[0].rb, &is);
        if (is && is <= opt->max_ins) isize[dir].push_back((uint64_t) is);
    }
    for (int d = 0; d < 4; ++d) {
        int32_t *o = lh + 3 * d; double *oa = as + 2 * d;
        o[0] = o[1] = o[2] = 0; oa[0] = oa[1] = 0;
        std::vector<uint64_t> &q = isize[d];
        if (q.size() < 10) { o[2] = 1; continue; }                               /* MIN_DIR_CNT */
        std::sort(q.begin(), q.end());                                           /* ks_introsort_64: plain keys, any sort */
        const size_t n = q.size();
        const int p25 = (int) q[(int) (.25 * n + .499)], p50 = (int) q[(int) (.50 * n + .499)], p75 = (int) q[(int) (.75 * n + .499)];
        (void) p50;
        int low = (int) (p25 - 2.0 * (p75 - p25) + .499);                        /* OUTLIER_BOUND */
        if (low < 1) low = 1;
        int high = (int) (p75 + 2.0 * (p75 - p25) + .499);
        double avg = 0; size_t x = 0;
        for (size_t i = 0; i < n; ++i) if (q[i] >= (uint64_t) low && q[i] <= (uint64_t) high) { avg += q[i]; ++x; }
        avg /= x;
        double sd = 0;
        for (size_t i = 0; i < n; ++i) if (q[i] >= (uint64_t) low && q[i] <= (uint64_t) high) sd += (q[i] - avg) * (q[i] - avg);
        sd = sqrt(sd / x);
        low = (int) (p25 - 3.0 * (p75 - p25) + .499);                            /* MAPPING_BOUND */
        high = (int) (p75 + 3.0 * (p75 - p25) + .499);
        if (low > avg - 4.0 * sd) low = (int) (avg - 4.0 * sd + .499);            /* MAX_STDDEV */
        if (high < avg + 4.0 * sd) high = (int) (avg + 4.0 * sd + .499);
        if (low < 1) low = 1;
        o[0] = low; o[1] = high; oa[0] = avg; oa[1] = sd;
    }
    size_t mx = 0;
    for (int d = 0; d < 4; ++d) mx = mx > isize[d].size() ? mx : isize[d].size();
    for (int d = 0; d < 4; ++d) if (lh[3 * d + 2] == 0 && isize[d].size() < mx * 0.05) lh[3 * d + 2] = 1;      /* MIN_DIR_RATIO */
}

/* mem_matesw: a = the anchor alignment of one read, ms = its mate's sequence, ma[0..*n_ma) the mate's regs (room for 4 more).
 * pes_lh[12] = low, high, failed.  Returns n (number of orientations aligned). */
extern "C" int bm2o_matesw(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const int32_t *pes_lh, const bm2_alnreg_t *a, int32_t l_ms,
                           const uint8_t *ms, bm2_alnreg_t *ma, int32_t *n_ma)
{
    const int64_t l_pac = x->l_pac;
    int skip[4], n = 0, nm = *n_ma;
    for (int r = 0; r < 4; ++r) skip[r] = pes_lh[3 * r + 2] ? 1 : 0;
    for (int i = 0; i < nm; ++i) {
        int64_t dist;
        const int r = o_infer_dir(l_pac, a->rb, ma[i].rb, &dist);
        if (dist >= pes_lh[3 * r] && dist <= pes_lh[3 * r + 1]) skip[r] = 1;
    }
    if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
    for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        const int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1);
        const int low = pes_lh[3 * r], high = pes_lh[3 * r + 1];
        std::vector<uint8_t> seq(ms, ms + l_ms);
        if (is_rev) for (int i = 0; i < l_ms; ++i) seq[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
        int64_t rb, re;
        if (!is_rev) {
            rb = is_larger ? a->rb + low : a->rb - high;
            re = (is_larger ? a->rb + high : a->rb - low) + l_ms;
        } else {
            rb = (is_larger ? a->rb + low : a->rb - high) - l_ms;
            re = is_larger ? a->rb + high : a->rb - low;
        }
        if (rb < 0) rb = 0;
        if (re > l_pac << 1) re = l_pac << 1;
        int rid = -1;
        if (rb < re) {                                                           /* bns_fetch_seq (src/bntseq.cpp:453-482) */
            const int64_t mid = (rb + re) >> 1;
            const int mrev = mid >= l_pac;
            rid = pos2rid(x, depos(x, mid));
            int64_t far_beg = x->ann_offset[rid], far_end = far_beg + x->ann_len[rid];
            if (mrev) { const int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
            rb = rb > far_beg ? rb : far_beg;
            re = re < far_end ? re : far_end;
        }
        if (a->rid == rid && re - rb >= opt->min_seed_len) {
            const int xtra = KX_SUBO | KX_START | (l_ms * opt->a < 250 ? KX_BYTE : 0) | (opt->min_seed_len * opt->a);
            int32_t al[7];
            bm2o_ksw_align2(l_ms, seq.data(), (int32_t) (re - rb), x->ref_string + rb, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, xtra, al);
            const int score = al[0], te = al[1], qe = al[2], score2 = al[3], tb = al[5], qb = al[6];
            if (score >= opt->min_seed_len && qb >= 0) {
                bm2_alnreg_t b; memset(&b, 0, sizeof(b));
                b.rid = a->rid;
                reg_set_is_alt(b, (a->n_comp_is_alt >> 30) & 3);
                b.qb = is_rev ? l_ms - (qe + 1) : qb;
                b.qe = is_rev ? l_ms - qb : qe + 1;
                b.rb = is_rev ? (l_pac << 1) - (rb + te + 1) : rb + tb;
                b.re = is_rev ? (l_pac << 1) - (rb + tb) : rb + te + 1;
                b.score = score; b.csub = score2; b.secondary = -1;
                b.seedcov = (int) ((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
                int i;
                for (i = 0; i < nm; ++i) if (ma[i].score < b.score) break;       /* keep ma sorted by score (:233-238) */
                for (int k = nm; k > i; --k) ma[k] = ma[k - 1];
                ma[i] = b; ++nm;
            }
            ++n;
        }
        if (n) nm = sort_dedup_patch(x, opt, nullptr, nm, ma);
    }
    *n_ma = nm;
    return n;
}


/* ================================================================================================
 * SAM stage, single-end (groundwork): mem_mark_primary_se, mem_approx_mapq_se, mem_reg2aln and the
 * record selection of mem_reg2sam.
 * ============================================================================================== */
static inline uint64_t o_hash_64(uint64_t key) {                 /* src/utils.h:117-128 */
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}
static inline int reg_is_alt(const bm2_alnreg_t &a) { return (a.n_comp_is_alt >> 30) & 3; }

static void mark_primary_core(const bm2_mem_opt_t *opt, int n, bm2_alnreg_t *a, std::vector<int> &z) {      /* :1392-1418 */
    int tmp = opt->a + opt->b;
    tmp = opt->o_del + opt->e_del > tmp ? opt->o_del + opt->e_del : tmp;
    tmp = opt->o_ins + opt->e_ins > tmp ? opt->o_ins + opt->e_ins : tmp;
    z.clear(); z.push_back(0);
    for (int i = 1; i < n; ++i) {
        size_t k;
        for (k = 0; k < z.size(); ++k) {
            const int j = z[k];
            const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * opt->mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (reg_is_alt(a[j]) || !reg_is_alt(a[i]))) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == z.size()) z.push_back(i);
        else a[i].secondary = z[k];
    }
}

static int mark_primary_se(const bm2_mem_opt_t *opt, int n, bm2_alnreg_t *a, int64_t id) {                  /* :1420-1468 */
    if (n == 0) return 0;
    int n_pri = 0;
    std::vector<int> z;
    for (int i = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = o_hash_64((uint64_t) (id + i));
        if (!reg_is_alt(a[i])) ++n_pri;
    }
    ks_introsort(a, n, [](const bm2_alnreg_t &p, const bm2_alnreg_t &q) {                                    /* alnreg_hlt */
        return p.score > q.score || (p.score == q.score && (reg_is_alt(p) < reg_is_alt(q) || (reg_is_alt(p) == reg_is_alt(q) && p.hash < q.hash)));
    });
    mark_primary_core(opt, n, a, z);
    for (int i = 0; i < n; ++i) {
        bm2_alnreg_t *p = &a[i];
        p->secondary_all = i;
        if (!reg_is_alt(*p) && p->secondary >= 0 && reg_is_alt(a[p->secondary])) p->alt_sc = a[p->secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        z.assign((size_t) n, 0);
        if (n_pri > 0) ks_introsort(a, n, [](const bm2_alnreg_t &p, const bm2_alnreg_t &q) {                 /* alnreg_hlt2 */
            return reg_is_alt(p) < reg_is_alt(q) || (reg_is_alt(p) == reg_is_alt(q) && (p.score > q.score || (p.score == q.score && p.hash < q.hash)));
        });
        for (int i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (int i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (reg_is_alt(a[i])) a[i].secondary = INT_MAX;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {
            for (int i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            mark_primary_core(opt, n_pri, a, z);
        }
    } else {
        for (int i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    return n_pri;
}

static int approx_mapq_se(const bm2_mem_opt_t *opt, const bm2_alnreg_t *a) {                                 /* :1470-1494 */
    int mapq, l, sub = a->sub ? a->sub : opt->min_seed_len * opt->a;
    double identity;
    sub = a->csub > sub ? a->csub : sub;
    if (sub >= a->score) return 0;
    l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int) (a->re - a->rb);
    identity = 1. - (double) (l * opt->a - a->score) / (opt->a + opt->b) / l;
    if (a->score == 0) mapq = 0;
    else if (opt->mapQ_coef_len > 0) {
        double tmp = l < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(l);
        tmp *= identity * identity;
        mapq = (int) (6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
    } else {
        mapq = (int) (30.0 * (1. - (double) sub / a->score) * log(a->seedcov) + .499);                      /* MEM_MAPQ_COEF */
        mapq = identity < 0.95 ? (int) (mapq * identity * identity + .499) : mapq;
    }
    if (a->sub_n > 0) mapq -= (int) (4.343 * log(a->sub_n + 1) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int) (mapq * (1. - a->frac_rep) + .499);
    return mapq;
}

static inline int o_infer_bw(int l1, int l2, int score, int a, int q, int r) {                               /* :1811-1818 */
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    int w = (int) ((double) ((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    const int d = l1 > l2 ? l1 - l2 : l2 - l1;
    if (w < d) w = d;
    return w;
}

struct OAln { int flag, rid, mapq, nm, score, sub, is_rev, is_alt, alt_sc; int64_t pos; std::vector<uint32_t> cigar; std::string md, XA; };

static OAln reg2aln(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int l_query, const uint8_t *query, const bm2_alnreg_t *ar) {   /* :1732-1805 */
    OAln a; a.flag = 0; a.rid = -1; a.mapq = 0; a.nm = 0; a.score = 0; a.sub = 0; a.is_rev = 0; a.is_alt = 0; a.alt_sc = 0; a.pos = -1;
    if (ar == 0 || ar->rb < 0 || ar->re < 0) { a.flag |= 0x4; return a; }
    const int qb = ar->qb, qe = ar->qe;
    const int64_t rb = ar->rb, re = ar->re;
    a.mapq = ar->secondary < 0 ? approx_mapq_se(opt, ar) : 0;
    if (ar->secondary >= 0) a.flag |= 0x100;
    int tmp = o_infer_bw(qe - qb, (int) (re - rb), ar->truesc, opt->a, opt->o_del, opt->e_del);
    int w2 = o_infer_bw(qe - qb, (int) (re - rb), ar->truesc, opt->a, opt->o_ins, opt->e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > opt->w) w2 = w2 < ar->w ? w2 : ar->w;
    int i = 0, score = 0, NM = -1, last_sc = -(1 << 30);
    do {
        w2 = w2 < opt->w << 2 ? w2 : opt->w << 2;
        gen_cigar_one(x, opt, w2, qe - qb, query + qb, rb, re, &score, a.cigar, &NM, a.md);
        if (score == last_sc || w2 == opt->w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar->truesc - opt->a);
    a.nm = NM;
    const int is_rev = (rb < x->l_pac ? rb : re - 1) >= x->l_pac;
    int64_t pos = depos(x, rb < x->l_pac ? rb : re - 1);
    a.is_rev = is_rev;
    if (!a.cigar.empty()) {                                      /* squeeze out leading or trailing deletions */
        if ((a.cigar[0] & 0xf) == 2) { pos += a.cigar[0] >> 4; a.cigar.erase(a.cigar.begin()); }
        else if ((a.cigar.back() & 0xf) == 2) a.cigar.pop_back();
    }
    if (qb != 0 || qe != l_query) {
        const int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
        if (clip5) a.cigar.insert(a.cigar.begin(), (uint32_t) clip5 << 4 | 3);
        if (clip3) a.cigar.push_back((uint32_t) clip3 << 4 | 3);
    }
    a.rid = pos2rid(x, pos);
    a.pos = pos - x->ann_offset[a.rid];
    a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
    a.is_alt = reg_is_alt(*ar); a.alt_sc = ar->alt_sc;
    return a;
}

/* mem_reorder_primary5 (src/bwamem.cpp:1496-1518), option -5: the primary with the smallest query start becomes record 0 */
static void reorder_primary5(int T, int n, bm2_alnreg_t *a) {
    int n_pri = 0, left_st = INT_MAX, left_k = -1;
    for (int k = 0; k < n; ++k) if (a[k].secondary < 0 && !reg_is_alt(a[k]) && a[k].score >= T) ++n_pri;
    if (n_pri <= 1) return;
    for (int k = 0; k < n; ++k) {
        const bm2_alnreg_t *p = &a[k];
        if (p->secondary >= 0 || reg_is_alt(*p) || p->score < T) continue;
        if (p->qb < left_st) { left_st = p->qb; left_k = k; }
    }
    if (left_k == 0) return;
    std::swap(a[0], a[left_k]);
    for (int k = 1; k < n; ++k) {
        bm2_alnreg_t *p = &a[k];
        if (p->secondary == 0) p->secondary = left_k; else if (p->secondary == left_k) p->secondary = 0;
        if (p->secondary_all == 0) p->secondary_all = left_k; else if (p->secondary_all == left_k) p->secondary_all = 0;
    }
}

extern "C" int bm2o_sam_se(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, bm2_alnreg_t *regs, const int64_t *read_off,
                           int64_t id_base, bm2o_aln **alns_out, int64_t *n_alns, uint32_t **cigar_out, int64_t *n_ops_out, char **md_out, int64_t *n_md_out)
{
    std::vector<bm2o_aln> out; std::vector<uint32_t> ops; std::string mds;
    auto push = [&](int read, const OAln &q) {
        bm2o_aln o; o.read = read; o.flag = q.flag; o.rid = q.rid; o.mapq = q.mapq; o.nm = q.nm; o.score = q.score; o.sub = q.sub; o.is_rev = q.is_rev;
        o.is_alt = q.is_alt; o.alt_sc = q.alt_sc; o.n_cigar = (int32_t) q.cigar.size(); o.n_md = (int32_t) q.md.size() + 1; o.pos = q.pos;
        o.cigar_off = (int64_t) ops.size(); o.md_off = (int64_t) mds.size();
        ops.insert(ops.end(), q.cigar.begin(), q.cigar.end()); mds += q.md; mds.push_back('\0');
        out.push_back(o);
    };
    for (int r = 0; r < reads->n_reads; ++r) {
        bm2_alnreg_t *a = regs + read_off[r];
        const int n = (int) (read_off[r + 1] - read_off[r]);
        const uint8_t *query = reads->codes + reads->offsets[r];
        const int l_query = (int) (reads->offsets[r + 1] - reads->offsets[r]);
        mark_primary_se(opt, n, a, id_base + r);
        if (opt->flag & 0x800) reorder_primary5(opt->T, n, a);                                                /* MEM_F_PRIMARY5 (:1329) */
        /* mem_reg2sam (:1521-1577), extra_flag = 0, no mate */
        std::vector<OAln> aa;
        int l = 0;
        for (int k = 0; k < n; ++k) {
            const bm2_alnreg_t *p = &a[k];
            if (p->score < opt->T) continue;
            if (p->secondary >= 0 && (reg_is_alt(*p) || !(opt->flag & 0x8))) continue;                       /* MEM_F_ALL */
            if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < a[p->secondary].score * opt->drop_ratio) continue;
            OAln q = reg2aln(x, opt, l_query, query, p);
            if (p->secondary >= 0) q.sub = -1;
            if (l && p->secondary < 0) q.flag |= (opt->flag & 0x10) ? 0x10000 : 0x800;                       /* MEM_F_NO_MULTI */
            if (!(opt->flag & 0x1000) && l && !reg_is_alt(*p) && q.mapq > aa[0].mapq) q.mapq = aa[0].mapq;    /* MEM_F_KEEP_SUPP_MAPQ */
            aa.push_back(q);
            ++l;
        }
        if (aa.empty()) push(r, reg2aln(x, opt, l_query, query, 0));
        else for (const OAln &q : aa) push(r, q);
    }
    const size_t n = out.size();
    *alns_out = (bm2o_aln *) malloc(sizeof(bm2o_aln) * (n + 1)); memcpy(*alns_out, out.data(), sizeof(bm2o_aln) * n);
    *cigar_out = (uint32_t *) malloc(4 * (ops.size() + 1)); memcpy(*cigar_out, ops.data(), 4 * ops.size());
    *md_out = (char *) malloc(mds.size() + 1); memcpy(*md_out, mds.data(), mds.size());
    *n_alns = (int64_t) n; *n_ops_out = (int64_t) ops.size(); *n_md_out = (int64_t) mds.size();
    return 0;
}


/* ================================================================================================
 * SAM stage, paired-end: mem_pair (src/bwamem_pair.cpp:285-346), mem_sam_pe (:353-552), mem_reg2sam with a mate, and the
 * columns of mem_aln2sam (src/bwamem.cpp:1592-1730).
 * ============================================================================================== */
namespace {
struct P64 { uint64_t x, y; };
inline bool p64_lt(const P64 &a, const P64 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }
inline int raw_mapq(int diff, int a) { return (int) (6.02 * diff / a + .499); }
inline int get_rlen(const std::vector<uint32_t> &c) { int l = 0; for (uint32_t v : c) { const int op = v & 0xf; if (op == 0 || op == 2) l += v >> 4; } return l; }

int o_mem_pair(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const int32_t *lh, const double *as, const std::vector<bm2_alnreg_t> a[2], int id,
               int *sub, int *n_sub, int z[2], const int n_pri[2])
{
    std::vector<P64> v, u;
    const int64_t l_pac = x->l_pac;
    for (int r = 0; r < 2; ++r)
        for (int i = 0; i < n_pri[r]; ++i) {
            const bm2_alnreg_t *e = &a[r][i];
            P64 key;
            key.x = (uint64_t) (e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb);
            key.x = (uint64_t) e->rid << 32 | (key.x - (uint64_t) x->ann_offset[e->rid]);
            key.y = (uint64_t) e->score << 32 | (uint64_t) (i << 2 | (e->rb >= l_pac) << 1 | r);
            v.push_back(key);
        }
    std::sort(v.begin(), v.end(), p64_lt);
    int y[4] = { -1, -1, -1, -1 };
    for (int i = 0; i < (int) v.size(); ++i) {
        for (int r = 0; r < 2; ++r) {
            const int dir = r << 1 | (int) (v[i].y >> 1 & 1);
            if (lh[3 * dir + 2]) continue;
            const int which = r << 1 | (int) ((v[i].y & 1) ^ 1);
            if (y[which] < 0) continue;
            for (int k = y[which]; k >= 0; --k) {
                if ((int) (v[k].y & 3) != which) continue;
                const int64_t dist = (int64_t) v[i].x - (int64_t) v[k].x;
                if (dist > lh[3 * dir + 1]) break;
                if (dist < lh[3 * dir]) continue;
                const double ns = (dist - as[2 * dir]) / as[2 * dir + 1];
                int q = (int) ((v[i].y >> 32) + (v[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a + .499);
                if (q < 0) q = 0;
                P64 p;
                p.y = (uint64_t) k << 32 | (uint64_t) i;
                p.x = (uint64_t) q << 32 | (o_hash_64(p.y ^ (uint64_t) (int64_t) (id << 8)) & 0xffffffffU);
                u.push_back(p);
            }
        }
        y[v[i].y & 3] = i;
    }
    int ret;
    if (!u.empty()) {
        int tmp = opt->a + opt->b;
        tmp = tmp > opt->o_del + opt->e_del ? tmp : opt->o_del + opt->e_del;
        tmp = tmp > opt->o_ins + opt->e_ins ? tmp : opt->o_ins + opt->e_ins;
        std::sort(u.begin(), u.end(), p64_lt);
        const int i = (int) (u.back().y >> 32), k = (int) (u.back().y << 32 >> 32);
        z[v[i].y & 1] = (int) (v[i].y << 32 >> 34);
        z[v[k].y & 1] = (int) (v[k].y << 32 >> 34);
        ret = (int) (u.back().x >> 32);
        *sub = u.size() > 1 ? (int) (u[u.size() - 2].x >> 32) : 0;
        *n_sub = 0;
        for (long j = (long) u.size() - 2; j >= 0; --j) if (*sub - (int) (u[j].x >> 32) <= tmp) ++*n_sub;
    } else { ret = 0; *sub = 0; *n_sub = 0; }
    return ret;
}

void gen_alt(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const char *const *names, int l_query, const uint8_t *query, const std::vector<bm2_alnreg_t> &a,
             std::vector<std::string> &XA);

/* the records mem_reg2sam keeps for one read (:1534-1560) */
void reg2sam_list(const bm2_index_desc *x, const bm2_mem_opt_t *opt, int l_query, const uint8_t *query, const std::vector<bm2_alnreg_t> &a, int extra_flag,
                  std::vector<OAln> &aa, const char *const *names = 0)
{
    aa.clear();
    std::vector<std::string> XA;
    if (names && !(opt->flag & 0x8)) gen_alt(x, opt, names, l_query, query, a, XA);
    int l = 0;
    for (size_t k = 0; k < a.size(); ++k) {
        const bm2_alnreg_t *p = &a[k];
        if (p->score < opt->T) continue;
        if (p->secondary >= 0 && (reg_is_alt(*p) || !(opt->flag & 0x8))) continue;
        if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < a[p->secondary].score * opt->drop_ratio) continue;
        OAln q = reg2aln(x, opt, l_query, query, p);
        if (!XA.empty()) q.XA = XA[k];
        q.flag |= extra_flag;
        if (p->secondary >= 0) q.sub = -1;
        if (l && p->secondary < 0) q.flag |= (opt->flag & 0x10) ? 0x10000 : 0x800;
        if (!(opt->flag & 0x1000) && l && !reg_is_alt(*p) && q.mapq > aa[0].mapq) q.mapq = aa[0].mapq;
        aa.push_back(q);
        ++l;
    }
    if (aa.empty()) { OAln t = reg2aln(x, opt, l_query, query, 0); t.flag |= extra_flag; aa.push_back(t); }
}

struct SamOut { std::vector<bm2o_samrec> recs; std::vector<uint32_t> ops; std::string mds; };
/* optional: the SAM text of every line after QNAME (names of the contigs, the read's codes and qualities) */
struct TextCtx { const char *const *names; const uint8_t *seq; const char *qual; int l_seq; std::string *text; };

/* mem_gen_alt (src/bwamem_extra.cpp:130-183): the XA string of every primary; call after mem_mark_primary_se */
void gen_alt(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const char *const *names, int l_query, const uint8_t *query, const std::vector<bm2_alnreg_t> &a,
             std::vector<std::string> &XA)
{
    const int n = (int) a.size();
    XA.assign((size_t) n, std::string());
    std::vector<int> cnt((size_t) n, 0); std::vector<char> has_alt((size_t) n, 0);
    const double xa_drop = opt->XA_drop_ratio;             /* get_pri_idx takes the float option as a double (src/bwamem_extra.cpp:122) */
    auto pri_idx = [&](int i) { const int k = a[i].secondary_all; return (k >= 0 && a[i].score >= a[k].score * xa_drop) ? k : -1; };
    int tot = 0;
    for (int i = 0; i < n; ++i) { const int r = pri_idx(i); if (r >= 0) { ++cnt[r]; ++tot; if (reg_is_alt(a[i])) has_alt[r] = 1; } }
    if (tot == 0) return;
    for (int i = 0; i < n; ++i) {
        const int r = pri_idx(i);
        if (r < 0) continue;
        if (cnt[r] > opt->max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt->max_XA_hits)) continue;
        const OAln t = reg2aln(x, opt, l_query, query, &a[i]);
        std::string e = names[t.rid]; e += ','; e += "+-"[t.is_rev]; e += std::to_string(t.pos + 1); e += ',';
        for (uint32_t v : t.cigar) { e += std::to_string(v >> 4); e += "MIDSHN"[v & 0xf]; }
        e += ','; e += std::to_string(t.nm); e += ';';
        XA[r] += e;
    }
}

void cigar_text(const bm2_mem_opt_t *opt, const OAln &p, int which, std::string &t) {               /* add_cigar (src/bwamem.cpp:1579-1590) */
    if (p.cigar.empty()) { t += '*'; return; }
    for (uint32_t v : p.cigar) {
        int c = v & 0xf;
        if (!(opt->flag & 0x200) && !p.is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
        t += std::to_string(v >> 4); t += "MIDSH"[c];
    }
}

/* the columns of mem_aln2sam for list[which] with mate m_ (may be null) */
void aln2sam(const bm2_mem_opt_t *opt, int read, const std::vector<OAln> &list, int which, const OAln *m_, SamOut &o, const TextCtx *tc = 0)
{
    OAln p = list[which], mt; const OAln *m = 0;
    if (m_) { mt = *m_; m = &mt; }
    p.flag |= m ? 0x1 : 0;
    p.flag |= p.rid < 0 ? 0x4 : 0;
    p.flag |= m && m->rid < 0 ? 0x8 : 0;
    if (p.rid < 0 && m && m->rid >= 0) { p.rid = m->rid; p.pos = m->pos; p.is_rev = m->is_rev; p.cigar.clear(); }
    if (m && m->rid < 0 && p.rid >= 0) { mt.rid = p.rid; mt.pos = p.pos; mt.is_rev = p.is_rev; mt.cigar.clear(); }
    p.flag |= p.is_rev ? 0x10 : 0;
    p.flag |= m && m->is_rev ? 0x20 : 0;
    bm2o_samrec r; memset(&r, 0, sizeof(r));
    r.read = read; r.flag = (p.flag & 0xffff) | (p.flag & 0x10000 ? 0x100 : 0);
    r.rid = p.rid; r.pos = p.rid >= 0 ? p.pos + 1 : 0; r.mapq = p.rid >= 0 ? p.mapq : 0;
    r.cigar_off = (int64_t) o.ops.size(); r.md_off = (int64_t) o.mds.size();
    if (p.rid >= 0)
        for (uint32_t v : p.cigar) {
            int c = v & 0xf;
            if (!(opt->flag & 0x200) && !p.is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
            o.ops.push_back((v >> 4) << 4 | (uint32_t) c);
        }
    r.n_cigar = (int32_t) (o.ops.size() - (size_t) r.cigar_off);
    r.rnext = -1; r.pnext = 0; r.tlen = 0;
    if (m && m->rid >= 0) {
        r.rnext = m->rid; r.pnext = m->pos + 1;
        if (p.rid == m->rid) {
            const int64_t p0 = p.pos + (p.is_rev ? get_rlen(p.cigar) - 1 : 0), p1 = m->pos + (m->is_rev ? get_rlen(m->cigar) - 1 : 0);
            if (m->cigar.empty() || p.cigar.empty()) r.tlen = 0;
            else r.tlen = -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0));
        }
    }
    if (!p.cigar.empty()) { r.nm = p.nm; o.mds += p.md; }          /* NM / MD only with a cigar (p.n_cigar) */
    o.mds.push_back('\0');
    r.n_md = (int32_t) (o.mds.size() - (size_t) r.md_off);
    r.score = p.score; r.sub = p.sub;
    if (p.cigar.empty()) r.n_cigar = 0;
    r.tlen_valid = 1;
    o.recs.push_back(r);
    if (!tc) return;
    /* the line after QNAME (src/bwamem.cpp:1614-1729) */
    std::string &t = *tc->text;
    t += std::to_string(r.flag); t += '\t';
    if (p.rid >= 0) {
        t += tc->names[p.rid]; t += '\t'; t += std::to_string(p.pos + 1); t += '\t'; t += std::to_string(p.mapq); t += '\t';
        cigar_text(opt, p, which, t);
    } else t += "*\t0\t0\t*";
    t += '\t';
    if (m && m->rid >= 0) {
        if (p.rid == m->rid) t += '='; else t += tc->names[m->rid];
        t += '\t'; t += std::to_string(m->pos + 1); t += '\t';
        if (p.rid == m->rid) t += std::to_string(r.tlen); else t += '0';
    } else t += "*\t0\t0";
    t += '\t';
    if (r.flag & 0x100) t += "*\t*";
    else {
        int qb = 0, qe = tc->l_seq;
        const bool trim = !p.cigar.empty() && which && !(opt->flag & 0x200) && !p.is_alt;
        auto is_clip = [](uint32_t v) { return (v & 0xf) == 3 || (v & 0xf) == 4; };
        if (!p.is_rev) {
            if (trim) { if (is_clip(p.cigar[0])) qb += p.cigar[0] >> 4; if (is_clip(p.cigar.back())) qe -= p.cigar.back() >> 4; }
            for (int i = qb; i < qe; ++i) t += "ACGTN"[tc->seq[i]];
            t += '\t';
            if (tc->qual) t.append(tc->qual + qb, (size_t) (qe - qb)); else t += '*';
        } else {
            if (trim) { if (is_clip(p.cigar[0])) qe -= p.cigar[0] >> 4; if (is_clip(p.cigar.back())) qb += p.cigar.back() >> 4; }
            for (int i = qe - 1; i >= qb; --i) t += "TGCAN"[tc->seq[i]];
            t += '\t';
            if (tc->qual) { for (int i = qe - 1; i >= qb; --i) t += tc->qual[i]; } else t += '*';
        }
    }
    if (!p.cigar.empty()) { t += "\tNM:i:"; t += std::to_string(p.nm); t += "\tMD:Z:"; t += p.md; }
    if (m && !m->cigar.empty()) { t += "\tMC:Z:"; cigar_text(opt, *m, which, t); }
    if (p.score >= 0) { t += "\tAS:i:"; t += std::to_string(p.score); }
    if (p.sub >= 0) { t += "\tXS:i:"; t += std::to_string(p.sub); }
    if (!(p.flag & 0x100)) {
        size_t i;
        for (i = 0; i < list.size(); ++i) if ((int) i != which && !(list[i].flag & 0x100)) break;
        if (i < list.size()) {
            t += "\tSA:Z:";
            for (i = 0; i < list.size(); ++i) {
                const OAln &q = list[i];
                if ((int) i == which || (q.flag & 0x100)) continue;
                t += tc->names[q.rid]; t += ','; t += std::to_string(q.pos + 1); t += ','; t += "+-"[q.is_rev]; t += ',';
                for (uint32_t v : q.cigar) { t += std::to_string(v >> 4); t += "MIDSH"[v & 0xf]; }
                t += ','; t += std::to_string(q.mapq); t += ','; t += std::to_string(q.nm); t += ';';
            }
        }
        if (p.alt_sc > 0) { char buf[64]; snprintf(buf, sizeof(buf), "\tpa:f:%.3f", (double) p.score / p.alt_sc); t += buf; }
    }
    if (!p.XA.empty()) { t += "\tXA:Z:"; t += p.XA; }
    t += '\n';
}
}  // namespace

static int sam_pe_core(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                       const int32_t *lh, const double *as, int64_t id_base, const char *const *names, const char *quals, std::string *text, SamOut &so)
{
    for (int pr = 0; pr < reads->n_reads >> 1; ++pr) {
        const int id = (int) (id_base + pr);
        std::vector<bm2_alnreg_t> a[2];
        const uint8_t *seq[2]; int l_seq[2];
        for (int i = 0; i < 2; ++i) {
            const int r = 2 * pr + i;
            a[i].assign(regs + read_off[r], regs + read_off[r + 1]);
            seq[i] = reads->codes + reads->offsets[r]; l_seq[i] = (int) (reads->offsets[r + 1] - reads->offsets[r]);
        }
        int extra_flag = 1, n_pri[2], z[2] = { 0, 0 }, o = 0, subo = 0, n_sub = 0;
        if (!(opt->flag & 0x20)) {                            /* !MEM_F_NO_RESCUE: mate SW for the best alignments (:378-412) */
            std::vector<bm2_alnreg_t> b[2];
            for (int i = 0; i < 2; ++i)
                for (size_t j = 0; j < a[i].size(); ++j)
                    if (a[i][j].score >= a[i][0].score - opt->pen_unpaired) b[i].push_back(a[i][j]);
            for (int i = 0; i < 2; ++i)
                for (size_t j = 0; j < b[i].size() && (int) j < opt->max_matesw; ++j) {
                    std::vector<bm2_alnreg_t> &ma = a[!i];
                    int32_t nm = (int32_t) ma.size();
                    ma.resize((size_t) nm + 4);
                    bm2o_matesw(x, opt, lh, &b[i][j], l_seq[!i], seq[!i], ma.data(), &nm);
                    ma.resize((size_t) nm);
                }
        }
        n_pri[0] = mark_primary_se(opt, (int) a[0].size(), a[0].data(), (int64_t) id << 1 | 0);
        n_pri[1] = mark_primary_se(opt, (int) a[1].size(), a[1].data(), (int64_t) id << 1 | 1);
        if (opt->flag & 0x800) {                                                                              /* MEM_F_PRIMARY5 (src/bwamem_pair.cpp:420-423) */
            reorder_primary5(opt->T, (int) a[0].size(), a[0].data());
            reorder_primary5(opt->T, (int) a[1].size(), a[1].data());
        }
        bool paired = false;
        std::vector<OAln> aa[2];
        OAln h[2];
        if (!(opt->flag & 0x4) && n_pri[0] && n_pri[1] && (o = o_mem_pair(x, opt, lh, as, a, id, &subo, &n_sub, z, n_pri)) > 0) {   /* !MEM_F_NOPAIRING */
            int is_multi[2];
            for (int i = 0; i < 2; ++i) {
                int j;
                for (j = 1; j < n_pri[i]; ++j) if (a[i][j].secondary < 0 && a[i][j].score >= opt->T) break;
                is_multi[i] = j < n_pri[i] ? 1 : 0;
            }
            if (!(is_multi[0] || is_multi[1])) {
                paired = true;
                int q_pe, q_se[2];
                const int score_un = a[0][0].score + a[1][0].score - opt->pen_unpaired;
                subo = subo > score_un ? subo : score_un;
                q_pe = raw_mapq(o - subo, opt->a);
                if (n_sub > 0) q_pe -= (int) (4.343 * log(n_sub + 1) + .499);
                if (q_pe < 0) q_pe = 0;
                if (q_pe > 60) q_pe = 60;
                q_pe = (int) (q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
                if (o > score_un) {
                    bm2_alnreg_t *c[2] = { &a[0][z[0]], &a[1][z[1]] };
                    for (int i = 0; i < 2; ++i) {
                        if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
                        q_se[i] = approx_mapq_se(opt, c[i]);
                    }
                    q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
                    q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
                    extra_flag |= 2;
                    q_se[0] = q_se[0] < raw_mapq(c[0]->score - c[0]->csub, opt->a) ? q_se[0] : raw_mapq(c[0]->score - c[0]->csub, opt->a);
                    q_se[1] = q_se[1] < raw_mapq(c[1]->score - c[1]->csub, opt->a) ? q_se[1] : raw_mapq(c[1]->score - c[1]->csub, opt->a);
                } else {
                    z[0] = z[1] = 0;
                    q_se[0] = approx_mapq_se(opt, &a[0][0]);
                    q_se[1] = approx_mapq_se(opt, &a[1][0]);
                }
                for (int i = 0; i < 2; ++i) {
                    const int k = a[i][z[i]].secondary_all;
                    if (k >= 0 && k < n_pri[i]) {
                        for (size_t j = 0; j < a[i].size(); ++j)
                            if (a[i][j].secondary_all == k || (int) j == k) a[i][j].secondary_all = z[i];
                        a[i][z[i]].secondary_all = -1;
                    }
                }
                std::vector<std::string> XA[2];
                if (names && !(opt->flag & 0x8)) for (int i = 0; i < 2; ++i) gen_alt(x, opt, names, l_seq[i], seq[i], a[i], XA[i]);
                for (int i = 0; i < 2; ++i) {
                    h[i] = reg2aln(x, opt, l_seq[i], seq[i], &a[i][z[i]]);
                    h[i].mapq = q_se[i];
                    h[i].flag |= 0x40 << i | extra_flag;
                    if (!XA[i].empty()) h[i].XA = XA[i][z[i]];
                    aa[i].push_back(h[i]);
                    if (n_pri[i] < (int) a[i].size()) {
                        const bm2_alnreg_t *p = &a[i][n_pri[i]];
                        if (p->score < opt->T || p->secondary >= 0 || !reg_is_alt(*p)) continue;
                        OAln g = reg2aln(x, opt, l_seq[i], seq[i], p);
                        g.flag |= 0x800 | 0x40 << i | extra_flag;
                        if (!XA[i].empty()) g.XA = XA[i][n_pri[i]];
                        aa[i].push_back(g);
                    }
                }
                for (int i = 0; i < 2; ++i) {
                    TextCtx tc = { names, seq[i], quals ? quals + reads->offsets[2 * pr + i] : 0, l_seq[i], text };
                    for (size_t k = 0; k < aa[i].size(); ++k) aln2sam(opt, 2 * pr + i, aa[i], (int) k, &h[!i], so, text ? &tc : 0);
                }
            }
        }
        if (!paired) {                                        /* no_pairing (:523-551) */
            for (int i = 0; i < 2; ++i) {
                int which = -1;
                if (!a[i].empty()) {
                    if (a[i][0].score >= opt->T) which = 0;
                    else if (n_pri[i] < (int) a[i].size() && a[i][n_pri[i]].score >= opt->T) which = n_pri[i];
                }
                h[i] = reg2aln(x, opt, l_seq[i], seq[i], which >= 0 ? &a[i][which] : 0);
            }
            if (!(opt->flag & 0x4) && h[0].rid == h[1].rid && h[0].rid >= 0) {
                int64_t dist;
                const int d = o_infer_dir(x->l_pac, a[0][0].rb, a[1][0].rb, &dist);
                if (!lh[3 * d + 2] && dist >= lh[3 * d] && dist <= lh[3 * d + 1]) extra_flag |= 2;
            }
            for (int i = 0; i < 2; ++i) {
                reg2sam_list(x, opt, l_seq[i], seq[i], a[i], (i == 0 ? 0x41 : 0x81) | extra_flag, aa[i], names);
                TextCtx tc = { names, seq[i], quals ? quals + reads->offsets[2 * pr + i] : 0, l_seq[i], text };
                for (size_t k = 0; k < aa[i].size(); ++k) aln2sam(opt, 2 * pr + i, aa[i], (int) k, &h[!i], so, text ? &tc : 0);
            }
        }
    }
    return 0;
}

extern "C" int bm2o_sam_pe(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                           const int32_t *lh, const double *as, int64_t id_base, bm2o_samrec **recs_out, int64_t *n_recs, uint32_t **cigar_out,
                           int64_t *n_ops_out, char **md_out, int64_t *n_md_out)
{
    SamOut so;
    const int rc = sam_pe_core(x, opt, reads, regs, read_off, lh, as, id_base, 0, 0, 0, so);
    if (rc) return rc;
    const size_t n = so.recs.size();
    *recs_out = (bm2o_samrec *) malloc(sizeof(bm2o_samrec) * (n + 1)); memcpy(*recs_out, so.recs.data(), sizeof(bm2o_samrec) * n);
    *cigar_out = (uint32_t *) malloc(4 * (so.ops.size() + 1)); memcpy(*cigar_out, so.ops.data(), 4 * so.ops.size());
    *md_out = (char *) malloc(so.mds.size() + 1); memcpy(*md_out, so.mds.data(), so.mds.size());
    *n_recs = (int64_t) n; *n_ops_out = (int64_t) so.ops.size(); *n_md_out = (int64_t) so.mds.size();
    return 0;
}


/* The SAM text of mem_sam_pe for every pair: each line from the FLAG column on (QNAME and the header are the caller's), tags NM MD MC AS XS SA pa XA
 * as mem_aln2sam writes them.  names: contig names; quals: qualities laid out like reads->codes (may be null: '*').  *text is malloc'd. */
extern "C" int bm2o_sam_pe_text(const bm2_index_desc *x, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const char *quals, const char *const *names,
                                const bm2_alnreg_t *regs, const int64_t *read_off, const int32_t *lh, const double *as, int64_t id_base, char **text, int64_t *len)
{
    SamOut so; std::string t;
    const int rc = sam_pe_core(x, opt, reads, regs, read_off, lh, as, id_base, names, quals, &t, so);
    if (rc) return rc;
    *text = (char *) malloc(t.size() + 1); memcpy(*text, t.data(), t.size()); (*text)[t.size()] = 0;
    *len = (int64_t) t.size();
    return 0;
}


/* Host probe for bench.py's cpu_baseline leg (no reference counterpart): aggregate rate of a fixed integer loop on `nthreads`
 * pthreads for `seconds`, in loop iterations per second.  rate(n) / rate(1) is the number of cores the box really gives `n`
 * threads (cgroup quotas, shared hosts and SMT siblings do not show in sched_getaffinity). */
#include <pthread.h>
#include <time.h>
struct ProbeArg { double seconds; volatile uint64_t iters; };
static double probe_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void *probe_thread(void *p) {
    ProbeArg *a = (ProbeArg *) p;
    const double t_end = probe_now() + a->seconds;
    uint64_t x = 88172645463325252ULL, n = 0;
    do {
        for (int i = 0; i < 100000; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; }
        n += 100000;
    } while (probe_now() < t_end);
    a->iters = n + (x == 0);
    return 0;
}
extern "C" double bm2o_cpu_probe(int32_t nthreads, double seconds) {
    if (nthreads < 1) nthreads = 1;
    std::vector<ProbeArg> args((size_t) nthreads);
    std::vector<pthread_t> th((size_t) nthreads);
    const double t0 = probe_now();
    for (int i = 0; i < nthreads; ++i) { args[i].seconds = seconds; args[i].iters = 0; pthread_create(&th[i], 0, probe_thread, &args[i]); }
    uint64_t tot = 0;
    for (int i = 0; i < nthreads; ++i) { pthread_join(th[i], 0); tot += args[i].iters; }
    const double dt = probe_now() - t0;
    return dt > 0 ? (double) tot / dt : 0.0;
}
