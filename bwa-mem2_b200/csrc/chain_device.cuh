// chain_device.cuh — seed chaining and chain filtering of ONE read (device logic, one read per thread).
//
// Replaces mem_chain_seeds + test_and_merge (reference src/bwamem.cpp:806-974, :357-399),
// mem_chain_weight (:429-448), mem_chain_flt (:506-624) and bns_intv2rid (src/bntseq.cpp:378-402).
// The reference's per-read B-tree of chains becomes an index array ordered by chain position inside
// the read's private stripe of the flat seed arrays; chains hold their seeds as linked lists until
// the filter has decided which chains survive, then the survivors are written out contiguously.
#pragma once
#include "hd.h"
#include "bm2_b200.h"

struct ContigView { int64_t l_pac; int32_t n_seqs; const int64_t *ann_off; const int32_t *ann_len; const int32_t *ann_alt; };

BM2_HD int64_t bns_depos_d(const ContigView &c, int64_t pos) { return pos >= c.l_pac ? (c.l_pac << 1) - 1 - pos : pos; }
BM2_HD int bns_pos2rid_d(const ContigView &c, int64_t pos_f) {
    if (pos_f >= c.l_pac) return -1;
    int left = 0, mid = 0, right = c.n_seqs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= c.ann_off[mid]) {
            if (mid == c.n_seqs - 1) break;
            if (pos_f < c.ann_off[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
// is pos_f (forward-strand coordinate) inside contig rid?  (the answer bns_pos2rid_d would give, without the search)
BM2_HD bool bns_pos_in_rid_d(const ContigView &c, int64_t pos_f, int rid) {
    return rid >= 0 && pos_f < c.l_pac && pos_f >= c.ann_off[rid] && (rid == c.n_seqs - 1 || pos_f < c.ann_off[rid + 1]);
}
// bns_intv2rid (src/bntseq.cpp:378-402).  *hint: contig of the caller's previous interval (most seeds of a read fall into the
// same contig), tried before the binary search; the second end is tested against the contig of the first.
BM2_HD int bns_intv2rid_d(const ContigView &c, int64_t rb, int64_t re, int *hint = nullptr) {
    if (rb < c.l_pac && re > c.l_pac) return -2;
    const int64_t pb = bns_depos_d(c, rb);
    int rid_b = (hint && bns_pos_in_rid_d(c, pb, *hint)) ? *hint : bns_pos2rid_d(c, pb);
    if (hint && rid_b >= 0) *hint = rid_b;
    if (!(rb < re)) return rid_b;
    const int64_t pe = bns_depos_d(c, re - 1);
    const int rid_e = bns_pos_in_rid_d(c, pe, rid_b) ? rid_b : bns_pos2rid_d(c, pe);
    return rid_b == rid_e ? rid_b : -1;
}

// ks_introsort (src/ksort.h:185-232) restated over an index array: identical comparison/swap
// sequence, hence the identical order among ties.  lt(i, j) compares ELEMENTS i and j.
template <class T, class LT> BM2_HD void ks_insertsort_d(T *a, long s, long t, LT &lt) {
    for (long i = s + 1; i < t; ++i)
        for (long j = i; j > s && lt(a[j], a[j - 1]); --j) bm2_swap(a[j], a[j - 1]);
}
template <class T, class LT> BM2_HD void ks_combsort_d(T *a, long n, LT &lt) {
    const double shrink = 1.2473309501039786540366528676643;
    long gap = n; bool swapped;
    do {
        if (gap > 2) { gap = (long) (gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
        swapped = false;
        for (long i = 0; i + gap < n; ++i) if (lt(a[i + gap], a[i])) { bm2_swap(a[i], a[i + gap]); swapped = true; }
    } while (swapped || gap > 2);
    if (gap != 1) ks_insertsort_d(a, 0, n, lt);
}
template <class T, class LT> BM2_HD void ks_introsort_d(T *a, long n, LT lt) {
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) bm2_swap(a[0], a[1]); return; }
    int d; for (d = 2; (1ul << d) < (unsigned long) n; ++d) {}
    long st_l[72], st_r[72]; int st_d[72]; int top = 0;
    long s = 0, t = n - 1; d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { ks_combsort_d(a + s, t - s + 1, lt); t = s; continue; }
            long i = s, j = t, k = i + ((j - i) >> 1) + 1;
            if (lt(a[k], a[i])) { if (lt(a[k], a[j])) k = j; }
            else k = lt(a[j], a[i]) ? i : j;
            T rp = a[k];
            if (k != t) bm2_swap(a[k], a[t]);
            for (;;) {
                do ++i; while (lt(a[i], rp));
                do --j; while (i <= j && lt(rp, a[j]));
                if (j <= i) break;
                bm2_swap(a[i], a[j]);
            }
            bm2_swap(a[i], a[t]);
            if (i - s > t - i) {
                if (i - s > 16) { st_l[top] = s; st_r[top] = i - 1; st_d[top] = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { st_l[top] = i + 1; st_r[top] = t; st_d[top] = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) { ks_insertsort_d(a, 0, n, lt); return; }
            --top; s = st_l[top]; t = st_r[top]; d = st_d[top];
        }
    }
}

struct ChainParams {
    int w, max_chain_gap, max_occ, min_chain_weight, max_chain_extend, min_seed_len;
    float mask_level, drop_ratio;
};

// working records inside the read's stripe
struct WSeed { int64_t rbeg; int32_t qbeg, len, next; int32_t score; };
struct WChain {
    int64_t pos;
    int64_t first_rbeg, last_rbeg;
    int32_t first_qbeg, last_qbeg, last_len;
    int32_t head, tail, n;          // linked list of WSeed
    int32_t rid, is_alt;
    int32_t w, kept, first;
};

struct FltRec { int32_t beg, end, w, is_alt, first, kept; };   // what mem_chain_flt scans, by sorted position

struct ChainStripe {               // all arrays have >= n_slots entries, private to the read
    WSeed *seeds;
    WChain *chains;
    int32_t *ord;                  // chain ids ordered by pos (the B-tree's in-order sequence); bits 28-31: the key's level in the tree
    int64_t *ordpos;               // pos of ord[i] (contiguous copy for the binary search)
    int32_t *srt;                  // filter: chain ids sorted by weight
    int32_t *kv;                   // filter: kept non-overlapping chains
    FltRec *flt;                   // filter: compact records by sorted position
};


// ---- the shape of the reference's chain tree ---------------------------------------------------------------------------------
// The reference keeps the chains of a read in a B-tree keyed by pos (KBTREE_INIT(chn, ...), src/bwamem.cpp:40-41; src/kbtree.h)
// and the tree accepts equal keys.  Which of several chains with the same pos kb_intervalp meets first, and next to which of them
// kb_putp places a new one, depends on the shape of the tree; reads inside short tandem repeats produce such chains.  The shape is
// kept without building nodes: the keys stay in ONE array in key order (ord / ordpos) and every key carries the level of the node
// that holds it (0 = leaf).  A node of level h is then the run of level-h keys between two neighbouring keys of a higher level, its
// children are the stretches between its keys, and splitting a full node is "raise the level of its median key by one".
// Order t = 5 (at most 9 keys per node): kb_init(chn, KB_DEFAULT_SIZE + 8) with the 48-byte mem_chain_t
// (src/kbtree.h:64, src/bwamem.cpp:845, src/bwamem.h:126-133).
#define BM2_CHAIN_TREE_T 5
BM2_HD int chain_ord_id(int32_t v) { return (int) (v & 0x0fffffff); }
BM2_HD int chain_ord_lvl(int32_t v) { return (int) ((uint32_t) v >> 28); }

// The scans of the level-per-key tree and the one-element shift of an insertion are O(keys); a read with thousands of chains (10 kbp reads:
// every repeat hit opens a chain) spends its whole chaining time there when ONE thread runs them.  Coop abstracts them: ChainSolo = the plain
// loops (one thread per read, and the host build), ChainWarp = the same searches by ballots over 32 keys at a time and the shift in blocks of
// 32, for a warp whose 32 lanes all run chain_read_d on the same read with the same data (redundantly: every lane computes and writes the
// same values, so no lane depends on another's store except inside these helpers, which synchronise).
struct ChainSolo {
    BM2_HD int find_up(const int32_t *ord, int from, int hi, int h) const { int b = from; while (b < hi && (int) ((uint32_t) ord[b] >> 28) != h) ++b; return b; }
    BM2_HD int find_down(const int32_t *ord, int from, int lo, int h) const { int a = from; while (a >= lo && (int) ((uint32_t) ord[a] >> 28) != h) --a; return a; }
    // number of keys of level lv in [clo, chi); *mth = index of the T-th of them (-1 if fewer)
    BM2_HD int count_level(const int32_t *ord, int clo, int chi, int lv, int T, int *mth) const {
        int cnt = 0, m = -1;
        for (int j = clo; j < chi; ++j) if ((int) ((uint32_t) ord[j] >> 28) == lv && ++cnt == T) m = j;
        *mth = m; return cnt;
    }
    BM2_HD void shift_up(int32_t *ord, int64_t *ordpos, int at, int n) const { for (int k = n; k > at; --k) { ord[k] = ord[k - 1]; ordpos[k] = ordpos[k - 1]; } }
    BM2_HD void sync() const {}
};
#if defined(__CUDACC__)
struct ChainWarp {
    int lane;
    BM2_D int find_up(const int32_t *ord, int from, int hi, int h) const {
        for (int base = from; base < hi; base += 32) {
            const int i = base + lane;
            const unsigned m = __ballot_sync(0xffffffffu, i < hi && (int) ((uint32_t) ord[i] >> 28) == h);
            if (m) return base + __ffs(m) - 1;
        }
        return hi;
    }
    BM2_D int find_down(const int32_t *ord, int from, int lo, int h) const {
        for (int base = from; base >= lo; base -= 32) {
            const int i = base - lane;
            const unsigned m = __ballot_sync(0xffffffffu, i >= lo && (int) ((uint32_t) ord[i] >> 28) == h);
            if (m) return base - (__ffs(m) - 1);
        }
        return lo - 1;
    }
    BM2_D int count_level(const int32_t *ord, int clo, int chi, int lv, int T, int *mth) const {
        int cnt = 0, m = -1;
        for (int base = clo; base < chi; base += 32) {
            const int j = base + lane;
            unsigned bm = __ballot_sync(0xffffffffu, j < chi && (int) ((uint32_t) ord[j] >> 28) == lv);
            const int c = __popc(bm);
            if (m < 0 && cnt + c >= T) { for (int s = 0; s < T - cnt - 1; ++s) bm &= bm - 1; m = base + __ffs(bm) - 1; }
            cnt += c;
        }
        *mth = m; return cnt;
    }
    BM2_D void shift_up(int32_t *ord, int64_t *ordpos, int at, int n) const {       // [at, n) moves up by one, from the top, 32 keys per step
        __syncwarp();
        for (int top = n - 1; top >= at; top -= 32) {
            const int k = top - lane;
            int32_t vo = 0; int64_t vp = 0;
            if (k >= at) { vo = ord[k]; vp = ordpos[k]; }
            __syncwarp();
            if (k >= at) { ord[k + 1] = vo; ordpos[k + 1] = vp; }
            __syncwarp();
        }
    }
    BM2_D void sync() const { __syncwarp(); }
};
#endif

// kb_intervalp (src/kbtree.h:158-175) when a key equal to k exists; q = first index with ordpos >= k (so ordpos[q] == k).
// Walks from the root: in each node the first key >= k (src/kbtree.h:125-139); an equal one ends the search, else descend
// into the child on its left.  Returns the index of the key met.
template <class Coop = ChainSolo>
BM2_HD int chain_tree_equal_d(const int32_t *ord, const int64_t *ordpos, int n, int height, int64_t k, int q, const Coop &co = Coop()) {
    int lo = 0, hi = n;                                  // the subtree spans [lo, hi)
    for (int h = height; h > 0; --h) {
        const int b = co.find_up(ord, q, hi, h);
        if (b < hi && ordpos[b] == k) return b;
        const int a = co.find_down(ord, q - 1, lo, h);
        lo = a + 1; hi = b;
    }
    return q;
}

// kb_putp (src/kbtree.h:181-235): splits every full node on the way down, returns the index at which the new key (level 0)
// goes.  q = first index with ordpos >= k; height / root_n = level and key count of the root, updated here.
template <class Coop = ChainSolo>
BM2_HD int chain_tree_put_d(int32_t *ord, const int64_t *ordpos, int n, int &height, int &root_n, int64_t k, int q, const Coop &co = Coop()) {
    const int full = 2 * BM2_CHAIN_TREE_T - 1;
    if (root_n == full) {                                // new root above the old one: the old root's median moves up
        int m = -1;
        co.count_level(ord, 0, n, height, BM2_CHAIN_TREE_T, &m);
        ++height;
        co.sync();
        ord[m] = (int32_t) ((uint32_t) chain_ord_id(ord[m]) | (uint32_t) height << 28);
        co.sync();
        root_n = 1;
    }
    int lo = 0, hi = n;
    for (int h = height; h > 0; --h) {
        const int qr = q > lo ? q : lo;                  // first index of the subtree with ordpos >= k (hi if none)
        const int b = co.find_up(ord, qr, hi, h);        // first key of this node >= k
        int clo, chi;
        if (b < hi && ordpos[b] == k) {                  // equal: the child on its RIGHT (src/kbtree.h:210)
            clo = b + 1; chi = co.find_up(ord, b + 1, hi, h);
        } else {
            const int a = co.find_down(ord, qr - 1, lo, h);
            clo = a + 1; chi = b;
        }
        int m = -1;
        const int cnt = co.count_level(ord, clo, chi, h - 1, BM2_CHAIN_TREE_T, &m);
        if (cnt == full) {                               // split the child: its median joins this node
            co.sync();
            ord[m] = (int32_t) ((uint32_t) chain_ord_id(ord[m]) | (uint32_t) h << 28);
            co.sync();
            if (h == height) ++root_n;
            if (k > ordpos[m]) clo = m + 1; else chi = m;
        }
        lo = clo; hi = chi;
    }
    if (height == 0) ++root_n;
    const int qr = q > lo ? (q < hi ? q : hi) : lo;
    return (qr < hi && ordpos[qr] == k) ? qr + 1 : qr;   // leaf: after the first equal key, else before the first larger one
}

// test_and_merge (src/bwamem.cpp:357-399) on the cached first/last seed of the chain
BM2_HD bool chain_test_and_merge(const ChainParams &p, int64_t l_pac, WChain &c, WSeed *seeds, int seed_id, int seed_rid) {
    const WSeed &s = seeds[seed_id];
    const int64_t qend = c.last_qbeg + c.last_len, rend = c.last_rbeg + c.last_len;
    if (seed_rid != c.rid) return false;
    if (s.qbeg >= c.first_qbeg && s.qbeg + s.len <= qend && s.rbeg >= c.first_rbeg && s.rbeg + s.len <= rend) return true;
    if ((c.last_rbeg < l_pac || c.first_rbeg < l_pac) && s.rbeg >= l_pac) return false;
    const int64_t x = s.qbeg - c.last_qbeg, y = s.rbeg - c.last_rbeg;
    if (y >= 0 && x - y <= p.w && y - x <= p.w && x - c.last_len < p.max_chain_gap && y - c.last_len < p.max_chain_gap) {
        seeds[c.tail].next = seed_id; c.tail = seed_id; ++c.n;
        c.last_qbeg = s.qbeg; c.last_rbeg = s.rbeg; c.last_len = s.len;
        return true;
    }
    return false;
}

BM2_HD int chain_weight_d(const WChain &c, const WSeed *seeds) {
    int64_t end = 0; int w = 0, tmp;
    for (int i = c.head, k = 0; k < c.n; ++k, i = seeds[i].next) {
        const WSeed &s = seeds[i];
        if (s.qbeg >= end) w += s.len; else if (s.qbeg + s.len > end) w += (int) (s.qbeg + s.len - end);
        end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
    }
    tmp = w; w = 0; end = 0;
    for (int i = c.head, k = 0; k < c.n; ++k, i = seeds[i].next) {
        const WSeed &s = seeds[i];
        if (s.rbeg >= end) w += s.len; else if (s.rbeg + s.len > end) w += (int) (s.rbeg + s.len - end);
        end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
    }
    w = w < tmp ? w : tmp;
    return w < (1 << 30) ? w : (1 << 30) - 1;
}

// Chains of one read.  smems[0..n_smem) ordered (m asc, n asc); sa[] holds, SMEM after SMEM, the
// reference positions of its sampled rows (count = min(s, max_occ)).  On return srt[0..n_kept) lists
// the surviving chains in the reference's output order (weight-sorted, kept != 0); returns n_kept.
// *frac_rep receives l_rep / l_seq.
template <class Coop = ChainSolo>
BM2_HD int chain_read_d(const ContigView &cv, const ChainParams &p, const bm2_smem *smems, int n_smem, const int64_t *sa,
                        int l_seq, ChainStripe &ws, float *frac_rep, const Coop &co = Coop())
{
    int b = 0, e = 0, l_rep = 0;
    for (int i = 0; i < n_smem; ++i) {
        const int sb = (int) smems[i].m, se = (int) smems[i].n + 1;
        if (smems[i].s <= p.max_occ) continue;
        if (sb > e) { l_rep += e - b; b = sb; e = se; }
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    *frac_rep = (float) l_rep / l_seq;

    int n_ch = 0, n_sd = 0, rid_hint = -1, tree_h = 0, root_n = 0;
    int64_t slot = 0;
    for (int i = 0; i < n_smem; ++i) {
        const bm2_smem &sm = smems[i];
        const int slen = (int) sm.n + 1 - (int) sm.m;
        const int64_t cnt = sm.s < p.max_occ ? sm.s : p.max_occ;
        for (int64_t t = 0; t < cnt; ++t, ++slot) {
            const int64_t rbeg = sa[slot];
            const int rid = bns_intv2rid_d(cv, rbeg, rbeg + slen, &rid_hint);
            if (rid < 0) continue;
            const int sid = n_sd;
            WSeed &s = ws.seeds[sid]; s.rbeg = rbeg; s.qbeg = (int) sm.m; s.len = slen; s.next = -1; s.score = slen;
            int lower = -1;
            int lo = 0;                           // first chain with pos >= rbeg
            if (n_ch) {
                int hi = n_ch;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (ws.ordpos[mid] < rbeg) lo = mid + 1; else hi = mid; }
                lower = (lo < n_ch && ws.ordpos[lo] == rbeg) ? chain_tree_equal_d(ws.ord, ws.ordpos, n_ch, tree_h, rbeg, lo, co) : lo - 1;
                if (lower >= 0) {
                    WChain &lc = ws.chains[chain_ord_id(ws.ord[lower])];
                    if (chain_test_and_merge(p, cv.l_pac, lc, ws.seeds, sid, rid)) {
                        if (lc.tail == sid) ++n_sd;      // appended: keeps its slot; contained: slot is reused
                        continue;
                    }
                }
            }
            ++n_sd;
            WChain &c = ws.chains[n_ch];
            c.pos = rbeg; c.first_rbeg = c.last_rbeg = rbeg; c.first_qbeg = c.last_qbeg = s.qbeg; c.last_len = slen;
            c.head = c.tail = sid; c.n = 1; c.rid = rid; c.is_alt = cv.ann_alt ? (cv.ann_alt[rid] != 0) : 0;
            c.w = 0; c.kept = 0; c.first = -1;
            const int at = chain_tree_put_d(ws.ord, ws.ordpos, n_ch, tree_h, root_n, rbeg, lo, co);
            co.shift_up(ws.ord, ws.ordpos, at, n_ch);
            ws.ord[at] = n_ch; ws.ordpos[at] = rbeg;      // a new key is always a leaf key (level 0)
            co.sync();
            ++n_ch;
        }
    }
    if (n_ch == 0) return 0;

    // ---- mem_chain_flt (src/bwamem.cpp:506-624) -------------------------------------------------------------
    int n = 0;
    for (int i = 0; i < n_ch; ++i) {
        WChain &c = ws.chains[chain_ord_id(ws.ord[i])];
        c.first = -1; c.kept = 0; c.w = chain_weight_d(c, ws.seeds);
        if (c.w >= p.min_chain_weight) ws.srt[n++] = chain_ord_id(ws.ord[i]);
    }
    if (n == 0) ws.srt[n++] = chain_ord_id(ws.ord[0]);     // reference quirk: range (0,1) is processed even when all were dropped
    {
        const WChain *chs = ws.chains;
        ks_introsort_d(ws.srt, (long) n, [chs](int x, int y) { return chs[x].w > chs[y].w; });
    }
    for (int i = 0; i < n; ++i) {
        const WChain &c = ws.chains[ws.srt[i]];
        FltRec f; f.beg = c.first_qbeg; f.end = c.last_qbeg + c.last_len; f.w = c.w; f.is_alt = c.is_alt; f.first = -1; f.kept = 0;
        ws.flt[i] = f;
    }
    int n_kv = 0;
    ws.flt[0].kept = 3; ws.kv[n_kv++] = 0;
    for (int i = 1; i < n; ++i) {
        const FltRec ai = ws.flt[i];
        int large_ovlp = 0, k;
        for (k = 0; k < n_kv; ++k) {
            const int j = ws.kv[k];
            const FltRec aj = ws.flt[j];
            const int b_max = aj.beg > ai.beg ? aj.beg : ai.beg;
            const int e_min = aj.end < ai.end ? aj.end : ai.end;
            if (e_min > b_max && (!aj.is_alt || ai.is_alt)) {
                const int li = ai.end - ai.beg, lj = aj.end - aj.beg;
                const int min_l = li < lj ? li : lj;
                if ((float) (e_min - b_max) >= (float) min_l * p.mask_level && min_l < p.max_chain_gap) {
                    large_ovlp = 1;
                    if (aj.first < 0) ws.flt[j].first = i;
                    if ((float) ai.w < (float) aj.w * p.drop_ratio && aj.w - ai.w >= p.min_seed_len << 1) break;
                }
            }
        }
        if (k == n_kv) { ws.kv[n_kv++] = i; ws.flt[i].kept = large_ovlp ? 2 : 3; }
    }
    for (int k = 0; k < n_kv; ++k) {
        const int f = ws.flt[ws.kv[k]].first;
        if (f >= 0) ws.flt[f].kept = 1;
    }
    int i, k;
    for (i = k = 0; i < n; ++i) {
        const int kept = ws.flt[i].kept;
        if (kept == 0 || kept == 3) continue;
        if (++k >= p.max_chain_extend) break;
    }
    for (; i < n; ++i) if (ws.flt[i].kept < 3) ws.flt[i].kept = 0;
    for (i = 0; i < n; ++i) { WChain &c = ws.chains[ws.srt[i]]; c.kept = ws.flt[i].kept; c.first = ws.flt[i].first; }
    int n_kept = 0;
    for (i = 0; i < n; ++i) if (ws.chains[ws.srt[i]].kept) ws.srt[n_kept++] = ws.srt[i];
    return n_kept;
}

// ---- mem_flt_chained_seeds (src/bwamem.cpp:472-504): only for reads long enough (>= 725 bp by default) ----------
struct SwParams { int a, o_del, e_del, o_ins, e_ins; int8_t mat[25]; };

// local Smith-Waterman score == kswr_t::score of ksw_align2/ksw_i16 (src/ksw.cpp:234-345); qlen, tlen < 200
BM2_HD int local_sw_score_d(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const SwParams &p) {
    const int oe_del = p.o_del + p.e_del, oe_ins = p.o_ins + p.e_ins;
    int H[201], E[201];
    for (int j = 0; j <= qlen; ++j) { H[j] = 0; E[j] = 0; }
    int gmax = 0;
    for (int i = 0; i < tlen; ++i) {
        int f = 0, diag = 0;
        const int tb = target[i];
        for (int j = 0; j < qlen; ++j) {
            int h = diag + p.mat[tb * 5 + query[j]];
            diag = H[j + 1];
            int e = E[j + 1];
            if (e > h) h = e;
            if (f > h) h = f;
            if (h < 0) h = 0;
            if (h > gmax) gmax = h;
            H[j + 1] = h;
            int t = h - oe_del; if (t < 0) t = 0;
            e -= p.e_del; if (e < 0) e = 0;
            E[j + 1] = e > t ? e : t;
            t = h - oe_ins; if (t < 0) t = 0;
            f -= p.e_ins; if (f < 0) f = 0;
            f = f > t ? f : t;
        }
    }
    return gmax;
}

// mem_seed_sw (src/bwamem.cpp:401-427)
BM2_HD int seed_sw_d(const ContigView &cv, const SwParams &p, const uint8_t *ref, int l_query, const uint8_t *query, const WSeed &s) {
    const int64_t l_pac = cv.l_pac;
    if (s.len >= 200) return -1;
    int qb = s.qbeg, qe = s.qbeg + s.len;
    int64_t rb = s.rbeg, re = s.rbeg + s.len;
    const int64_t mid = (rb + re) >> 1;
    qb -= 50; qb = qb > 0 ? qb : 0;
    qe += 50; qe = qe < l_query ? qe : l_query;
    rb -= 50; rb = rb > 0 ? rb : 0;
    re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
    if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
    if (qe - qb >= 200 || re - rb >= 200) return -1;
    {
        const int is_rev = mid >= l_pac;
        const int rid = bns_pos2rid_d(cv, bns_depos_d(cv, mid));
        int64_t far_beg = cv.ann_off[rid], far_end = far_beg + cv.ann_len[rid];
        if (is_rev) { const int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
        rb = rb > far_beg ? rb : far_beg;
        re = re < far_end ? re : far_end;
    }
    return local_sw_score_d(qe - qb, query + qb, (int) (re - rb), ref + rb, p);
}

// mem_flt_chained_seeds (src/bwamem.cpp:472-504) in three steps, so that a warp can share the local alignments of ONE read (long reads:
// hundreds of seeds per read, each a <= 200 x 200 local alignment - the chain stage's whole cost for 10 kbp reads):
//   list:  the seeds of the surviving chains, in chain order, into ws.ord (free after chain_read_d)        [one lane]
//   score: s.score = seed_sw_d(...) for every listed seed, independent of each other                        [all lanes, strided]
//   apply: drop the seeds whose score is below min_hsp (computed on the host with the reference's double arithmetic); seeds keep their
//          order, survivors get score = SW score (or len * a when the alignment was skipped)                [one lane]
BM2_HD int chain_flt_list_d(const ChainStripe &ws, int n_kept) {
    int T = 0;
    for (int k = 0; k < n_kept; ++k) {
        const WChain &c = ws.chains[ws.srt[k]];
        for (int t = 0, i = c.head; t < c.n; ++t, i = ws.seeds[i].next) ws.ord[T++] = i;
    }
    return T;
}

BM2_HD void chain_flt_score_d(const ContigView &cv, const SwParams &p, const uint8_t *ref, int l_query, const uint8_t *query, const ChainStripe &ws,
                              int T, int first, int step) {
    for (int t = first; t < T; t += step) { WSeed &s = ws.seeds[ws.ord[t]]; s.score = seed_sw_d(cv, p, ref, l_query, query, s); }
}

BM2_HD void chain_flt_apply_d(const SwParams &p, int min_hsp, const ChainStripe &ws, int n_kept) {
    for (int k = 0; k < n_kept; ++k) {
        WChain &c = ws.chains[ws.srt[k]];
        int prev = -1, kept = 0, i = c.head;
        for (int t = 0; t < c.n; ++t) {
            WSeed &s = ws.seeds[i];
            const int nxt = s.next;
            const int sc = s.score;
            if (sc < 0 || sc >= min_hsp) {
                s.score = sc < 0 ? s.len * p.a : sc;
                if (prev < 0) c.head = i; else ws.seeds[prev].next = i;
                prev = i; ++kept;
            }
            i = nxt;
        }
        c.n = kept;
        if (prev >= 0) { ws.seeds[prev].next = -1; c.tail = prev; }
    }
}

BM2_HD void chain_flt_seeds_d(const ContigView &cv, const SwParams &p, const uint8_t *ref, int l_query, const uint8_t *query,
                              int min_hsp, const ChainStripe &ws, int n_kept)
{
    const int T = chain_flt_list_d(ws, n_kept);
    chain_flt_score_d(cv, p, ref, l_query, query, ws, T, 0, 1);
    chain_flt_apply_d(p, min_hsp, ws, n_kept);
}

// Writes the surviving chains of one read contiguously (reference order) and counts the extension
// work they imply: one reg per seed, a left job iff qbeg > 0, a right job iff the seed does not
// reach the read end (src/bwamem.cpp:2229, :2324).
BM2_HD void chain_finalize_d(const ChainStripe &ws, int n_kept, float frac_rep, int seqid, int l_seq, bm2_chain *out_chain,
                             bm2_seed *out_seed, int *n_seed_out, int *n_left, int *n_right)
{
    int ns = 0, nl = 0, nr = 0;
    for (int k = 0; k < n_kept; ++k) {
        const WChain &c = ws.chains[ws.srt[k]];
        bm2_chain &o = out_chain[k];
        o.pos = c.pos; o.seqid = seqid; o.rid = c.rid; o.n_seeds = c.n; o.seed_off = ns;
        o.w = c.w; o.kept = c.kept; o.first = c.first; o.is_alt = c.is_alt; o.frac_rep = frac_rep; o._pad = 0;
        for (int i = c.head, t = 0; t < c.n; ++t, i = ws.seeds[i].next) {
            const WSeed &s = ws.seeds[i];
            bm2_seed &d = out_seed[ns++];
            d.rbeg = s.rbeg; d.qbeg = s.qbeg; d.len = s.len; d.score = s.score; d.chain = k;
            if (s.qbeg) ++nl;
            if (s.qbeg + s.len != l_seq) ++nr;
        }
    }
    *n_seed_out = ns; *n_left = nl; *n_right = nr;
}
