// sam_device.cuh — the SAM stage of ONE read pair after mate rescue (device logic, groundwork for SURVEY §8(f) items 1-3: no kernel
// launches this yet).
//
// Replaces, per pair, what mem_sam_pe does after its rescue block (reference src/bwamem_pair.cpp:414-552): mem_mark_primary_se
// (src/bwamem.cpp:1392-1468), mem_pair (src/bwamem_pair.cpp:285-346), the paired / unpaired MAPQ logic, mem_approx_mapq_se
// (src/bwamem.cpp:1470-1494), mem_reg2aln (:1732-1805, over gen_cigar_d of cigar_device.cuh), the record selection of mem_reg2sam
// (:1534-1560) and the columns of mem_aln2sam (:1592-1730) except the text.  Double-precision libm calls are replaced by tables the
// host fills with the SAME libm the reference uses (log of small integers; the insert-size term of mem_pair per orientation and
// distance), so that the device need not reproduce glibc's log / erfc bit for bit.
// tests/host_emul/sam_emul.cpp checks it against the oracle (bm2o_sam_pe), which equals the reference's SAM byte for byte.
#pragma once
#include "mate_device.cuh"
#include "cigar_device.cuh"

// why a pair could not be finished with the scratch it was given (bits of *overflow; the driver grows that buffer and re-runs)
#define BM2_OVF_LOG 1          // an argument of log() beyond the host-filled table
#define BM2_OVF_PAIR_TERM 2    // an insert size outside the host-filled pairing-term table
#define BM2_OVF_POOL 8         // CIGAR / MD storage of the records
#define BM2_OVF_RECORDS 16     // records per read (aa_cap)
#define BM2_OVF_KSW_LIST 32    // score2 candidates of one local alignment (bcap)
#define BM2_OVF_WINDOW 64      // a rescue window longer than MateScratch::tcap

struct SamParams {
    ExtParams ep;                 // a, b, gaps, w, mat
    int T, flag, min_seed_len, pen_unpaired;
    float mask_level, drop_ratio;
    float mapQ_coef_len; int mapQ_coef_fac;
    float XA_drop_ratio; int max_XA_hits, max_XA_hits_alt;
};
struct SamTables {
    const double *log_tab; int n_log;          // log_tab[k] = log((double) k), k < n_log (k = 0 unused)
    const double *pair_term[4]; int64_t pair_lo[4], pair_hi[4];   // pair_term[d][dist - lo] = .721 * log(2 * erfc(|dist - avg| / std * M_SQRT1_2)) * a
};
struct SamAln {                    // mem_aln_t subset
    int flag, rid, mapq, nm, score, sub, is_rev, is_alt, alt_sc, n_cigar, n_md;
    int reg;                      // index of the region it was made from (the key of its XA entries), -1: none
    int64_t pos;
    uint32_t *cigar; char *md;    // storage provided by the caller: l_query + rlen + 4 ops, 2*l_query + 7*rlen + 16 bytes
};
struct SamRec {                    // the columns of one SAM line (bm2o_samrec without the offsets)
    int flag, rid, mapq, rnext, nm, score, sub, n_cigar, n_md;
    int reg;                      // SamAln::reg
    int alt_sc;                   // > 0: the pa tag is score / alt_sc (src/bwamem.cpp:1713-1714; not printed on 0x100 records)
    int is_alt;                   // the record's region lies on an ALT contig
    int n_mc;                     // operations of the MC tag (the mate's CIGAR as this line prints it), stored after the record's own in ops[]
    int64_t pos, pnext, tlen;
};

BM2_HD uint64_t sam_hash64_d(uint64_t key) {           // hash_64 (src/utils.h:117-128)
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}
BM2_HD int sam_is_alt_d(const bm2_alnreg_t &a) { return (a.n_comp_is_alt >> 30) & 3; }

BM2_HD void sam_mark_primary_core_d(const SamParams &p, int n, bm2_alnreg_t *a, int32_t *z) {       // :1392-1418
    int tmp = p.ep.a + p.ep.b;
    tmp = p.ep.o_del + p.ep.e_del > tmp ? p.ep.o_del + p.ep.e_del : tmp;
    tmp = p.ep.o_ins + p.ep.e_ins > tmp ? p.ep.o_ins + p.ep.e_ins : tmp;
    int nz = 0;
    z[nz++] = 0;
    for (int i = 1; i < n; ++i) {
        int k;
        for (k = 0; k < nz; ++k) {
            const int j = z[k];
            const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * p.mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (sam_is_alt_d(a[j]) || !sam_is_alt_d(a[i]))) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i;
        else a[i].secondary = z[k];
    }
}

// mem_mark_primary_se (:1420-1468); z, idx: n ints each.  Returns n_pri.
BM2_HD int sam_mark_primary_se_d(const SamParams &p, int n, bm2_alnreg_t *a, int64_t id, int32_t *z, int32_t *idx) {
    if (n == 0) return 0;
    int n_pri = 0;
    for (int i = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = sam_hash64_d((uint64_t) (id + i));
        if (!sam_is_alt_d(a[i])) ++n_pri;
        idx[i] = i;
    }
    {
        const bm2_alnreg_t *ra = a;
        ks_introsort_d(idx, (long) n, [ra](int xi, int yi) {                                    // alnreg_hlt
            const bm2_alnreg_t &x = ra[xi], &y = ra[yi];
            return x.score > y.score || (x.score == y.score && (sam_is_alt_d(x) < sam_is_alt_d(y) || (sam_is_alt_d(x) == sam_is_alt_d(y) && x.hash < y.hash)));
        });
    }
    permute_regs_d(a, idx, n);
    sam_mark_primary_core_d(p, n, a, z);
    for (int i = 0; i < n; ++i) {
        bm2_alnreg_t *q = &a[i];
        q->secondary_all = i;
        if (!sam_is_alt_d(*q) && q->secondary >= 0 && sam_is_alt_d(a[q->secondary])) q->alt_sc = a[q->secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        if (n_pri > 0) {
            for (int i = 0; i < n; ++i) idx[i] = i;
            const bm2_alnreg_t *ra = a;
            ks_introsort_d(idx, (long) n, [ra](int xi, int yi) {                                // alnreg_hlt2
                const bm2_alnreg_t &x = ra[xi], &y = ra[yi];
                return sam_is_alt_d(x) < sam_is_alt_d(y) || (sam_is_alt_d(x) == sam_is_alt_d(y) && (x.score > y.score || (x.score == y.score && x.hash < y.hash)));
            });
            permute_regs_d(a, idx, n);
        }
        for (int i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (int i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (sam_is_alt_d(a[i])) a[i].secondary = 0x7fffffff;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {
            for (int i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            sam_mark_primary_core_d(p, n_pri, a, z);
        }
    } else {
        for (int i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    return n_pri;
}

BM2_HD int sam_mapq_se_d(const SamParams &p, const SamTables &tb, const bm2_alnreg_t *a, int *overflow) {     // mem_approx_mapq_se (:1470-1494)
    int mapq, l, sub = a->sub ? a->sub : p.min_seed_len * p.ep.a;
    sub = a->csub > sub ? a->csub : sub;
    if (sub >= a->score) return 0;
    l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int) (a->re - a->rb);
    const double identity = 1. - (double) (l * p.ep.a - a->score) / (p.ep.a + p.ep.b) / l;
    if (a->score == 0) mapq = 0;
    else if (p.mapQ_coef_len > 0) {
        if (l >= tb.n_log) { *overflow |= BM2_OVF_LOG; return 0; }
        double tmp = l < p.mapQ_coef_len ? 1. : p.mapQ_coef_fac / tb.log_tab[l];
        tmp *= identity * identity;
        mapq = (int) (6.02 * (a->score - sub) / p.ep.a * tmp * tmp + .499);
    } else {
        if (a->seedcov < 0 || a->seedcov >= tb.n_log) { *overflow |= BM2_OVF_LOG; return 0; }
        mapq = (int) (30.0 * (1. - (double) sub / a->score) * tb.log_tab[a->seedcov] + .499);
        mapq = identity < 0.95 ? (int) (mapq * identity * identity + .499) : mapq;
    }
    if (a->sub_n > 0) {
        if (a->sub_n + 1 >= tb.n_log) { *overflow |= BM2_OVF_LOG; return 0; }
        mapq -= (int) (4.343 * tb.log_tab[a->sub_n + 1] + .499);
    }
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int) (mapq * (1. - a->frac_rep) + .499);
    return mapq;
}

BM2_HD int sam_infer_bw_d(int l1, int l2, int score, int a, int q, int r) {                       // infer_bw (:1811-1818)
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    int w = (int) ((double) ((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    const int d = l1 > l2 ? l1 - l2 : l2 - l1;
    if (w < d) w = d;
    return w;
}

// mem_reg2aln (:1732-1805).  ar == null: an unmapped record.  he: 2*(l_query+1) ints; z: backtrack cells for the widest band.
BM2_HD void sam_reg2aln_d(const SamParams &p, const SamTables &tb, const ContigView &cv, const uint8_t *ref, int l_query, const uint8_t *query,
                          const bm2_alnreg_t *ar, int32_t *he, const CigarZ &z, SamAln *a, int *overflow)
{
    a->reg = -1; a->flag = 0; a->rid = -1; a->mapq = 0; a->nm = 0; a->score = 0; a->sub = 0; a->is_rev = 0; a->is_alt = 0; a->alt_sc = 0; a->n_cigar = 0; a->n_md = 0; a->pos = -1;
    if (ar == 0 || ar->rb < 0 || ar->re < 0) { a->flag |= 0x4; return; }
    const int qb = ar->qb, qe = ar->qe;
    const int64_t rb = ar->rb, re = ar->re;
    a->mapq = ar->secondary < 0 ? sam_mapq_se_d(p, tb, ar, overflow) : 0;
    if (ar->secondary >= 0) a->flag |= 0x100;
    int tmp = sam_infer_bw_d(qe - qb, (int) (re - rb), ar->truesc, p.ep.a, p.ep.o_del, p.ep.e_del);
    int w2 = sam_infer_bw_d(qe - qb, (int) (re - rb), ar->truesc, p.ep.a, p.ep.o_ins, p.ep.e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > p.ep.w) w2 = w2 < ar->w ? w2 : ar->w;
    CigarParams cp; for (int k = 0; k < 25; ++k) cp.mat[k] = p.ep.mat[k];
    cp.o_del = p.ep.o_del; cp.e_del = p.ep.e_del; cp.o_ins = p.ep.o_ins; cp.e_ins = p.ep.e_ins;
    int i = 0, score = 0, NM = -1, nc = 0, nmd = 0, last_sc = -(1 << 30);
    do {
        w2 = w2 < p.ep.w << 2 ? w2 : p.ep.w << 2;
        gen_cigar_d(cp, cv.l_pac, ref, w2, qe - qb, query + qb, rb, re, he, z, &score, a->cigar + 1, &nc, &NM, a->md, &nmd);   // +1: room for a 5' clip
        if (score == last_sc || w2 == p.ep.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar->truesc - p.ep.a);
    a->nm = NM; a->n_md = nmd;
    const int is_rev = (rb < cv.l_pac ? rb : re - 1) >= cv.l_pac;
    int64_t pos = bns_depos_d(cv, rb < cv.l_pac ? rb : re - 1);
    a->is_rev = is_rev;
    uint32_t *c = a->cigar + 1;
    if (nc > 0) {                                                     // squeeze out a leading or trailing deletion
        if ((c[0] & 0xf) == 2) { pos += c[0] >> 4; ++c; --nc; }
        else if ((c[nc - 1] & 0xf) == 2) --nc;
    }
    if (qb != 0 || qe != l_query) {
        const int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
        if (clip5) { --c; c[0] = (uint32_t) clip5 << 4 | 3; ++nc; }
        if (clip3) c[nc++] = (uint32_t) clip3 << 4 | 3;
    }
    if (c != a->cigar) for (int k = 0; k < nc; ++k) a->cigar[k] = c[k];
    a->n_cigar = nc;
    a->rid = bns_pos2rid_d(cv, pos);
    a->pos = pos - cv.ann_off[a->rid];
    a->score = ar->score; a->sub = ar->sub > ar->csub ? ar->sub : ar->csub;
    a->is_alt = sam_is_alt_d(*ar); a->alt_sc = ar->alt_sc;
}

struct SamP64 { uint64_t x, y; };

// mem_pair (src/bwamem_pair.cpp:285-346).  v: n_pri[0] + n_pri[1] entries.  The reference collects every candidate pair in a growing
// array `u`, sorts it and reads off the best entry, the score of the second best and the number of entries within a small score
// distance of the second best.  A read pair inside a tandem repeat has 10^5 candidates, so nothing is stored here: the candidates
// are enumerated twice (best two, then the count); the keys are the reference's, so the order - hash tie-break included - is too.
template <class F>
BM2_HD void sam_pair_enum_d(const SamTables &tb, const MatePes &pes, const SamP64 *v, int nv, int id, int *overflow, F &f) {
    int y[4] = { -1, -1, -1, -1 };
    for (int i = 0; i < nv; ++i) {
        for (int r = 0; r < 2; ++r) {
            const int dir = r << 1 | (int) (v[i].y >> 1 & 1);
            if (pes.failed[dir]) continue;
            const int which = r << 1 | (int) ((v[i].y & 1) ^ 1);
            if (y[which] < 0) continue;
            for (int k = y[which]; k >= 0; --k) {
                if ((int) (v[k].y & 3) != which) continue;
                const int64_t dist = (int64_t) v[i].x - (int64_t) v[k].x;
                if (dist > pes.high[dir]) break;
                if (dist < pes.low[dir]) continue;
                if (dist < tb.pair_lo[dir] || dist > tb.pair_hi[dir]) { *overflow |= BM2_OVF_PAIR_TERM; continue; }
                int q = (int) ((double) ((v[i].y >> 32) + (v[k].y >> 32)) + tb.pair_term[dir][dist - tb.pair_lo[dir]] + .499);
                if (q < 0) q = 0;
                SamP64 e;
                e.y = (uint64_t) k << 32 | (uint64_t) i;
                e.x = (uint64_t) q << 32 | (sam_hash64_d(e.y ^ (uint64_t) (int64_t) (id << 8)) & 0xffffffffU);
                f(e);
            }
        }
        y[v[i].y & 3] = i;
    }
}

BM2_HD int sam_pair_d(const SamParams &p, const SamTables &tb, const ContigView &cv, const MatePes &pes, bm2_alnreg_t *const a[2], int id, int *sub, int *n_sub,
                      int z[2], const int n_pri[2], SamP64 *v, int *overflow)
{
    const int64_t l_pac = cv.l_pac;
    int nv = 0;
    for (int r = 0; r < 2; ++r)
        for (int i = 0; i < n_pri[r]; ++i) {
            const bm2_alnreg_t *e = &a[r][i];
            SamP64 key;
            key.x = (uint64_t) (e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb);
            key.x = (uint64_t) e->rid << 32 | (key.x - (uint64_t) cv.ann_off[e->rid]);
            key.y = (uint64_t) e->score << 32 | (uint64_t) (i << 2 | (e->rb >= l_pac) << 1 | r);
            v[nv++] = key;
        }
    auto lt = [](const SamP64 &s, const SamP64 &t) { return s.x < t.x || (s.x == t.x && s.y < t.y); };
    ks_introsort_d(v, (long) nv, lt);
    long long nu = 0;
    SamP64 best; best.x = 0; best.y = 0;
    int second = 0;                                        // score of the second entry from the top of the sorted list
    auto top2 = [&](const SamP64 &e) {
        if (nu == 0 || lt(best, e)) { if (nu) second = (int) (best.x >> 32); best = e; }
        else if (nu == 1 || (int) (e.x >> 32) > second) second = (int) (e.x >> 32);
        ++nu;
    };
    sam_pair_enum_d(tb, pes, v, nv, id, overflow, top2);
    *sub = 0; *n_sub = 0;
    if (nu == 0) return 0;
    const int i = (int) (best.y >> 32), k = (int) (best.y << 32 >> 32);
    z[v[i].y & 1] = (int) (v[i].y << 32 >> 34);
    z[v[k].y & 1] = (int) (v[k].y << 32 >> 34);
    if (nu > 1) {
        int tmp = p.ep.a + p.ep.b;
        tmp = tmp > p.ep.o_del + p.ep.e_del ? tmp : p.ep.o_del + p.ep.e_del;
        tmp = tmp > p.ep.o_ins + p.ep.e_ins ? tmp : p.ep.o_ins + p.ep.e_ins;
        *sub = second;
        long long cnt = 0;                                 // entries below the top one with sub - score <= tmp
        const int sb = second;
        auto count = [&](const SamP64 &e) { if (sb - (int) (e.x >> 32) <= tmp) ++cnt; };
        sam_pair_enum_d(tb, pes, v, nv, id, overflow, count);
        if (sb - (int) (best.x >> 32) <= tmp) --cnt;       // the top entry itself
        *n_sub = cnt > 0x7fffffff ? 0x7fffffff : (int) cnt;
    }
    return (int) (best.x >> 32);
}

BM2_HD int sam_rlen_d(const SamAln &a) { int l = 0; for (int k = 0; k < a.n_cigar; ++k) { const int op = a.cigar[k] & 0xf; if (op == 0 || op == 2) l += a.cigar[k] >> 4; } return l; }
BM2_HD int sam_raw_mapq_d(int diff, int a) { return (int) (6.02 * diff / a + .499); }

// The columns of mem_aln2sam (src/bwamem.cpp:1592-1640) for p (the which-th record of its read) with mate m (may be null).  The printed
// CIGAR letters are returned in ops[0..rec->n_cigar) as len << 4 | index into "MIDSH", followed by the rec->n_mc operations of the MC tag.
BM2_HD void sam_aln2rec_d(const SamParams &prm, const SamAln &p_, int which, const SamAln *m_, SamRec *r, uint32_t *ops)
{
    int flag = p_.flag, rid = p_.rid, is_rev = p_.is_rev, n_cigar = p_.n_cigar; int64_t pos = p_.pos;
    int m_rid = -1, m_is_rev = 0, m_ncig = 0; int64_t m_pos = -1;
    if (m_) { m_rid = m_->rid; m_is_rev = m_->is_rev; m_ncig = m_->n_cigar; m_pos = m_->pos; }
    flag |= m_ ? 0x1 : 0;
    flag |= rid < 0 ? 0x4 : 0;
    flag |= m_ && m_rid < 0 ? 0x8 : 0;
    if (rid < 0 && m_ && m_rid >= 0) { rid = m_rid; pos = m_pos; is_rev = m_is_rev; n_cigar = 0; }
    if (m_ && m_rid < 0 && rid >= 0) { m_rid = rid; m_pos = pos; m_is_rev = is_rev; m_ncig = 0; }
    flag |= is_rev ? 0x10 : 0;
    flag |= m_ && m_is_rev ? 0x20 : 0;
    r->flag = (flag & 0xffff) | (flag & 0x10000 ? 0x100 : 0);
    r->rid = rid; r->pos = rid >= 0 ? pos + 1 : 0; r->mapq = rid >= 0 ? p_.mapq : 0;
    r->n_cigar = 0;
    if (rid >= 0)
        for (int k = 0; k < n_cigar; ++k) {
            int c = p_.cigar[k] & 0xf;
            if (!(prm.flag & 0x200) && !p_.is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
            ops[r->n_cigar++] = (p_.cigar[k] >> 4) << 4 | (uint32_t) c;
        }
    r->rnext = -1; r->pnext = 0; r->tlen = 0;
    if (m_ && m_rid >= 0) {
        r->rnext = m_rid; r->pnext = m_pos + 1;
        if (rid == m_rid) {
            SamAln pc = p_; pc.n_cigar = n_cigar;
            const int64_t p0 = pos + (is_rev ? sam_rlen_d(pc) - 1 : 0);
            int64_t p1 = m_pos;
            if (m_is_rev && m_ncig) { SamAln mc = *m_; mc.n_cigar = m_ncig; p1 += sam_rlen_d(mc) - 1; } else if (m_is_rev) p1 += -1;
            if (m_ncig == 0 || n_cigar == 0) r->tlen = 0;
            else r->tlen = -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0));
        }
    }
    r->nm = n_cigar ? p_.nm : 0; r->n_md = n_cigar ? p_.n_md : 1;
    r->score = p_.score; r->sub = p_.sub; r->reg = p_.reg; r->alt_sc = p_.alt_sc; r->is_alt = p_.is_alt;
    r->n_mc = 0;                                              // MC:Z: = add_cigar(opt, m, str, which) (src/bwamem.cpp:1686, :1579-1590)
    if (m_ && m_ncig)
        for (int k = 0; k < m_ncig; ++k) {
            int c = m_->cigar[k] & 0xf;
            if (!(prm.flag & 0x200) && !m_->is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
            ops[r->n_cigar + r->n_mc++] = (m_->cigar[k] >> 4) << 4 | (uint32_t) c;
        }
}

// mem_reorder_primary5 (src/bwamem.cpp:1496-1518), option -5: the primary with the smallest query start becomes record 0
BM2_HD void sam_reorder_primary5_d(int T, int n, bm2_alnreg_t *a) {
    int n_pri = 0, left_st = 0x7fffffff, left_k = -1;
    for (int k = 0; k < n; ++k) if (a[k].secondary < 0 && !sam_is_alt_d(a[k]) && a[k].score >= T) ++n_pri;
    if (n_pri <= 1) return;
    for (int k = 0; k < n; ++k) {
        const bm2_alnreg_t &q = a[k];
        if (q.secondary >= 0 || sam_is_alt_d(q) || q.score < T) continue;
        if (q.qb < left_st) { left_st = q.qb; left_k = k; }
    }
    if (left_k == 0) return;
    { const bm2_alnreg_t t = a[0]; a[0] = a[left_k]; a[left_k] = t; }
    for (int k = 1; k < n; ++k) {
        bm2_alnreg_t &q = a[k];
        if (q.secondary == 0) q.secondary = left_k; else if (q.secondary == left_k) q.secondary = 0;
        if (q.secondary_all == 0) q.secondary_all = left_k; else if (q.secondary_all == left_k) q.secondary_all = 0;
    }
}

// mem_gen_alt (src/bwamem_extra.cpp:130-183): the entries of the XA tags of one read, in the reference's order.  An entry belongs to
// the region r = secondary_all of the hit it describes and is printed with every record made from region r (SamAln::reg).
// emit(r, t): t.rid, t.is_rev, t.pos, t.cigar[0..n_cigar), t.nm are the fields of the entry; t's storage is reused by the next one.
// cnt / has_alt: n ints each.  Call after mem_mark_primary_se (and after the paired path's secondary_all fix-up).
template <class EmitXa>
BM2_HD void sam_gen_alt_d(const SamParams &p, const SamTables &tb, const ContigView &cv, const uint8_t *ref, int l_query, const uint8_t *query,
                          const bm2_alnreg_t *a, int n, int32_t *cnt, int32_t *has_alt, int32_t *he, const CigarZ &zz, SamAln &t, EmitXa &emit, int *overflow)
{
    const double drop = p.XA_drop_ratio;                  // get_pri_idx takes the float option as a double (src/bwamem_extra.cpp:122)
    int tot = 0;
    for (int i = 0; i < n; ++i) { cnt[i] = 0; has_alt[i] = 0; }
    for (int i = 0; i < n; ++i) {
        const int k = a[i].secondary_all;
        if (k < 0 || !(a[i].score >= a[k].score * drop)) continue;
        ++cnt[k]; ++tot;
        if (sam_is_alt_d(a[i])) has_alt[k] = 1;
    }
    if (tot == 0) return;
    for (int i = 0; i < n; ++i) {
        const int r = a[i].secondary_all;
        if (r < 0 || !(a[i].score >= a[r].score * drop)) continue;
        if (cnt[r] > p.max_XA_hits_alt || (!has_alt[r] && cnt[r] > p.max_XA_hits)) continue;
        sam_reg2aln_d(p, tb, cv, ref, l_query, query, &a[i], he, zz, &t, overflow);
        emit(r, t);
    }
}

// Scratch of one pair for the SAM stage.
struct SamScratch {
    int32_t *z, *idx;              // max(n0, n1) + 4 ints each
    SamP64 *v;                     // n0 + n1 entries
    int32_t *he; CigarZ zz;        // global alignment: 2 * (max read length + 1) ints; backtrack cells
    SamAln *aa[2]; int aa_cap;     // per read: records to print (regions + 2)
    uint32_t *cig_pool; long long cig_cap; char *md_pool; long long md_cap;      // storage of the records' CIGAR / MD
    uint32_t *ops;                 // printed CIGAR of one record + its MC tag (longest record of each read)
};

struct SamPool { uint32_t *c; long long cc, cu; char *m; long long mc, mu; };
BM2_HD bool sam_alloc_d(SamPool &pl, int l_query, const bm2_alnreg_t *ar, SamAln *a) {
    const long long rlen = ar && ar->re > ar->rb ? ar->re - ar->rb : 0;
    const long long nc = l_query + rlen + 4, nm = 2LL * l_query + 7 * rlen + 16;
    if (pl.cu + nc > pl.cc || pl.mu + nm > pl.mc) return false;
    a->cigar = pl.c + pl.cu; a->md = pl.m + pl.mu; pl.cu += nc; pl.mu += nm;
    return true;
}

// mem_sam_pe after the rescue block (src/bwamem_pair.cpp:414-552) for one pair whose regions a[i][0..n[i]) already went through mate
// rescue.  emit(read_in_pair, record index, SamRec, printed ops, md pointer) is called once per SAM line in output order;
// emit_xa(read_in_pair, region, SamAln) once per XA entry (before the records; SamRec::reg of a record names its entries' region).
template <class Emit, class EmitXa>
BM2_HD void sam_pe_pair_d(const SamParams &p, const SamTables &tb, const ContigView &cv, const MatePes &pes, const uint8_t *ref,
                          const uint8_t *const seq[2], const int l_seq[2], bm2_alnreg_t *const a[2], const int n[2], int id, const SamScratch &sc,
                          Emit &emit, EmitXa &emit_xa, int *overflow)
{
    SamPool pl = { sc.cig_pool, sc.cig_cap, 0, sc.md_pool, sc.md_cap, 0 };
    int extra_flag = 1, n_pri[2], z[2] = { 0, 0 }, o = 0, subo = 0, n_sub = 0, n_aa[2] = { 0, 0 };
    n_pri[0] = sam_mark_primary_se_d(p, n[0], a[0], (int64_t) id << 1 | 0, sc.z, sc.idx);
    n_pri[1] = sam_mark_primary_se_d(p, n[1], a[1], (int64_t) id << 1 | 1, sc.z, sc.idx);
    if (p.flag & 0x800) { sam_reorder_primary5_d(p.T, n[0], a[0]); sam_reorder_primary5_d(p.T, n[1], a[1]); }      // MEM_F_PRIMARY5 (src/bwamem_pair.cpp:420-423)
    bool paired = false;
    SamAln h[2];
    auto out_all = [&]() {
        if (!(p.flag & 0x8))                                  // !MEM_F_ALL: XA entries (src/bwamem_pair.cpp:484-487, src/bwamem.cpp:1529-1530)
            for (int i = 0; i < 2; ++i) {
                if (n[i] == 0) continue;
                bm2_alnreg_t widest; widest.rb = 0; widest.re = 0;
                for (int k = 0; k < n[i]; ++k) if (a[i][k].re - a[i][k].rb > widest.re) widest.re = a[i][k].re - a[i][k].rb;
                SamAln t;
                if (!sam_alloc_d(pl, l_seq[i], &widest, &t)) { *overflow |= BM2_OVF_POOL; return; }
                auto one = [&](int r, const SamAln &e) { emit_xa(i, r, e); };
                sam_gen_alt_d(p, tb, cv, ref, l_seq[i], seq[i], a[i], n[i], sc.z, sc.idx, sc.he, sc.zz, t, one, overflow);
            }
        for (int i = 0; i < 2; ++i)
            for (int k = 0; k < n_aa[i]; ++k) {
                SamRec r;
                sam_aln2rec_d(p, sc.aa[i][k], k, &h[!i], &r, sc.ops);
                emit(i, k, r, sc.ops, sc.aa[i][k].n_cigar ? sc.aa[i][k].md : "");
            }
    };
    if (!(p.flag & 0x4) && n_pri[0] && n_pri[1] &&
        (o = sam_pair_d(p, tb, cv, pes, a, id, &subo, &n_sub, z, n_pri, sc.v, overflow)) > 0) {
        int is_multi[2];
        for (int i = 0; i < 2; ++i) {
            int j;
            for (j = 1; j < n_pri[i]; ++j) if (a[i][j].secondary < 0 && a[i][j].score >= p.T) break;
            is_multi[i] = j < n_pri[i] ? 1 : 0;
        }
        if (!(is_multi[0] || is_multi[1])) {
            paired = true;
            int q_pe, q_se[2];
            const int score_un = a[0][0].score + a[1][0].score - p.pen_unpaired;
            subo = subo > score_un ? subo : score_un;
            q_pe = sam_raw_mapq_d(o - subo, p.ep.a);
            if (n_sub > 0) { if (n_sub + 1 >= tb.n_log) *overflow |= BM2_OVF_LOG; else q_pe -= (int) (4.343 * tb.log_tab[n_sub + 1] + .499); }
            if (q_pe < 0) q_pe = 0;
            if (q_pe > 60) q_pe = 60;
            q_pe = (int) (q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
            if (o > score_un) {
                bm2_alnreg_t *c[2] = { &a[0][z[0]], &a[1][z[1]] };
                for (int i = 0; i < 2; ++i) {
                    if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
                    q_se[i] = sam_mapq_se_d(p, tb, c[i], overflow);
                }
                q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
                q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
                extra_flag |= 2;
                q_se[0] = q_se[0] < sam_raw_mapq_d(c[0]->score - c[0]->csub, p.ep.a) ? q_se[0] : sam_raw_mapq_d(c[0]->score - c[0]->csub, p.ep.a);
                q_se[1] = q_se[1] < sam_raw_mapq_d(c[1]->score - c[1]->csub, p.ep.a) ? q_se[1] : sam_raw_mapq_d(c[1]->score - c[1]->csub, p.ep.a);
            } else {
                z[0] = z[1] = 0;
                q_se[0] = sam_mapq_se_d(p, tb, &a[0][0], overflow);
                q_se[1] = sam_mapq_se_d(p, tb, &a[1][0], overflow);
            }
            for (int i = 0; i < 2; ++i) {
                const int k = a[i][z[i]].secondary_all;
                if (k >= 0 && k < n_pri[i]) {
                    for (int j = 0; j < n[i]; ++j)
                        if (a[i][j].secondary_all == k || j == k) a[i][j].secondary_all = z[i];
                    a[i][z[i]].secondary_all = -1;
                }
            }
            for (int i = 0; i < 2; ++i) {
                if (!sam_alloc_d(pl, l_seq[i], &a[i][z[i]], &h[i])) { *overflow |= BM2_OVF_POOL; return; }
                sam_reg2aln_d(p, tb, cv, ref, l_seq[i], seq[i], &a[i][z[i]], sc.he, sc.zz, &h[i], overflow);
                h[i].mapq = q_se[i]; h[i].reg = z[i];
                h[i].flag |= 0x40 << i | extra_flag;
                sc.aa[i][n_aa[i]++] = h[i];
                if (n_pri[i] < n[i]) {
                    const bm2_alnreg_t *q = &a[i][n_pri[i]];
                    if (q->score < p.T || q->secondary >= 0 || !sam_is_alt_d(*q)) continue;
                    SamAln g;
                    if (!sam_alloc_d(pl, l_seq[i], q, &g)) { *overflow |= BM2_OVF_POOL; return; }
                    sam_reg2aln_d(p, tb, cv, ref, l_seq[i], seq[i], q, sc.he, sc.zz, &g, overflow);
                    g.flag |= 0x800 | 0x40 << i | extra_flag; g.reg = n_pri[i];
                    sc.aa[i][n_aa[i]++] = g;
                }
            }
            out_all();
        }
    }
    if (paired) return;
    // no_pairing (:523-551)
    for (int i = 0; i < 2; ++i) {
        int which = -1;
        if (n[i]) {
            if (a[i][0].score >= p.T) which = 0;
            else if (n_pri[i] < n[i] && a[i][n_pri[i]].score >= p.T) which = n_pri[i];
        }
        if (!sam_alloc_d(pl, l_seq[i], which >= 0 ? &a[i][which] : 0, &h[i])) { *overflow |= BM2_OVF_POOL; return; }
        sam_reg2aln_d(p, tb, cv, ref, l_seq[i], seq[i], which >= 0 ? &a[i][which] : 0, sc.he, sc.zz, &h[i], overflow);
    }
    if (!(p.flag & 0x4) && h[0].rid == h[1].rid && h[0].rid >= 0) {
        int64_t dist;
        const int d = mate_infer_dir_d(cv.l_pac, a[0][0].rb, a[1][0].rb, &dist);
        if (!pes.failed[d] && dist >= pes.low[d] && dist <= pes.high[d]) extra_flag |= 2;
    }
    for (int i = 0; i < 2; ++i) {                             // mem_reg2sam (src/bwamem.cpp:1534-1566)
        const int ef = (i == 0 ? 0x41 : 0x81) | extra_flag;
        int l = 0;
        for (int k = 0; k < n[i]; ++k) {
            const bm2_alnreg_t *q = &a[i][k];
            if (q->score < p.T) continue;
            if (q->secondary >= 0 && (sam_is_alt_d(*q) || !(p.flag & 0x8))) continue;
            if (q->secondary >= 0 && q->secondary < 0x7fffffff && q->score < a[i][q->secondary].score * p.drop_ratio) continue;
            if (n_aa[i] >= sc.aa_cap) { *overflow |= BM2_OVF_RECORDS; break; }
            SamAln &t = sc.aa[i][n_aa[i]];
            if (!sam_alloc_d(pl, l_seq[i], q, &t)) { *overflow |= BM2_OVF_POOL; return; }
            sam_reg2aln_d(p, tb, cv, ref, l_seq[i], seq[i], q, sc.he, sc.zz, &t, overflow);
            t.flag |= ef; t.reg = k;
            if (q->secondary >= 0) t.sub = -1;
            if (l && q->secondary < 0) t.flag |= (p.flag & 0x10) ? 0x10000 : 0x800;
            if (!(p.flag & 0x1000) && l && !sam_is_alt_d(*q) && t.mapq > sc.aa[i][0].mapq) t.mapq = sc.aa[i][0].mapq;
            ++n_aa[i]; ++l;
        }
        if (n_aa[i] == 0) {
            SamAln &t = sc.aa[i][0];
            if (!sam_alloc_d(pl, l_seq[i], 0, &t)) { *overflow |= BM2_OVF_POOL; return; }
            sam_reg2aln_d(p, tb, cv, ref, l_seq[i], seq[i], 0, sc.he, sc.zz, &t, overflow);
            t.flag |= ef;
            n_aa[i] = 1;
        }
    }
    out_all();
}

// The single-end branch of worker_sam for one read (src/bwamem.cpp:1320-1334): mem_mark_primary_se, -5 reordering, then mem_reg2sam
// (:1521-1577) without a mate.  a[0..n) the read's regions (modified: marking), id = the read's index in the run.
// sc: z / idx n + 4 ints each, he, zz, aa[0] (aa_cap >= n + 1 records), pools, ops as for a pair.
template <class Emit, class EmitXa>
BM2_HD void sam_se_read_d(const SamParams &p, const SamTables &tb, const ContigView &cv, const uint8_t *ref, const uint8_t *seq, int l_seq, bm2_alnreg_t *a, int n,
                          int64_t id, const SamScratch &sc, Emit &emit, EmitXa &emit_xa, int *overflow)
{
    SamPool pl = { sc.cig_pool, sc.cig_cap, 0, sc.md_pool, sc.md_cap, 0 };
    sam_mark_primary_se_d(p, n, a, id, sc.z, sc.idx);
    if (p.flag & 0x800) sam_reorder_primary5_d(p.T, n, a);
    if (!(p.flag & 0x8) && n > 0) {                               // XA entries (mem_gen_alt inside mem_reg2sam, :1529-1530)
        bm2_alnreg_t widest; widest.rb = 0; widest.re = 0;
        for (int k = 0; k < n; ++k) if (a[k].re - a[k].rb > widest.re) widest.re = a[k].re - a[k].rb;
        SamAln t;
        if (!sam_alloc_d(pl, l_seq, &widest, &t)) { *overflow |= BM2_OVF_POOL; return; }
        auto one = [&](int r, const SamAln &e) { emit_xa(0, r, e); };
        sam_gen_alt_d(p, tb, cv, ref, l_seq, seq, a, n, sc.z, sc.idx, sc.he, sc.zz, t, one, overflow);
    }
    int n_aa = 0, l = 0;
    SamAln *aa = sc.aa[0];
    for (int k = 0; k < n; ++k) {
        const bm2_alnreg_t *q = &a[k];
        if (q->score < p.T) continue;
        if (q->secondary >= 0 && (sam_is_alt_d(*q) || !(p.flag & 0x8))) continue;
        if (q->secondary >= 0 && q->secondary < 0x7fffffff && q->score < a[q->secondary].score * p.drop_ratio) continue;
        if (n_aa >= sc.aa_cap) { *overflow |= BM2_OVF_RECORDS; break; }
        SamAln &t = aa[n_aa];
        if (!sam_alloc_d(pl, l_seq, q, &t)) { *overflow |= BM2_OVF_POOL; return; }
        sam_reg2aln_d(p, tb, cv, ref, l_seq, seq, q, sc.he, sc.zz, &t, overflow);
        t.reg = k;
        if (q->secondary >= 0) t.sub = -1;
        if (l && q->secondary < 0) t.flag |= (p.flag & 0x10) ? 0x10000 : 0x800;
        if (!(p.flag & 0x1000) && l && !sam_is_alt_d(*q) && t.mapq > aa[0].mapq) t.mapq = aa[0].mapq;
        ++n_aa; ++l;
    }
    if (n_aa == 0) {
        SamAln &t = aa[0];
        if (!sam_alloc_d(pl, l_seq, 0, &t)) { *overflow |= BM2_OVF_POOL; return; }
        sam_reg2aln_d(p, tb, cv, ref, l_seq, seq, 0, sc.he, sc.zz, &t, overflow);
        n_aa = 1;
    }
    for (int k = 0; k < n_aa; ++k) {
        SamRec r;
        sam_aln2rec_d(p, aa[k], k, (const SamAln *) 0, &r, sc.ops);
        emit(0, k, r, sc.ops, aa[k].n_cigar ? aa[k].md : "");
    }
}

