// sam_layout.cuh — the scratch of ONE read pair in the SAM stage (mate rescue, pairing, MAPQ, CIGAR / NM / MD, records, XA entries):
// capacities that cannot overflow, computed from what the host knows before the launch, and the carving of a pair's arena.
//
// The reference grows every one of these buffers on demand (kvec / realloc in mem_matesw, mem_pair, mem_reg2aln, ksw_align2:
// src/bwamem_pair.cpp:150-346, src/bwamem.cpp:1732-1805, src/ksw.cpp:188-196).  A kernel cannot, so each capacity here is a bound
// that follows from the algorithm, stated next to it.  tests/host_emul/sam_emul.cpp runs the device logic inside arenas carved by this
// header with guard words between the pieces, so a bound that is too small shows up on the host (tests/test_oracle_sam_pe.py,
// scripts/torture.py) before a kernel ever uses it.
#pragma once
#include "sam_device.cuh"

struct SamPairShape {            // known before the launch (seam 2's output and the reads)
    int n[2];                    // regions of read 0 / 1
    int l_seq[2];
    long long max_rlen[2];       // widest region (re - rb) of each read, 0 if none
    long long sum_rlen[2];       // sum of (re - rb) over the regions of each read
    long long max_zcells[2];     // largest backtrack matrix among the regions of each read (sam_reg_zcells_d)
};

// Backtrack cells mem_reg2aln can need for a region: ksw_global2 keeps min(l_query, 2w + 1) columns per reference base, and its band is
// at most max(4 * opt->w, |rlen - l_query| + 3) (the doubling loop of src/bwamem.cpp:1762-1768 stops at opt->w << 2; src/bwa.cpp:292-300).
BM2_HD long long sam_reg_zcells_d(int w_opt, const bm2_alnreg_t &r) {
    const long long lq = r.qe - r.qb, rl = r.re - r.rb;
    if (lq <= 0 || rl <= 0) return 0;
    long long w = 4LL * w_opt, d = rl > lq ? rl - lq : lq - rl;
    if (d + 3 > w) w = d + 3;
    const long long ncol = lq < 2 * w + 1 ? lq : 2 * w + 1;
    return ncol * rl;
}

// shape of read i of a pair from its regions (host side of the launch)
BM2_HD void sam_shape_read_d(SamPairShape &s, int i, int l_seq, const bm2_alnreg_t *a, int n, int w_opt) {
    s.n[i] = n; s.l_seq[i] = l_seq; s.max_rlen[i] = 0; s.sum_rlen[i] = 0; s.max_zcells[i] = 0;
    for (int k = 0; k < n; ++k) {
        const long long rl = a[k].re - a[k].rb, zc = sam_reg_zcells_d(w_opt, a[k]);
        s.sum_rlen[i] += rl;
        if (rl > s.max_rlen[i]) s.max_rlen[i] = rl;
        if (zc > s.max_zcells[i]) s.max_zcells[i] = zc;
    }
}

struct SamPairCaps {
    int acap[2];                 // regions of read i after rescue: every mem_matesw call adds at most one per orientation with statistics (:233-238),
                                 //   and at most min(n[!i], max_matesw) anchors of the other read call it (:398-407)
    int bcap[2];                 // anchor copies: n[i] + 1
    int max_l;                   // longer read
    int tcap;                    // longest rescue window: high - low + l_ms (mate_window_max_d)
    int kcap;                    // score-2 candidates of one local alignment: rows of the window, neighbours merged
    long long rlen_cap[2];       // widest region of read i at SAM time: an original one or a rescued one (inside its window)
    int zi;                      // z / idx / keys entries: max(acap) + 4
    int nv;                      // pairing keys: acap[0] + acap[1] + 4
    int aa_cap;                  // records of one read: its regions + 2 (mem_reg2sam prints each region at most once, + the ALT / unmapped one)
    long long zz_cells;          // backtrack matrix of one global alignment (sam_reg_zcells_d; rescued regions: l_query x window)
    long long pool_ops, pool_md; // CIGAR / MD storage of all records of the pair + one XA entry at a time (sam_alloc_d's sizes)
    int ops_cap;                 // printed CIGAR of one record + its MC tag
    int recs_cap, xa_cap;        // output: records of the pair, XA entries of the pair (one per region at most)
    long long out_ops, out_md;   // output: printed operations of records and XA entries, MD bytes
    size_t scratch_bytes;        // arena of the pair (pieces aligned to 16 bytes), outputs not included
};

BM2_HD size_t sam_align16_d(size_t v) { return (v + 15) & ~(size_t) 15; }

// a / e_del: match score and deletion extension penalty (mem_opt_t): a local alignment of positive score over l query bases spans fewer than
// l + l * a / e_del reference bases (every deleted base costs at least e_del, the matches earn at most l * a).
BM2_HD SamPairCaps sam_pair_caps_d(const SamPairShape &s, const MatePes &pes, int max_matesw, bool rescue, int a, int e_del)
{
    SamPairCaps c;
    c.max_l = s.l_seq[0] > s.l_seq[1] ? s.l_seq[0] : s.l_seq[1];
    c.tcap = 16;
    long long resc[2] = { 0, 0 }, rlen_resc[2] = { 0, 0 };               // rescued regions read i can receive, and how wide one can be
    int n_dirs = 0;
    for (int r = 0; r < 4; ++r) n_dirs += pes.failed[r] ? 0 : 1;          // a call aligns at most one window per orientation that has statistics
    for (int i = 0; i < 2; ++i) {
        const int calls = !rescue ? 0 : (s.n[!i] < max_matesw ? s.n[!i] : max_matesw);
        resc[i] = (long long) n_dirs * calls;
        c.acap[i] = s.n[i] + (int) resc[i] + 4;
        c.bcap[i] = s.n[i] + 1;
        const int win = mate_window_max_d(pes, s.l_seq[i]) + 16;          // read i is the mate that gets aligned into the window
        c.rlen_cap[i] = s.max_rlen[i];
        if (resc[i] > 0) {
            if (win > c.tcap) c.tcap = win;
            const long long span = (long long) s.l_seq[i] + (long long) s.l_seq[i] * a / (e_del > 0 ? e_del : 1) + 2;
            rlen_resc[i] = span < win ? span : win;
            if (rlen_resc[i] > c.rlen_cap[i]) c.rlen_cap[i] = rlen_resc[i];
        }
    }
    c.kcap = c.tcap / 2 + 2;
    c.zi = (c.acap[0] > c.acap[1] ? c.acap[0] : c.acap[1]) + 4;
    c.nv = c.acap[0] + c.acap[1] + 4;
    c.aa_cap = c.zi;
    c.zz_cells = 16; c.pool_ops = 0; c.pool_md = 0; c.ops_cap = 16; c.out_ops = 0; c.out_md = 0;
    for (int i = 0; i < 2; ++i) {
        const long long lq = s.l_seq[i], widest = c.rlen_cap[i];
        if (s.max_zcells[i] + 16 > c.zz_cells) c.zz_cells = s.max_zcells[i] + 16;
        if (resc[i] > 0 && lq * rlen_resc[i] + 16 > c.zz_cells) c.zz_cells = lq * rlen_resc[i] + 16;      // a rescued region: at most l_query columns
        // every region printed once (its own rlen), rescued ones bounded by the window, + h[i], the ALT record, the unmapped record, one XA entry
        const long long extra = 4;
        const long long ops = s.n[i] * (lq + 4) + s.sum_rlen[i] + resc[i] * (lq + rlen_resc[i] + 4) + extra * (lq + widest + 4);
        const long long md = s.n[i] * (2 * lq + 16) + 7 * s.sum_rlen[i] + resc[i] * (2 * lq + 7 * rlen_resc[i] + 16) + extra * (2 * lq + 7 * widest + 16);
        c.pool_ops += ops; c.pool_md += md;
        // output: records as above; XA entries: each region at most once more
        c.out_ops += 2 * ops; c.out_md += md;
        c.ops_cap += (int) (lq + widest + 4);                                 // a record's own CIGAR + the mate's (MC tag)
    }
    // output: every record of read i also carries the MC operations (the CIGAR of the mate's record)
    for (int i = 0; i < 2; ++i) c.out_ops += (long long) (c.acap[i] + 4) * (s.l_seq[!i] + c.rlen_cap[!i] + 4);
    c.recs_cap = c.acap[0] + c.acap[1] + 4;
    c.xa_cap = c.acap[0] + c.acap[1];
    size_t b = 0;
    b += sam_align16_d(sizeof(bm2_alnreg_t) * (size_t) (c.acap[0] + c.acap[1] + c.bcap[0] + c.bcap[1]) + 64);
    b += sam_align16_d((size_t) c.max_l + 1) + sam_align16_d((size_t) c.tcap);                          // rev, tmp
    b += sam_align16_d(4 * (size_t) (3 * (c.max_l + 16))) + 2 * sam_align16_d(4 * (size_t) c.kcap);      // ksw, bsc, bpos
    b += 2 * sam_align16_d(4 * (size_t) c.zi) + sam_align16_d(sizeof(TailSortKey) * (size_t) c.zi);     // z, idx, keys
    b += sam_align16_d(sizeof(SamP64) * (size_t) c.nv) + sam_align16_d(4 * (size_t) (2 * (c.max_l + 2)));   // v, he
    b += sam_align16_d((size_t) c.zz_cells);
    b += 2 * sam_align16_d(sizeof(SamAln) * (size_t) c.aa_cap);
    b += sam_align16_d(4 * (size_t) c.pool_ops) + sam_align16_d((size_t) c.pool_md) + sam_align16_d(4 * (size_t) c.ops_cap);
    c.scratch_bytes = b + 16 * 24;                                       // room for the guard words of the host test build
    return c;
}

// Carves the arena.  guard != 0 (host test build): a 16-byte guard pattern is written after every piece and can be checked with
// sam_arena_guards_ok_d.  The kernel passes guard = 0.
struct SamArena {
    bm2_alnreg_t *a[2], *b[2];
    MateScratch ms;
    SamScratch sc;
    uint8_t *guards[24]; int n_guards;
};

BM2_HD void sam_arena_carve_d(uint8_t *base, const SamPairCaps &c, int guard, SamArena *ar)
{
    size_t off = 0;
    ar->n_guards = 0;
    auto take = [&](size_t bytes) -> uint8_t * {
        uint8_t *p = base + off;
        off += sam_align16_d(bytes);
        if (guard) { uint8_t *g = base + off; for (int k = 0; k < 16; ++k) g[k] = (uint8_t) (0xA5 ^ k); ar->guards[ar->n_guards++] = g; off += 16; }
        return p;
    };
    uint8_t *regs = take(sizeof(bm2_alnreg_t) * (size_t) (c.acap[0] + c.acap[1] + c.bcap[0] + c.bcap[1]) + 64);
    ar->a[0] = (bm2_alnreg_t *) regs; ar->a[1] = ar->a[0] + c.acap[0]; ar->b[0] = ar->a[1] + c.acap[1]; ar->b[1] = ar->b[0] + c.bcap[0];
    ar->ms.rev = take((size_t) c.max_l + 1);
    ar->ms.tmp = take((size_t) c.tcap); ar->ms.tcap = c.tcap;
    ar->ms.ksw = (int32_t *) take(4 * (size_t) (3 * (c.max_l + 16)));
    ar->ms.bsc = (int32_t *) take(4 * (size_t) c.kcap); ar->ms.bpos = (int32_t *) take(4 * (size_t) c.kcap); ar->ms.bcap = c.kcap;
    ar->sc.z = (int32_t *) take(4 * (size_t) c.zi);
    ar->sc.idx = (int32_t *) take(4 * (size_t) c.zi); ar->ms.idx = ar->sc.idx;          // used one after the other
    ar->ms.keys = (TailSortKey *) take(sizeof(TailSortKey) * (size_t) c.zi);
    ar->sc.v = (SamP64 *) take(sizeof(SamP64) * (size_t) c.nv);
    ar->sc.he = (int32_t *) take(4 * (size_t) (2 * (c.max_l + 2)));
    ar->sc.zz.base = take((size_t) c.zz_cells); ar->sc.zz.stride = 1;
    ar->sc.aa[0] = (SamAln *) take(sizeof(SamAln) * (size_t) c.aa_cap);
    ar->sc.aa[1] = (SamAln *) take(sizeof(SamAln) * (size_t) c.aa_cap); ar->sc.aa_cap = c.aa_cap;
    ar->sc.cig_pool = (uint32_t *) take(4 * (size_t) c.pool_ops); ar->sc.cig_cap = c.pool_ops;
    ar->sc.md_pool = (char *) take((size_t) c.pool_md); ar->sc.md_cap = c.pool_md;
    ar->sc.ops = (uint32_t *) take(4 * (size_t) c.ops_cap);
}

BM2_HD bool sam_arena_guards_ok_d(const SamArena &ar) {
    for (int g = 0; g < ar.n_guards; ++g) for (int k = 0; k < 16; ++k) if (ar.guards[g][k] != (uint8_t) (0xA5 ^ k)) return false;
    return true;
}
