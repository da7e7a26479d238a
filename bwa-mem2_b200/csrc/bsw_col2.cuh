// bsw_col2.cuh — the extension DP of ONE job with TWO ADJACENT COLUMNS per packed 16-bit instruction
// (DPX: VIADDMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2, VIMNMX.U16x2).
//
// Same DP as bsw_extend_one (bsw.cu; reference src/bandedSWA.cpp:116-237 with the SIMD-wrapper band and z-drop,
// :2905-2926) for jobs whose scores fit 8 bits (h0 + min(qlen, tlen) * a <= 255: every job of a 2x151 bp read) and
// whose query is at most 256 columns long.  Columns 2p (low half) and 2p+1 (high half) of a row share every packed
// register.  M, E, and the row maximum are independent per column; only F runs along the row,
//     F(j+1) = max(F(j) - e_ins, M(j) - oe_ins, 0),
// and crosses from the low to the high half inside a pair and from the high half of a pair to the low half of the
// next (shifts written as multiplies: FMA pipe).  Both columns belong to the SAME job, so - unlike two jobs per thread
// (bsw_pair.cuh) - the band, the exit row and the trip count are shared: the lanes of a warp diverge no more than in
// the one-cell-per-instruction kernel.  Odd band edges are single cells (at most two per row).
//
// Per pair of cells: 1 LDS + 1 STS of the state {H_lo, E_lo, H_hi, E_hi} (4 x 8 bit), 1 LDS of the two selector bytes, one PRMT
// for both substitution scores (the per-column selector byte is precomputed from the query), 10 ALU-pipe + 7 FMA-pipe instructions
// (unpacking H, the row-maximum key and all packing are multiply-adds on the FMA pipe).
//
// Written as BM2_HD so that tests/host_emul/bsw_col2_emul.cpp runs the very same code on the CPU against the oracle.
#pragma once
#include "bsw_pair.cuh"          // p2_* packed-halfword primitives (device intrinsics + portable host definitions)

#if defined(__CUDA_ARCH__)
BM2_D uint32_t c2_shr16(uint32_t x) { return __umulhi(x, 65536u); }      // x >> 16 on the FMA pipe
#else
inline uint32_t c2_shr16(uint32_t x) { return x >> 16; }
#endif

// selector byte of one query base (codes > 4 count as 4 = N): PRMT nibbles {qb | 8 (sign of that byte), qb}
BM2_HD uint32_t c2_selector_byte(int qb) { const uint32_t q = qb > 4 ? 4u : (uint32_t) qb; return q * 0x11u + 0x80u; }

// can a job go to the column-pair kernel?  (+ the caller's class rule: 8-bit scores, qlen <= 256)
BM2_HD bool c2_params_ok(const BswParams &p) { return p2_params_ok(p); }

// Mem: ldw/stw (state word of column pair p = columns 2p, 2p+1), ldh/sth (16-bit state {H, E << 8} of one column),
// sel16(p) (the two selector bytes of pair p: columns 2p, 2p+1).
// SAME_OE: o_del + e_del == o_ins + e_ins (the default scoring), one packed add per pair less.
//
// Round 2: ONE control flow for all 32 lanes of a warp.  The row is the pairs pb = beg >> 1 .. pe = (end - 1) >> 1; a pair that
// sticks out of the band on the left (odd beg) or on the right (odd end) is run like any other pair on a masked input word:
//   * dead low half (column beg - 1): H and E read as 0, so M = E' = h = 0 and nothing flows into column beg (F enters as 0, the
//     left neighbour's H is h1 = 0 because beg > 0); its state bytes are written back unchanged;
//   * dead high half (column end): H and E read as 0, its stored state becomes {H(i, end - 1), 0} - exactly the reference's
//     eh[end] = {h1, 0} (src/bandedSWA.cpp:203) - and its key is masked out of the row maximum.
// Every lane therefore executes: masked first pair, the pair loop (trip counts differ, code does not), masked last pair.  (Round 1
// ran odd edges as single 32-bit cells and the pair loop as quads with remainder blocks: five code paths per row that the lanes of
// a warp took at different times - 24-26 of 32 lanes active per instruction.)
// F crosses the halves with one multiply and one PRMT per pair: with U = {F(2q), F(2q)},
//   f2 = max(U + {0, -e}, {0, t(2q)}, 0) = {F(2q), F(2q+1)},  fb = max(f2 + {-e, -e}, {t(2q), t(2q+1)}, 0) -> high half F(2q+2),
// t = M - oe_ins (round 1: two multiplies up, a merge and a shift down per pair).
template <bool SAME_OE, class Mem, int REG_SHRINK = 1, int UNR = 4>
BM2_HD void bsw_col2_extend(const Mem &mem, const uint8_t *tptr, int tstride, int qlen, int tlen, int h0, const BswParams &p,
                            BswOut &o, unsigned long long &cells)
{
    const int oe_del = p.o_del + p.e_del, oe_ins = p.o_ins + p.e_ins, e_del = p.e_del, e_ins = p.e_ins;
    const uint32_t n_oe_del = ((uint32_t) (-oe_del) & 0xFFFFu) * 0x10001u, n_e_del = ((uint32_t) (-e_del) & 0xFFFFu) * 0x10001u;
    const uint32_t n_oe_ins = ((uint32_t) (-oe_ins) & 0xFFFFu) * 0x10001u, n_e_ins = ((uint32_t) (-e_ins) & 0xFFFFu) * 0x10001u;
    const uint32_t n_e_ins_hi = n_e_ins & 0xFFFF0000u;                                       // {0, -e_ins}
    const uint32_t k256 = qlen >= 0 ? 256u : 0u;       // 256, opaque to the compiler: h * k256 + c stays a multiply-add (FMA pipe), not an LEA (ALU pipe)
    // first row (bandedSWA.cpp:141-144): columns 0..qlen, E = 0
    {
        int h = h0;
        mem.sth(0, (uint32_t) h);
        h = h0 > oe_ins ? h0 - oe_ins : 0;
        for (int j = 1; j <= qlen; ++j) {
            mem.sth(j, (uint32_t) h);
            h = h > e_ins ? h - e_ins : 0;
        }
    }
    // band (SIMD wrapper arithmetic, bandedSWA.cpp:2905-2926)
    int w = p.w;
    const BswQuirk qk = bsw_quirk(qlen, tlen, h0, p);
    {
        unsigned t1 = ((unsigned) (qlen * p.a) + (unsigned) (p.end_bonus - p.o_ins)) & qk.band_mask;
        int max_ins = (int) (t1 / (unsigned) e_ins) + 1; if (max_ins < 1) max_ins = 1;
        unsigned t2 = ((unsigned) (qlen * p.a) + (unsigned) (p.end_bonus - p.o_del)) & qk.band_mask;
        int max_del = (int) (t2 / (unsigned) e_del) + 1; if (max_del < 1) max_del = 1;
        if (w > max_ins) w = max_ins;
        if (w > max_del) w = max_del;
    }
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    unsigned long long ncell = 0;
    int tb_next = tlen > 0 ? (int) tptr[0] : 0;           // the target base of a row is fetched one row ahead (global-memory latency)
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1 = 0;
        if (beg == 0) { h1 = h0 - (p.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        const int tb = tb_next;
        if (i + 1 < tlen) tb_next = (int) tptr[(long long) (i + 1) * tstride];
        const uint32_t tbl = p2_score_table(tb, p.a, p.b);
        // Row maximum with its column, per half: key = H << 8 | c, c = 0xFF at the pair that set it and one less for every pair after it (ONE
        // packed add-max per pair: the running key minus one against the pair's {H << 8 | 0xFF}; an equal H further right wins, as in the
        // reference's row scan).  The column is pe - (0xFF - c) afterwards.  c never wraps: a row has at most 129 pairs.
        uint32_t key2 = 0x00FF00FFu;
        uint32_t fw = 0, lw = 0;            // the state words just written for the first / last pair of the row (band shrink below)
        if (end > beg) {
            const int pb = beg >> 1, pe = (end - 1) >> 1;
            const uint32_t in_first = (beg & 1) ? 0xFFFF0000u : 0xFFFFFFFFu, in_last = (end & 1) ? 0x0000FFFFu : 0xFFFFFFFFu;
            uint32_t U = 0;                                            // {F(2q), F(2q)}
            uint32_t hp = (uint32_t) h1 << 16;                         // packed h of the pair to the left: its high half is H(i, 2q - 1)
            // the DP of one pair on the (masked) state word wv; returns the new state word
            auto pair = [&](const uint32_t wv, const uint32_t sel, const uint32_t keymask) -> uint32_t {
                const uint32_t e = p2_prmt(wv, 0u, 0x4341u);
                const uint32_t hd = p2_mad(e, 0xFFFFFF00u, wv);                          // wv - (e << 8) on the FMA pipe
                const uint32_t s = p2_prmt(tbl, 0xFFFFFFFFu, sel);                       // low 16 selector bits: columns 2q, 2q+1
                const uint32_t M = p2_addmin_relu(hd, s, p2_mad(hd, 128u, 0u));          // hd ? max(hd + s, 0) : 0   (s <= 127)
                const uint32_t td = p2_add(M, n_oe_del);
                const uint32_t en = p2_addmax_relu(e, n_e_del, td);
                const uint32_t ti = SAME_OE ? td : p2_add(M, n_oe_ins);
                const uint32_t f2 = p2_addmax_relu(U, n_e_ins_hi, p2_mad(ti, 65536u, 0u));   // {F(2q), F(2q+1)}
                const uint32_t fb = p2_addmax_relu(f2, n_e_ins, ti);                     // high half: F(2q+2)
                U = p2_prmt(fb, fb, 0x3232u);
                const uint32_t h = p2_max3(M, e, f2);
                const uint32_t hs = p2_mad(h, 65536u, c2_shr16(hp));                     // {H(i, 2q-1), H(i, 2q)}
                hp = h;
                key2 = p2_addmaxu(key2, 0xFFFFFFFFu, p2_mad(h, k256, 0x00FF00FFu) & keymask);   // {H << 8 | 0xFF - pairs since}: see below
                return p2_mad(en, 256u, hs);
            };
            // The state word and the selectors of a pair are loaded one pair AHEAD of their use (the accesses are volatile asm, so the
            // compiler keeps this order): the shared-memory latency of pair q + 1 runs under the arithmetic of pair q.
            const bool only = pb == pe;
            const uint32_t old = mem.ldw(pb), sel0 = mem.sel16(pb);
            uint32_t wv = 0, sl = 0;
            if (!only) { wv = mem.ldw(pb + 1); sl = mem.sel16(pb + 1); }
            {   // first pair (also the last one when the band is a single pair)
                const uint32_t in_and = only ? (in_first & in_last) : in_first;
                const uint32_t nw = pair(old & in_and, sel0, only ? in_last : 0xFFFFFFFFu);
                fw = lw = (nw & in_first) | (old & ~in_first);
                mem.stw(pb, fw);
            }
            if (!only) {
#if defined(__CUDA_ARCH__)
#pragma unroll UNR
#endif
                for (int q = pb + 1; q < pe; ++q) {                       // (unrolled x UNR, remainder first)
                    const uint32_t wn = mem.ldw(q + 1), sn = mem.sel16(q + 1);
                    mem.stw(q, pair(wv, sl, 0xFFFFFFFFu));
                    wv = wn; sl = sn;
                }
                lw = pair(wv & in_last, sl, in_last);
                mem.stw(pe, lw);
            }
            h1 = (int) ((end & 1) ? (hp & 0xFFFFu) : (hp >> 16));      // H(i, end - 1)
            ncell += (unsigned) (end - beg);
        }
        int m = 0, mj = 0;
        if (end > beg) {
            const int pe = (end - 1) >> 1;
            const int ka = (int) (key2 & 0xFFFFu), kb = (int) (key2 >> 16);
            const int ja = 2 * (pe - (0xFF - (ka & 0xFF))), jb = 2 * (pe - (0xFF - (kb & 0xFF))) + 1;
            const int ha = ka >> 8, hb = kb >> 8;
            const bool hi = hb > ha || (hb == ha && jb > ja);
            m = hi ? hb : ha; mj = hi ? jb : ja;
        }
#ifdef BM2_COL2_TRACE
        BM2_COL2_TRACE(beg, end);                              // lane-utilisation studies (tests/host_emul/bsw_col2_emul.cpp)
#endif
        if (!(end > beg && (end & 1))) mem.sth(end, (uint32_t) h1);     // (odd end: the masked last pair has written {h1, 0} there)
        if ((end > beg ? end : beg) == qlen) {
            if (h1 >= gscore) best_ie = i;
            if (h1 > gscore) gscore = h1;
        }
        if (m == 0) break;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            int d = mj - i; d = d < 0 ? -d : d;
            if (d > max_off) max_off = d;
            if (0 > qk.zthr) break;
        } else {
            const int di = i - best_i, dj = mj - best_j;
            const int pen = di > dj ? di - dj : dj - di;       // SIMD z-drop: no e_del/e_ins factor, no `zdrop > 0` guard (ZSCORE8/16)
            if (best - m - pen > qk.zthr) break;
        }
        // Band shrink (src/bandedSWA.cpp:222-225): beg = first column of [beg, end) with a non-zero state, end = last one + 2.  The two
        // columns at either edge are decided from the words just written (registers); shared memory is only scanned when both are zero
        // (round 2: the scan loops - a load, a compare and a branch per step, 19 of 32 lanes - took 11 % of the kernel's samples).
        int j = beg;                                           // (here end > beg: an empty row has m == 0 and left the loop above)
        if (REG_SHRINK == 2) {                                 // the column at either edge from the words just written, then the scans
            const uint32_t f16 = (beg & 1) ? fw >> 16 : fw & 0xFFFFu;
            if (f16 == 0u) for (++j; j < end && mem.ldh(j) == 0u; ++j) {}
            beg = j;
            j = end;
            if (h1 == 0) {
                const uint32_t l16 = (end & 1) ? lw & 0xFFFFu : lw >> 16;
                --j;
                if (j >= beg && l16 == 0u) for (--j; j >= beg && mem.ldh(j) == 0u; --j) {}
            }
            end = j + 2 < qlen ? j + 2 : qlen;
            continue;
        }
        if (!REG_SHRINK) {                                     // round 1's scans (A/B measurements: BM2_BSW_REGSHRINK=0)
            for (; j < end && mem.ldh(j) == 0u; ++j) {}
            beg = j;
            j = end;
            if (h1 == 0) for (--j; j >= beg && mem.ldh(j) == 0u; --j) {}
            end = j + 2 < qlen ? j + 2 : qlen;
            continue;
        }
        {
            bool scan = false;
            if (!(beg & 1)) {
                if ((fw & 0xFFFFu) == 0u) { j = beg + 1; if (j < end && (fw >> 16) == 0u) { j = beg + 2; scan = true; } }
            } else if ((fw >> 16) == 0u) { j = beg + 1; scan = true; }
            if (scan) for (; j < end && mem.ldh(j) == 0u; ++j) {}
        }
        beg = j;
        j = end;                                               // column `end` holds {h1, 0}: nothing to look at in the usual case
        if (h1 == 0) {
            bool scan = false;
            --j;
            if (end & 1) {                                     // column end - 1 is the low half of the last word
                if (j >= beg && (lw & 0xFFFFu) == 0u) { --j; scan = true; }
            } else if (j >= beg && (lw >> 16) == 0u) {         // column end - 1 is its high half, end - 2 its low half
                --j;
                if (j >= beg && (lw & 0xFFFFu) == 0u) { --j; scan = true; }
            }
            if (scan) for (; j >= beg && mem.ldh(j) == 0u; --j) {}
        }
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    o.score = best; o.qle = best_j + 1; o.tle = best_i + 1; o.gtle = best_ie + 1; o.gscore = gscore; o.max_off = max_off;
    cells += ncell;
}
