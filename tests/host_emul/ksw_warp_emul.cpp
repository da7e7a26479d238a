// ksw_warp_emul.cpp — TEST-ONLY host driver of the warp formulation of the mate-rescue local alignment (bwa-mem2_b200/csrc/ksw_warp.cuh):
// the per-lane phases are the product's functions; the 32 lanes are a loop, the warp scan and the reductions plain loops in lane order.
// Checked against the oracle / the reference's golden vectors by tests/test_oracle_ksw.py.  Never part of the product.
#include <vector>
#include <cstdint>
#include "ksw_warp.cuh"

template <int T>
static KswRes pass_lanes_t(int size, int qlen, const uint8_t *query, int qstride, int comp, int tlen, const uint8_t *target, int tstride, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const KswShape s = ksw_shape_d(size, qlen, mat, o_del, e_del, o_ins, e_ins);
    int qmax = 0;
    for (int a = 0; a < 25; ++a) if (mat[a] > qmax) qmax = mat[a];
    const int minsc = (xtra & BM2_KSW_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & BM2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
    std::vector<KswLaneT<T>> L(32);
    for (int l = 0; l < 32; ++l) ksw_lane_init_d(s, l, query, qstride, mat, L[l], comp);
    KswRowState st; st.gmax = 0; st.te = -1; st.n_b = 0; st.last_sc = 0; st.last_pos = -2; st.stop = false;
    for (int i = 0; i < tlen && !st.stop; ++i) {
        const int tbase = target[(long long) i * tstride] > 4 ? 4 : target[(long long) i * tstride];
        int last[32];
        for (int l = 0; l < 32; ++l) last[l] = L[l].hlast;
        KswSummary m[32], in[32];
        for (int l = 0; l < 32; ++l) m[l] = ksw_lane_phase_a_d(s, tbase, l ? last[l - 1] : 0, L[l]);
        KswSummary run; run.v_seg = BM2_KSW_NONE; run.v_full = BM2_KSW_NONE;                  // exclusive prefix maxima of the scan keys, lane order
        for (int l = 0; l < 32; ++l) {
            in[l] = ksw_scan_entering_d(s, L[l], run);
            const KswSummary key = ksw_scan_keys_d(s, L[l], m[l]);
            if (key.v_seg > run.v_seg) run.v_seg = key.v_seg;
            if (key.v_full > run.v_full) run.v_full = key.v_full;
        }
        int rowmax = 0;
        for (int l = 0; l < 32; ++l) { const int r = ksw_lane_phase_b_d(s, in[l], L[l]); if (r > rowmax) rowmax = r; }
        bool took;
        ksw_row_end_d(s, i, rowmax, minsc, endsc, st, bsc, bpos, bcap, overflow, &took, true);
        if (took) for (int l = 0; l < 32; ++l) for (int c = 0; c < L[l].ncol; ++c) L[l].Hbest[c] = L[l].H[c];
    }
    KswRes r; r.score = size == 1 ? (st.gmax + s.shift < 255 ? st.gmax : 255) : st.gmax; r.te = st.te; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    if (size == 2 || r.score != 255) {
        int mx = -1;
        for (int l = 0; l < 32; ++l) for (int c = 0; c < L[l].ncol; ++c) if (L[l].Hbest[c] > mx) { mx = L[l].Hbest[c]; r.qe = L[l].col0 + c; }
        if (st.n_b) {
            const int d = (r.score + qmax - 1) / qmax, low = st.te - d, high = st.te + d;
            for (int k = 0; k < st.n_b; ++k) if ((bpos[k] < low || bpos[k] > high) && bsc[k] > r.score2) { r.score2 = bsc[k]; r.te2 = bpos[k]; }
        }
    }
    return r;
}

// the same choice of the lane capacity as ksw_pass_warp_d
static KswRes pass_lanes(int size, int qlen, const uint8_t *query, int qstride, int comp, int tlen, const uint8_t *target, int tstride, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const int p = size == 1 ? 16 : 8, nlen = (qlen + p - 1) / p * p;
    switch (ksw_lane_width_d((nlen + 31) / 32)) {
    case 5: return pass_lanes_t<5>(size, qlen, query, qstride, comp, tlen, target, tstride, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    case 8: return pass_lanes_t<8>(size, qlen, query, qstride, comp, tlen, target, tstride, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    default: return pass_lanes_t<BM2_KSW_CMAX>(size, qlen, query, qstride, comp, tlen, target, tstride, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    }
}

// the query is query[0], query[qstride], ... (complemented if comp), as ksw_align2_warp_d takes it
extern "C" int emul_ksw_warp_align2_q(int32_t qlen, const uint8_t *query, int32_t qstride, int32_t comp, int32_t tlen, const uint8_t *target, const int8_t *mat,
                                      int32_t o_del, int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out)
{
    if (qlen > 32 * BM2_KSW_CMAX - 15 || !ksw_scan_ok_d(e_ins, qlen)) return -1;
    std::vector<int32_t> bsc((size_t) tlen / 2 + 2), bpos((size_t) tlen / 2 + 2);
    std::vector<uint8_t> tmp((size_t) tlen + 1);
    int overflow = 0;
    const int size = (xtra & BM2_KSW_XBYTE) ? 1 : 2;
    KswRes r = pass_lanes(size, qlen, query, qstride, comp, tlen, target, 1, mat, o_del, e_del, o_ins, e_ins, xtra, bsc.data(), bpos.data(), (int) bsc.size(), &overflow);
    if (!((xtra & BM2_KSW_XSTART) == 0 || ((xtra & BM2_KSW_XSUBO) && r.score < (xtra & 0xffff)))) {          // ksw_align2's second pass (ksw_device.cuh)
        for (int i = 0; i <= r.te; ++i) tmp[i] = target[r.te - i];
        for (int i = r.te + 1; i < tlen; ++i) tmp[i] = target[i];
        int ov2 = 0;
        const KswRes rr = pass_lanes(size, r.qe + 1, query + (long long) r.qe * qstride, -qstride, comp, tlen, tmp.data(), 1, mat, o_del, e_del, o_ins, e_ins, BM2_KSW_XSTOP | r.score, bsc.data(),
                                     bpos.data(), (int) bsc.size(), &ov2);
        if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    }
    out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
    return overflow;
}

extern "C" int emul_ksw_warp_align2(int32_t qlen, const uint8_t *query, int32_t tlen, const uint8_t *target, const int8_t *mat, int32_t o_del,
                                    int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out)
{
    return emul_ksw_warp_align2_q(qlen, query, 1, 0, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, out);
}
