// ksw_emul.cpp — TEST-ONLY host build of the mate-rescue local alignment's device logic (bwa-mem2_b200/csrc/ksw_device.cuh) so
// that its one-sweep formulation can be checked against the oracle's lane-by-lane restatement of the reference's striped kernels
// (which is pinned to the reference binary).  Never part of the product.
#include <vector>
#include <cstdint>
#include "ksw_device.cuh"

extern "C" int emul_ksw_align2(int32_t qlen, const uint8_t *query, int32_t tlen, const uint8_t *target, const int8_t *mat, int32_t o_del,
                               int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out)
{
    std::vector<int32_t> scratch((size_t) 3 * (qlen + 16), 0x5A5A5A5A), bsc((size_t) tlen / 2 + 2), bpos((size_t) tlen / 2 + 2);      // rows of the window, neighbours merged
    std::vector<uint8_t> tmp((size_t) tlen + 1, 0xEE);
    int overflow = 0;
    KswRes r = ksw_align2_d(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, scratch.data(), bsc.data(), bpos.data(), (int) bsc.size(), tmp.data(), &overflow);
    out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
    return overflow;
}
