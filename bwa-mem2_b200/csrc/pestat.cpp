// pestat.cpp — insert-size statistics of one chunk of read pairs (host code, no device work: one pass over the best alignment
// region of every read, as in the reference, where this is a serial step between the two worker phases).
//
// Replaces mem_pestat (reference src/bwamem_pair.cpp:81-148; cal_sub :67-79, mem_infer_dir :57-65) as mem_process_seqs calls it once
// per chunk (src/bwamem.cpp:1368-1378).  Its result feeds mate rescue and pairing.  The moments are summed over
// the sorted insert sizes, as the reference does, so avg / std are the same doubles.
#include "bm2_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
// score of the best region overlapping the top one on the query by mask_level, else the score of a bare seed
int second_best(const bm2_mem_opt_t *opt, const bm2_alnreg_t *a, int64_t n) {
    const int top_len = a[0].qe - a[0].qb;
    for (int64_t j = 1; j < n; ++j) {
        const int beg = std::max(a[j].qb, a[0].qb), end = std::min(a[j].qe, a[0].qe);
        if (end <= beg) continue;
        const int shorter = std::min(a[j].qe - a[j].qb, top_len);
        if (end - beg >= shorter * opt->mask_level) return a[j].score;
    }
    return opt->min_seed_len * opt->a;
}
}  // namespace

extern "C" int bm2_pestat(const bm2_mem_opt_t *opt, int64_t l_pac, int32_t n_reads, const bm2_alnreg_t *regs, const int64_t *read_off, bm2_pestat_t pes[4])
{
    if (!opt || !pes || n_reads < 0 || (n_reads > 0 && (!read_off || (!regs && read_off[n_reads] > 0)))) return 1;
    const double min_ratio = 0.8, outlier_bound = 2.0, mapping_bound = 3.0, max_stddev = 4.0, min_dir_ratio = 0.05;      // src/bwamem_pair.cpp:47-53
    const size_t min_dir_cnt = 10;
    memset(pes, 0, 4 * sizeof(bm2_pestat_t));
    std::vector<uint64_t> sizes[4];
    for (int32_t pr = 0; pr + 1 < n_reads; pr += 2) {
        const int64_t n0 = read_off[pr + 1] - read_off[pr], n1 = read_off[pr + 2] - read_off[pr + 1];
        if (n0 == 0 || n1 == 0) continue;
        const bm2_alnreg_t *r0 = regs + read_off[pr], *r1 = regs + read_off[pr + 1];
        if (second_best(opt, r0, n0) > min_ratio * r0->score || second_best(opt, r1, n1) > min_ratio * r1->score) continue;       // not unique enough
        if (r0->rid != r1->rid) continue;
        // orientation and distance of the two starts on the forward-reverse text (mem_infer_dir)
        const bool rev0 = r0->rb >= l_pac, rev1 = r1->rb >= l_pac;
        const int64_t p1 = rev0 == rev1 ? r1->rb : (l_pac << 1) - 1 - r1->rb;
        const int64_t dist = p1 > r0->rb ? p1 - r0->rb : r0->rb - p1;
        const int dir = (rev0 == rev1 ? 0 : 1) ^ (p1 > r0->rb ? 0 : 3);
        if (dist && dist <= opt->max_ins) sizes[dir].push_back((uint64_t) dist);
    }
    size_t most = 0;
    for (int d = 0; d < 4; ++d) {
        std::vector<uint64_t> &q = sizes[d];
        most = std::max(most, q.size());
        if (q.size() < min_dir_cnt) { pes[d].failed = 1; continue; }
        std::sort(q.begin(), q.end());
        const double n = (double) q.size();
        const int p25 = (int) q[(size_t) (int) (.25 * n + .499)], p75 = (int) q[(size_t) (int) (.75 * n + .499)];
        int low = std::max(1, (int) (p25 - outlier_bound * (p75 - p25) + .499));
        int high = (int) (p75 + outlier_bound * (p75 - p25) + .499);
        double sum = 0; size_t cnt = 0;
        for (uint64_t v : q) if (v >= (uint64_t) low && v <= (uint64_t) high) { sum += v; ++cnt; }
        if (cnt == 0) return 2;                                   // the reference asserts (:126)
        const double avg = sum / cnt;
        double ss = 0;
        for (uint64_t v : q) if (v >= (uint64_t) low && v <= (uint64_t) high) ss += (v - avg) * (v - avg);
        const double sd = sqrt(ss / cnt);
        low = (int) (p25 - mapping_bound * (p75 - p25) + .499);
        high = (int) (p75 + mapping_bound * (p75 - p25) + .499);
        if (low > avg - max_stddev * sd) low = (int) (avg - max_stddev * sd + .499);
        if (high < avg + max_stddev * sd) high = (int) (avg + max_stddev * sd + .499);
        pes[d].low = std::max(1, low); pes[d].high = high; pes[d].avg = avg; pes[d].std = sd;
    }
    for (int d = 0; d < 4; ++d)
        if (!pes[d].failed && sizes[d].size() < most * min_dir_ratio) pes[d].failed = 1;
    return 0;
}
