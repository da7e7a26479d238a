// SMEM access-locality study (CPU only): replays the index accesses of the three SMEM passes (the kernels' own device logic, fm_device.cuh)
// through an LRU cache scaled to the index size and counts the 64-byte lines that would come from DRAM, for the current layout and for
// alternatives.  Built and run by scripts/study_smem_locality.py.
#include <vector>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <algorithm>
struct Acc { int64_t p1, p2, s; int phase; };
static std::vector<Acc> g_acc; static int g_phase = 0;
#define BM2_TRACE_EXT(p1, p2, s) g_acc.push_back(Acc{ (p1), (p2), (s), g_phase })
#include "fm_device.cuh"

struct Lru {             // set-associative LRU over 64-byte line ids
    int ways; size_t sets; std::vector<int64_t> tag; std::vector<uint32_t> age; uint32_t clock = 0;
    Lru(size_t lines, int w) : ways(w), sets(std::max<size_t>(1, lines / w)), tag(sets * w, -1), age(sets * w, 0) {}
    bool access(int64_t line) {
        const size_t s = (size_t) ((uint64_t) line * 0x9E3779B97F4A7C15ULL >> 20) % sets;
        int64_t *t = &tag[s * ways]; uint32_t *a = &age[s * ways];
        ++clock;
        int victim = 0;
        for (int w = 0; w < ways; ++w) { if (t[w] == line) { a[w] = clock; return true; } if (a[w] < a[victim]) victim = w; }
        t[victim] = line; a[victim] = clock;
        return false;
    }
};

extern "C" void smem_study(const bm2_index_desc *idx, const uint8_t *codes, const int64_t *offs, int n_reads, long long cache_lines, int kmer_k,
                           double *out /* [variant 0..3][phase 0..4][ext, lines, misses] */)
{
    FmIndexView fm; fm.cp_occ = idx->cp_occ; fm.sa_ms = idx->sa_ms_byte; fm.sa_ls = idx->sa_ls_word; fm.sentinel = idx->sentinel_index;
    for (int i = 0; i < 5; ++i) fm.count[i] = idx->count[i];
    SmemParams sp; sp.min_seed_len = 19; sp.split_len = 28; sp.split_width = 10; sp.max_mem_intv = 20;
    Lru c0((size_t) cache_lines, 16), c1((size_t) cache_lines, 16), c2((size_t) cache_lines, 16), c3((size_t) cache_lines, 16);
    std::vector<FmPrev> scratch(1024);
    memset(out, 0, sizeof(double) * 4 * 5 * 3);
    auto O = [&](int v, int ph, int k) -> double & { return out[(v * 5 + ph) * 3 + k]; };
    std::vector<std::vector<Acc>> blk;
    const int block = 8192;                                  // reads in flight in the replay
    for (int r = 0; r < n_reads; ++r) {
        const uint8_t *q = codes + offs[r]; const int len = (int) (offs[r + 1] - offs[r]);
        QPlain qq = { q };
        g_acc.clear();
        unsigned ne = 0;
        struct S { int x, mi; std::vector<FmPrev> l; };
        std::vector<S> tasks, tasks2; std::vector<std::pair<int, int>> reseeds; bool pass1 = true;
        auto sink = [&](int x, int mi, const FmPrev *list, int nl) { S t; t.x = x; t.mi = mi; t.l.assign(list, list + nl); tasks.push_back(t); };
        auto sink2 = [&](int x, int mi, const FmPrev *list, int nl) { S t; t.x = x; t.mi = mi; t.l.assign(list, list + nl); tasks2.push_back(t); };
        auto emit = [&](int m, int n, int64_t, int64_t, int64_t ss) { if (pass1 && n + 1 - m >= sp.split_len && ss <= sp.split_width) reseeds.push_back({ (n + 1 + m) >> 1, (int) (ss + 1) }); };
        g_phase = 0; fm_forward(fm, qq, len, 0, 1, false, scratch.data(), sink, ne);
        g_phase = 1; for (auto &t : tasks) fm_backward_rows(fm, qq, t.x, t.mi, sp.min_seed_len, t.l.data(), (int) t.l.size(), emit, ne);
        pass1 = false;
        g_phase = 2; for (auto &rs : reseeds) fm_forward(fm, qq, len, rs.first, rs.second, true, scratch.data(), sink2, ne);
        g_phase = 3; for (auto &t : tasks2) fm_backward_rows(fm, qq, t.x, t.mi, sp.min_seed_len, t.l.data(), (int) t.l.size(), emit, ne);
        g_phase = 4; fm_smem_pass3(fm, qq, len, sp, emit, ne);
        blk.push_back(g_acc);
        if ((int) blk.size() < block && r + 1 < n_reads) continue;
        // replay the block: the reads advance in lock step, one access each per round (as the threads of the GPU kernels do)
        const size_t nb = blk.size();
        std::vector<size_t> cur(nb, 0); std::vector<int> stepv(nb, 0), uniq(nb, 0), pph(nb, -1); std::vector<int64_t> ps(nb, -1);
        bool any = true;
        while (any) { any = false;
        for (size_t bi = 0; bi < nb; ++bi) {
            if (cur[bi] >= blk[bi].size()) continue;
            any = true;
            const Acc &a = blk[bi][cur[bi]++];
            int &step = stepv[bi]; int64_t &prev_s = ps[bi]; int &prev_phase = pph[bi]; bool in_unique = uniq[bi] != 0;
            const bool fwd = a.phase == 0 || a.phase == 2 || a.phase == 4;
            if (a.phase != prev_phase || (fwd && a.s > prev_s)) { step = 0; in_unique = false; }      // a new forward search: the interval grows only there
            prev_phase = a.phase; prev_s = a.s;
            const int64_t l1 = a.p1 >> 6, l2 = a.p2 >> 6;
            // variant 0: the current layout (one 64-byte checkpoint per 64 rows)
            O(0, a.phase, 0) += 1; O(0, a.phase, 1) += l1 == l2 ? 1 : 2;
            O(0, a.phase, 2) += !c0.access(l1); if (l2 != l1) O(0, a.phase, 2) += !c0.access(l2);
            // variant 1: one 64-byte line per 128 rows (2-bit packed BWT + counts): half the table
            { const int64_t h1 = a.p1 >> 7, h2 = a.p2 >> 7; O(1, a.phase, 0) += 1; O(1, a.phase, 1) += h1 == h2 ? 1 : 2;
              O(1, a.phase, 2) += !c1.access(h1); if (h2 != h1) O(1, a.phase, 2) += !c1.access(h2); }
            // variant 2: a k-mer table answers the first kmer_k - 1 extensions of every forward search with one (uncached) fetch
            if (fwd && step < kmer_k - 1) { if (step == 0) { O(2, a.phase, 0) += 1; O(2, a.phase, 1) += 1; O(2, a.phase, 2) += 1; } }
            else { O(2, a.phase, 0) += 1; O(2, a.phase, 1) += l1 == l2 ? 1 : 2; O(2, a.phase, 2) += !c2.access(l1); if (l2 != l1) O(2, a.phase, 2) += !c2.access(l2); }
            // variant 3: pass-1 forward stretches with a unique interval (s == 1) verified against the reference text: 8 accesses per stretch
            if (a.phase == 0 && a.s == 1) { if (!in_unique) { in_unique = true; O(3, a.phase, 0) += 8; O(3, a.phase, 1) += 8; O(3, a.phase, 2) += 8; } }
            else { O(3, a.phase, 0) += 1; O(3, a.phase, 1) += l1 == l2 ? 1 : 2; O(3, a.phase, 2) += !c3.access(l1); if (l2 != l1) O(3, a.phase, 2) += !c3.access(l2); }
            ++step;
            uniq[bi] = in_unique;
        }
        }
        blk.clear();
    }
}
