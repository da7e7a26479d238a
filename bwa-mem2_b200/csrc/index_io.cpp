// index_io.cpp — native host loader of the bwa-mem2 on-disk index (input contract of the hot path).
// Formats: <prefix>.bwt.2bit.64 (writer reference src/FMI_search.cpp:154-297, reader :384-460),
// <prefix>.0123 (src/FMI_search.cpp:325-362, read at src/fastmap.cpp:860-881),
// <prefix>.ann (src/bntseq.cpp:73-104, reader :106-177), optional <prefix>.alt (:188-228).
#include "bm2_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>

namespace {
struct OwnedIndex {
    bm2_index_desc d;
    std::vector<void *> allocs;
    std::vector<std::string> names;
};
std::string g_io_error;

bool read_exact(FILE *f, void *dst, size_t bytes) {
    char *p = (char *) dst;
    while (bytes) {
        size_t chunk = bytes > (size_t(1) << 30) ? (size_t(1) << 30) : bytes;
        if (fread(p, 1, chunk, f) != chunk) return false;
        p += chunk; bytes -= chunk;
    }
    return true;
}
}  // namespace

extern "C" const char *bm2_index_io_error(void) { return g_io_error.c_str(); }

extern "C" void bm2_index_free(bm2_index_desc *idx) {
    if (!idx) return;
    OwnedIndex *o = reinterpret_cast<OwnedIndex *>(idx);   // d is the first member
    for (void *p : o->allocs) free(p);
    delete o;
}

extern "C" int bm2_index_load(const char *prefix, bm2_index_desc **out) {
    if (!prefix || !out) return 1;
    *out = nullptr;
    OwnedIndex *o = new OwnedIndex();
    memset(&o->d, 0, sizeof(o->d));
    auto fail = [&](const std::string &m) { g_io_error = m; bm2_index_free(&o->d); return 1; };
    auto alloc = [&](size_t bytes) { void *p = malloc(bytes ? bytes : 1); o->allocs.push_back(p); return p; };

    std::string pre(prefix);
    FILE *f = fopen((pre + ".bwt.2bit.64").c_str(), "rb");
    if (!f) return fail("cannot open " + pre + ".bwt.2bit.64");
    int64_t N = 0, cnt[5];
    if (!read_exact(f, &N, 8) || N <= 0 || !read_exact(f, cnt, 40)) { fclose(f); return fail("bad .bwt.2bit.64 header"); }
    o->d.reference_seq_len = N;
    for (int i = 0; i < 5; ++i) o->d.count[i] = cnt[i] + 1;          // FMI_search.cpp:433-436
    size_t n_occ = (size_t) (N >> 6) + 1, n_sa = (size_t) (N >> 3) + 1;
    void *occ = alloc(n_occ * sizeof(bm2_cp_occ)), *ms = alloc(n_sa), *ls = alloc(n_sa * 4);
    if (!occ || !ms || !ls) { fclose(f); return fail("out of memory loading index"); }
    if (!read_exact(f, occ, n_occ * sizeof(bm2_cp_occ)) || !read_exact(f, ms, n_sa) || !read_exact(f, ls, n_sa * 4) ||
        !read_exact(f, &o->d.sentinel_index, 8)) { fclose(f); return fail("truncated .bwt.2bit.64"); }
    fclose(f);
    o->d.cp_occ = (const bm2_cp_occ *) occ; o->d.sa_ms_byte = (const int8_t *) ms; o->d.sa_ls_word = (const uint32_t *) ls;

    // .ann
    f = fopen((pre + ".ann").c_str(), "r");
    if (!f) return fail("cannot open " + pre + ".ann");
    long long l_pac = 0; int n_seqs = 0; unsigned seed = 0;
    if (fscanf(f, "%lld%d%u", &l_pac, &n_seqs, &seed) != 3 || n_seqs < 0) { fclose(f); return fail("bad .ann header"); }
    o->d.l_pac = l_pac; o->d.n_seqs = n_seqs;
    int64_t *off = (int64_t *) alloc((size_t) n_seqs * 8);
    int32_t *len = (int32_t *) alloc((size_t) n_seqs * 4), *alt = (int32_t *) alloc((size_t) n_seqs * 4);
    std::map<std::string, int> by_name;
    for (int i = 0; i < n_seqs; ++i) {
        unsigned gi; char name[8192];
        if (fscanf(f, "%u%8191s", &gi, name) != 2) { fclose(f); return fail("bad .ann record"); }
        int c;
        while ((c = fgetc(f)) != '\n' && c != EOF) {}                  // rest of line = annotation
        long long o_; int l_, namb;
        if (fscanf(f, "%lld%d%d", &o_, &l_, &namb) != 3) { fclose(f); return fail("bad .ann record"); }
        off[i] = o_; len[i] = l_; alt[i] = 0;
        by_name[name] = i;
    }
    fclose(f);
    o->d.ann_offset = off; o->d.ann_len = len; o->d.ann_is_alt = alt;
    if (2 * (int64_t) l_pac + 1 != N) return fail(".ann l_pac does not match the BWT length");
    // .alt (optional): first column = contig name, lines starting with '@' ignored
    f = fopen((pre + ".alt").c_str(), "r");
    if (f) {
        char line[65536];
        while (fgets(line, sizeof(line), f)) {
            if (line[0] == '@') continue;
            size_t k = strcspn(line, "\t\r\n");
            line[k] = 0;
            auto it = by_name.find(line);
            if (it != by_name.end()) alt[it->second] = 1;
        }
        fclose(f);
    }
    // .0123
    f = fopen((pre + ".0123").c_str(), "rb");
    if (!f) return fail("cannot open " + pre + ".0123");
    void *ref = alloc((size_t) l_pac * 2);
    if (!ref || !read_exact(f, ref, (size_t) l_pac * 2)) { fclose(f); return fail("truncated .0123"); }
    fclose(f);
    o->d.ref_string = (const uint8_t *) ref;
    *out = &o->d;
    return 0;
}
