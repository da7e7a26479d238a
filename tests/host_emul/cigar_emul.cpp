// cigar_emul.cpp — TEST-ONLY host build of the seam-3 device logic (bwa-mem2_b200/csrc/cigar_device.cuh, what cigar_kernel runs
// per thread) with the same worst-case output stripes and compaction as cigar.cu, checked against the oracle / the reference's
// golden vectors on a machine without a GPU.  Never part of the product.
#include <vector>
#include <cstdlib>
#include <cstring>
#include <climits>
#include "cigar_device.cuh"

extern "C" int emul_gen_cigar(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_cigar_req *reqs,
                              int64_t n, bm2_cigar_rec **recs_out, uint32_t **cigar_out, int64_t *n_ops_out, char **md_out, int64_t *n_md_out)
{
    CigarParams p; memcpy(p.mat, opt->mat, 25); p.o_del = opt->o_del; p.e_del = opt->e_del; p.o_ins = opt->o_ins; p.e_ins = opt->e_ins;
    std::vector<bm2_cigar_rec> recs((size_t) n);
    std::vector<uint32_t> ops; std::vector<char> mds;
    for (int64_t r = 0; r < n; ++r) {
        const bm2_cigar_req &q = reqs[r];
        if (q.read < 0 || q.read >= reads->n_reads) return 1;
        const int64_t ro = reads->offsets[q.read], rl = reads->offsets[q.read + 1] - ro;
        if (q.qb < 0 || q.qe > rl) return 1;
        const int lq = q.qe - q.qb;
        const long long rlen = q.re > q.rb ? q.re - q.rb : 0;
        const long long zc = cigar_z_cells_d(p, idx->l_pac, q.w, lq, q.rb, q.re);
        std::vector<uint8_t> zbuf((size_t) zc + 1, 0xEE);
        std::vector<int32_t> he((size_t) 2 * (lq > 0 ? lq + 1 : 1), 0x5A5A5A5A);
        std::vector<uint32_t> cig((size_t) (lq > 0 ? lq : 0) + (size_t) rlen + 2);
        std::vector<char> md((size_t) 2 * (lq > 0 ? lq : 0) + (size_t) 7 * rlen + 16);
        CigarZ z = { zbuf.data(), 1 };
        bm2_cigar_rec &o = recs[(size_t) r];
        int score = INT32_MIN, nc = 0, nm = -1, nmd = 0;
        gen_cigar_d(p, idx->l_pac, idx->ref_string, q.w, lq, reads->codes + ro + q.qb, q.rb, q.re, he.data(), z, &score, cig.data(), &nc, &nm, md.data(), &nmd);
        if ((size_t) nc > cig.size() || (size_t) nmd > md.size()) return 2;
        o.score = score; o.n_cigar = nc; o.nm = nm; o.n_md = nmd; o.cigar_off = (int64_t) ops.size(); o.md_off = (int64_t) mds.size();
        ops.insert(ops.end(), cig.begin(), cig.begin() + nc);
        mds.insert(mds.end(), md.begin(), md.begin() + nmd);
    }
    *recs_out = (bm2_cigar_rec *) malloc(sizeof(bm2_cigar_rec) * (size_t) (n + 1)); memcpy(*recs_out, recs.data(), sizeof(bm2_cigar_rec) * (size_t) n);
    *cigar_out = (uint32_t *) malloc(4 * (ops.size() + 1)); memcpy(*cigar_out, ops.data(), 4 * ops.size());
    *md_out = (char *) malloc(mds.size() + 1); memcpy(*md_out, mds.data(), mds.size());
    *n_ops_out = (int64_t) ops.size(); *n_md_out = (int64_t) mds.size();
    return 0;
}
