// mate_emul.cpp — TEST-ONLY host build of the mate-rescue device logic (bwa-mem2_b200/csrc/mate_device.cuh) so that it can be checked
// against the oracle's mem_matesw restatement (pinned to the reference's own function).  Never part of the product.
#include <vector>
#include <cstring>
#include "mate_device.cuh"

// regs of pair p: reads 2p, 2p+1 (read_off has n_reads + 1 entries).  out0/out1: the two reads' regions after the rescue block,
// off0/off1 their per-pair offsets (n_pairs + 1).  Returns the number of alignments done, -1 on a scratch overflow.
extern "C" long long emul_mate_rescue(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs,
                                      const int64_t *read_off, const int32_t *pes_lh, bm2_alnreg_t *out, int64_t out_cap, int64_t *out_off)
{
    ContigView cv; cv.l_pac = idx->l_pac; cv.n_seqs = idx->n_seqs; cv.ann_off = idx->ann_offset; cv.ann_len = idx->ann_len; cv.ann_alt = idx->ann_is_alt;
    ExtParams ep; ep.a = opt->a; ep.b = opt->b; ep.o_del = opt->o_del; ep.e_del = opt->e_del; ep.o_ins = opt->o_ins; ep.e_ins = opt->e_ins; ep.w = opt->w;
    ep.pen_clip5 = opt->pen_clip5; ep.pen_clip3 = opt->pen_clip3; ep.max_chain_gap = opt->max_chain_gap; ep.mask_level_redun = opt->mask_level_redun;
    memcpy(ep.mat, opt->mat, 25);
    MatePes pes;
    for (int d = 0; d < 4; ++d) { pes.low[d] = pes_lh[3 * d]; pes.high[d] = pes_lh[3 * d + 1]; pes.failed[d] = pes_lh[3 * d + 2]; }
    long long total = 0; int64_t pos = 0; int overflow = 0;
    out_off[0] = 0;
    for (int p = 0; p < reads->n_reads >> 1; ++p) {
        const uint8_t *seq[2]; int l_seq[2], n[2];
        std::vector<bm2_alnreg_t> a[2], b[2];
        int max_l = 0;
        for (int i = 0; i < 2; ++i) {
            const int r = 2 * p + i;
            seq[i] = reads->codes + reads->offsets[r]; l_seq[i] = (int) (reads->offsets[r + 1] - reads->offsets[r]);
            n[i] = (int) (read_off[r + 1] - read_off[r]);
            if (l_seq[i] > max_l) max_l = l_seq[i];
        }
        for (int i = 0; i < 2; ++i) {
            a[i].assign(regs + read_off[2 * p + i], regs + read_off[2 * p + i + 1]);
            const int calls = n[!i] < opt->max_matesw ? n[!i] : opt->max_matesw;        // mem_matesw calls that can add to read i
            a[i].resize((size_t) n[i] + 4 * (size_t) calls + 4);
            b[i].resize((size_t) n[i] + 1);
        }
        const size_t nreg = a[0].size() + a[1].size();
        const int tcap = mate_window_max_d(pes, max_l) + 16;            // the driver sizes the window scratch from the statistics
        std::vector<uint8_t> rev((size_t) max_l + 1), tmp((size_t) tcap);
        std::vector<int32_t> ksw((size_t) 3 * (max_l + 16)), bsc((size_t) tcap / 2 + 2), bpos((size_t) tcap / 2 + 2), idxv(nreg + 8);
        std::vector<TailSortKey> keys(nreg + 8);
        MateScratch sc = { rev.data(), tmp.data(), tcap, ksw.data(), bsc.data(), bpos.data(), tcap / 2 + 2, idxv.data(), keys.data() };
        bm2_alnreg_t *ap[2] = { a[0].data(), a[1].data() }, *bp[2] = { b[0].data(), b[1].data() };
        total += mate_rescue_pair_d(cv, ep, opt->min_seed_len, opt->pen_unpaired, opt->max_matesw, pes, idx->ref_string, seq, l_seq, ap, n, bp, sc, &overflow);
        for (int i = 0; i < 2; ++i) {
            if (pos + n[i] > out_cap) return -2;
            memcpy(out + pos, a[i].data(), sizeof(bm2_alnreg_t) * (size_t) n[i]);
            pos += n[i];
            out_off[2 * p + i + 1] = pos;
        }
    }
    return overflow ? -1 : total;
}
