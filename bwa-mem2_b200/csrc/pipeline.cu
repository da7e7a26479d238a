// pipeline.cu — seam 2: index upload and the seed -> chain -> extend pipeline (under construction).
#include "bm2_common.cuh"
#include "bm2_ctx.h"

int bm2_upload_index(bm2_ctx *ctx, const bm2_index_desc *idx) {
    bm2_ctx *ctx_for_error = ctx;
    auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
        void *p = nullptr;
        BM2_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 1));
        ctx->idx_allocs.push_back(p);
        if (bytes) BM2_CUDA_OK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        *dst = p;
        return 0;
    };
    DevIndex &d = ctx->idx;
    d.N = idx->reference_seq_len; d.l_pac = idx->l_pac; d.sentinel = idx->sentinel_index;
    for (int i = 0; i < 5; ++i) d.count[i] = idx->count[i];
    d.n_seqs = idx->n_seqs;
    size_t n_occ = (size_t) (d.N >> 6) + 1, n_sa = (size_t) (d.N >> 3) + 1;
    if (up(idx->cp_occ, n_occ * sizeof(bm2_cp_occ), (const void **) &d.cp_occ)) return 1;
    if (up(idx->sa_ms_byte, n_sa, (const void **) &d.sa_ms)) return 1;
    if (up(idx->sa_ls_word, n_sa * 4, (const void **) &d.sa_ls)) return 1;
    if (up(idx->ref_string, (size_t) d.l_pac * 2, (const void **) &d.ref)) return 1;
    if (up(idx->ann_offset, (size_t) d.n_seqs * 8, (const void **) &d.ann_off)) return 1;
    if (up(idx->ann_len, (size_t) d.n_seqs * 4, (const void **) &d.ann_len)) return 1;
    std::vector<int32_t> zeros;
    const int32_t *alt = idx->ann_is_alt;
    if (!alt) { zeros.assign(d.n_seqs, 0); alt = zeros.data(); }
    if (up(alt, (size_t) d.n_seqs * 4, (const void **) &d.ann_alt)) return 1;
    d.loaded = true;
    return 0;
}

void bm2_free_index(bm2_ctx *ctx) {
    for (void *p : ctx->idx_allocs) cudaFree(p);
    ctx->idx_allocs.clear();
    ctx->idx.loaded = false;
}

// ---- seam 2 entry points (filled in stage by stage) ---------------------------------------------
extern "C" int bm2_collect_smems(bm2_ctx *ctx, const bm2_read_batch *, bm2_smem_result *) {
    bm2_set_error(ctx, "bm2_collect_smems: not implemented in this build"); return 1;
}
extern "C" int bm2_seed_chain(bm2_ctx *ctx, const bm2_read_batch *, bm2_chain_result *) {
    bm2_set_error(ctx, "bm2_seed_chain: not implemented in this build"); return 1;
}
extern "C" int bm2_seed_chain_extend(bm2_ctx *ctx, const bm2_read_batch *, bm2_reg_result *) {
    bm2_set_error(ctx, "bm2_seed_chain_extend: not implemented in this build"); return 1;
}
extern "C" int bm2_last_stage_ms(const bm2_ctx *ctx, const char *const **names, const float **ms, int *n) {
    if (!ctx) return 1;
    *names = ctx->stage_names.data(); *ms = ctx->stage_ms.data(); *n = (int) ctx->stage_ms.size();
    return 0;
}
