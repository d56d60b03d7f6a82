// Batch twins of SeqCS.DOS.distance and SeqCS.DOS.median (src/seqCS.ml:701-774, 985-1084): the per-locus policy the
// OCaml side applies around the alignment stubs -- empty-sequence rules, shorter-first ordering, the `swaped` flag
// (src/sequence.ml:633-649), deltaw for the linear model (src/sequence.ml:868-925) -- restated once in C++ above the
// batch entry points, so that the candidate seam can hand over (a[p], b[p]) in any order.  Host-side composition only:
// every alignment goes through poy_batch_cost_affine / poy_batch_cost_linear / poy_batch_align_affine.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

poy_status dos_fail(poy_ctx *ctx, poy_status s, const char *msg) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", msg);
    return s;
}

// Sequence.is_empty (src/sequence.ml:241-251) and Sequence.count_gaps (seq_CAML_count, src/seq.c:644-669) of every pool
// sequence, computed once per pool from the device bytes (pools made by poy_pool_from_device have no host copy) and
// cached in the pool.
poy_status ensure_flags(poy_ctx *ctx, const poy_pool *cpool) {
    poy_pool *pool = const_cast<poy_pool *>(cpool);     // lazily filled cache fields
    if (pool->h_empty) return POY_OK;
    cudaSetDevice(ctx->device);
    std::vector<uint8_t> bytes((size_t)std::max<int64_t>(pool->nbytes, 1));
    cudaError_t e = cudaMemcpyAsync(bytes.data(), pool->d_data, (size_t)pool->nbytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return dos_fail(ctx, POY_ERR_CUDA, cudaGetErrorString(e));
    uint8_t *em = (uint8_t *)malloc((size_t)pool->nseq + 1);
    int32_t *gc = (int32_t *)malloc(sizeof(int32_t) * ((size_t)pool->nseq + 1));
    if (!em || !gc) { free(em); free(gc); return dos_fail(ctx, POY_ERR_NOMEM, "host allocation failed"); }
    for (int q = 0; q < pool->nseq; ++q) {
        int32_t nongap = 0, gapbit = 0;
        for (int64_t x = pool->h_off[q]; x < pool->h_off[q + 1]; ++x) { nongap += bytes[x] != 16; gapbit += (bytes[x] & 16) != 0; }
        em[q] = nongap == 0;
        gc[q] = gapbit;
    }
    pool->h_gapcnt = gc;
    pool->h_empty = em;
    return POY_OK;
}

inline int64_t seq_len(const poy_pool *pool, int s) { return pool->h_off[s + 1] - pool->h_off[s]; }

}  // namespace

// DOS.distance: cost-only alignment under c2_ORIGINAL; an empty sequence on either side yields missing_distance
// (src/seqCS.ml:705-709).  Affine model: Sequence.Align.cost_2 = algn_CAML_cost_affine_3, either order.  Linear model:
// shorter first, deltaw = max(|len a - len b|, 8) (src/seqCS.ml:716-718) folded into deltawh = gaps + deltaw_calc
// (src/sequence.ml:875-914).
extern "C" poy_status poy_dos_distance(poy_ctx *ctx, const poy_cm *c2_original, const poy_pool *pool, int32_t n,
                                       const int32_t *a, const int32_t *b, int32_t missing_distance, int32_t *cost) {
    if (!ctx || !c2_original || !pool || n < 0 || (n > 0 && (!a || !b || !cost))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    for (int p = 0; p < n; ++p)
        if (a[p] < 0 || a[p] >= pool->nseq || b[p] < 0 || b[p] >= pool->nseq) return dos_fail(ctx, POY_ERR_ARG, "pair index out of range");
    poy_status s = ensure_flags(ctx, pool);
    if (s != POY_OK) return s;
    const bool affine = c2_original->h.cost_model_type == 1;
    std::vector<int32_t> idx, s1, s2, dwh;
    idx.reserve(n); s1.reserve(n); s2.reserve(n);
    for (int p = 0; p < n; ++p) {
        if (pool->h_empty[a[p]] || pool->h_empty[b[p]]) { cost[p] = missing_distance; continue; }
        idx.push_back(p);
        if (affine) { s1.push_back(a[p]); s2.push_back(b[p]); continue; }
        const int64_t la = seq_len(pool, a[p]), lb = seq_len(pool, b[p]);
        const bool swaped = la > lb;
        const int x1 = swaped ? b[p] : a[p], x2 = swaped ? a[p] : b[p];
        const int64_t l1 = seq_len(pool, x1), l2 = seq_len(pool, x2);
        const int64_t deltaw = std::max<int64_t>(la > lb ? la - lb : lb - la, 8);
        const int64_t lower = (int64_t)((double)l1 * 0.10), dif = l1 - l2;
        const int64_t dcalc = dif < lower ? lower : deltaw;
        s1.push_back(x1); s2.push_back(x2);
        dwh.push_back((int32_t)(std::max(pool->h_gapcnt[a[p]], pool->h_gapcnt[b[p]]) + dcalc));
    }
    const int m = (int)idx.size();
    if (m == 0) return POY_OK;
    std::vector<int32_t> c((size_t)m);
    s = affine ? poy_batch_cost_affine(ctx, c2_original, pool, m, s1.data(), s2.data(), c.data())
               : poy_batch_cost_linear(ctx, c2_original, pool, m, s1.data(), s2.data(), dwh.data(), c.data());
    if (s != POY_OK) return s;
    for (int q = 0; q < m; ++q) cost[idx[q]] = c[q];
    return POY_OK;
}

// DOS.median, affine model: an empty child yields the other child with cost 0 (src/seqCS.ml:991-1039); otherwise
// Sequence.Align.align_affine_3 under c2_FULL with the shorter sequence first and swaped = len a > len b
// (src/sequence.ml:633-649).  Pair p owns the slot [out_off[p], out_off[p] + len_a + len_b + 2) of `median`, its median
// sequence is right-justified there and out_len[p] bytes long; cost2[p] is the alignment cost.
extern "C" poy_status poy_dos_median(poy_ctx *ctx, const poy_cm *c2_full, const poy_pool *pool, int32_t n, const int32_t *a,
                                     const int32_t *b, const int64_t *out_off, int32_t *cost2, uint8_t *median,
                                     int32_t *out_len) {
    if (!ctx || !c2_full || !pool || n < 0 || (n > 0 && (!a || !b || !out_off || !cost2 || !median || !out_len))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    if (c2_full->h.cost_model_type != 1)
        return dos_fail(ctx, POY_ERR_MODEL, "poy_dos_median restates the affine branch of DOS.median; linear models compose "
                                            "poy_batch_align_linear + poy_batch_ancestor_2");
    for (int p = 0; p < n; ++p)
        if (a[p] < 0 || a[p] >= pool->nseq || b[p] < 0 || b[p] >= pool->nseq) return dos_fail(ctx, POY_ERR_ARG, "pair index out of range");
    poy_status s = ensure_flags(ctx, pool);
    if (s != POY_OK) return s;
    std::vector<int32_t> idx, si, sj;
    std::vector<uint8_t> sw;
    std::vector<int64_t> off;
    for (int p = 0; p < n; ++p) {
        if (pool->h_empty[a[p]] || pool->h_empty[b[p]]) continue;
        const bool swaped = seq_len(pool, a[p]) > seq_len(pool, b[p]);
        idx.push_back(p);
        si.push_back(swaped ? b[p] : a[p]); sj.push_back(swaped ? a[p] : b[p]);
        sw.push_back(swaped ? 1 : 0);
        off.push_back(out_off[p]);
    }
    const int m = (int)idx.size();
    if (m > 0) {
        std::vector<int32_t> c((size_t)m), len4(4 * (size_t)m);
        s = poy_batch_align_affine(ctx, c2_full, pool, m, si.data(), sj.data(), sw.data(), off.data(), c.data(), median, nullptr,
                                   nullptr, nullptr, len4.data(), nullptr);
        if (s != POY_OK) return s;
        for (int q = 0; q < m; ++q) { cost2[idx[q]] = c[q]; out_len[idx[q]] = len4[4 * (size_t)q]; }
    }
    // empty children (after the batch: its read-back covers the whole output range)
    cudaSetDevice(ctx->device);
    bool copied = false;
    for (int p = 0; p < n; ++p) {
        const bool ea = pool->h_empty[a[p]] != 0, eb = pool->h_empty[b[p]] != 0;
        if (!ea && !eb) continue;
        const int keep = ea ? b[p] : a[p];
        const int64_t len = seq_len(pool, keep), cap = seq_len(pool, a[p]) + seq_len(pool, b[p]) + 2;
        cudaError_t e = cudaMemcpyAsync(median + out_off[p] + cap - len, pool->d_data + pool->h_off[keep], (size_t)len,
                                        cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) return dos_fail(ctx, POY_ERR_CUDA, cudaGetErrorString(e));
        cost2[p] = 0; out_len[p] = (int32_t)len;
        copied = true;
    }
    if (copied) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return dos_fail(ctx, POY_ERR_CUDA, cudaGetErrorString(e));
    }
    return POY_OK;
}
