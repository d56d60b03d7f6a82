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

inline int64_t seq_len(const poy_pool *pool, int s) { return pool->h_off[s + 1] - pool->h_off[s]; }

}  // namespace

// Sequence.Align.recost x x cm (src/sequence.ml:1244-1307) of pool sequences against themselves: what DOS.median
// charges for a median with one empty child when the matrix has a non-zero diagonal (src/seqCS.ml:992-996)
poy_status dos_self_recost(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int n, const int32_t *ids, int32_t *cost) {
    void *v;
    poy_status s = poy_scratch(ctx, SL_STORE2, (size_t)n * (8 + 4 + 4) + 256, &v);
    if (s != POY_OK) return s;
    int64_t *d_off = (int64_t *)v;
    int32_t *d_len = (int32_t *)(d_off + n), *d_c = d_len + n;
    std::vector<int64_t> off((size_t)n);
    std::vector<int32_t> len((size_t)n);
    for (int q = 0; q < n; ++q) { off[q] = pool->h_off[ids[q]]; len[q] = (int32_t)seq_len(pool, ids[q]); }
    CK(cudaMemcpyAsync(d_off, off.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_len, len.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_aligned_cost(ctx, cm, n, pool->d_data, pool->d_data, d_off, d_len, 2, d_c));
    CK(cudaMemcpyAsync(cost, d_c, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

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
    bind_device(ctx);
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

// ---- DOS.median on the device ---------------------------------------------------------------------------------------
namespace {

// aligned rows of the linear entry point sit right-justified in their slots: (offset, length) per pair for the
// column-wise kernels (both rows of a pair have the same length)
__global__ void k_rows_of_slots(int n, const int64_t *__restrict__ slot_end, const int *__restrict__ len2, int64_t *off, int *len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l = len2[2 * p];
    off[p] = slot_end[p] - 2 - l;      // rows end at slot start + len1 + len2 = slot end - 2
    len[p] = l;
}
__global__ void k_take_len(int n, const int *__restrict__ len4, int stride, int *len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) len[p] = len4[(size_t)stride * p];
}

}  // namespace

// The alignment half of DOS.median for m pairs with two non-empty children, everything on the device:
//   affine model  Sequence.Align.align_affine_3 under c2_full, shorter sequence first, swaped = len a > len b
//                 (src/seqCS.ml:1051-1056, src/sequence.ml:633-649); the median is right-justified in its slot
//   linear model  Sequence.Align.align_2 (deltaw = None: src/sequence.ml:875-914, 1061-1082) + ancestor_2 of the rows in
//                 (a, b) order (src/seqCS.ml:1058-1071); the median is left-justified in its slot
// d_slot / d_slot_end: start / end of pair q's slot (capacity len a + len b + 2) in d_median; d_mlen[q] = median length.
// Uses the scratch slot SL_STORE2 for the intermediate rows.
poy_status dos_median_device(poy_ctx *ctx, const poy_cm *c2_full, const poy_pool *pool, int m, const int32_t *a, const int32_t *b,
                             const int64_t *h_slot, const int64_t *d_slot, const int64_t *d_slot_end, int64_t slot_total,
                             int32_t *d_cost, uint8_t *d_median, int32_t *d_mlen, bool *right_justified) {
    std::vector<int32_t> si((size_t)m), sj((size_t)m);
    std::vector<uint8_t> sw((size_t)m);
    for (int q = 0; q < m; ++q) {
        const bool swaped = seq_len(pool, a[q]) > seq_len(pool, b[q]);
        si[q] = swaped ? b[q] : a[q]; sj[q] = swaped ? a[q] : b[q]; sw[q] = swaped ? 1 : 0;
    }
    poy_status s;
    void *v;
    if (c2_full->h.cost_model_type == 1) {
        if ((s = poy_scratch(ctx, SL_STORE2, sizeof(int32_t) * 4 * (size_t)m + 256, &v)) != POY_OK) return s;
        int32_t *d_len4 = (int32_t *)v;
        s = align_split(ctx, c2_full, pool, m, si.data(), sj.data(), sw.data(), d_slot, d_cost, d_median, nullptr, nullptr, nullptr,
                        d_len4, nullptr);
        if (s != POY_OK) return s;
        k_take_len<<<(m + 255) / 256, 256, 0, ctx->stream>>>(m, d_len4, 4, d_mlen);
        ctx->launches++;
        CK(cudaGetLastError());
        *right_justified = true;
        return POY_OK;
    }
    // linear / no-alignment model
    if ((s = ensure_flags(ctx, pool)) != POY_OK) return s;
    std::vector<int32_t> dwh((size_t)m);
    for (int q = 0; q < m; ++q) {
        const int64_t l1 = seq_len(pool, si[q]), l2 = seq_len(pool, sj[q]);
        const int64_t lower = (int64_t)((double)l1 * 0.10), dif = l1 - l2;
        dwh[q] = (int32_t)(std::max(pool->h_gapcnt[a[q]], pool->h_gapcnt[b[q]]) + (dif < lower ? lower / 2 : 2));
    }
    const size_t A = ((size_t)slot_total + 255) & ~(size_t)255;
    if ((s = poy_scratch(ctx, SL_STORE2, 2 * A + (size_t)m * (8 + 4 + 8 + 1) + 1024, &v)) != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v;
    uint8_t *d_r1 = cur; cur += A;
    uint8_t *d_r2 = cur; cur += A;
    int64_t *d_off = (int64_t *)cur; cur += 8 * (size_t)m;
    int32_t *d_len2 = (int32_t *)cur; cur += 8 * (size_t)m;
    int32_t *d_len = (int32_t *)cur; cur += 4 * (size_t)m;
    uint8_t *d_sw = cur;
    CK(cudaMemcpyAsync(d_sw, sw.data(), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));      // sw is a local
    // create_edited_2 allocates len1 + len2 per row (src/sequence.ml:1019-1033): the slots here have two bytes to spare
    s = align_split(ctx, c2_full, pool, m, si.data(), sj.data(), sw.data(), d_slot, d_cost, nullptr, nullptr, d_r1, d_r2, d_len2,
                    nullptr, dwh.data());
    if (s != POY_OK) return s;
    (void)h_slot;
    // rows end at slot start + len1 + len2 = slot end - 2
    k_rows_of_slots<<<(m + 255) / 256, 256, 0, ctx->stream>>>(m, d_slot_end, d_len2, d_off, d_len);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(launch_ancestor_2(ctx, c2_full, m, d_r1, d_r2, d_off, d_len, d_slot, d_median, d_mlen, d_sw, -2));
    *right_justified = false;
    return POY_OK;
}

// DOS.median (src/seqCS.ml:985-1084): an empty child yields the other child; its cost is 0 when c2_original is an
// identity matrix (zero diagonal), else Sequence.Align.recost x x c2_original (src/seqCS.ml:992-996, 1027-1031).
// Otherwise the alignment under c2_FULL (see dos_median_device).  Pair p owns the slot
// [out_off[p], out_off[p] + len_a + len_b + 2) of `median`; its median is RIGHT-justified there, out_len[p] bytes long;
// cost2[p] is the alignment cost.  `c2_original` may be NULL: the identity rule then reads c2_full's flag.
extern "C" poy_status poy_dos_median2(poy_ctx *ctx, const poy_cm *c2_full, const poy_cm *c2_original, const poy_pool *pool, int32_t n,
                                      const int32_t *a, const int32_t *b, const int64_t *out_off, int32_t *cost2, uint8_t *median,
                                      int32_t *out_len) {
    if (!ctx || !c2_full || !pool || n < 0 || (n > 0 && (!a || !b || !out_off || !cost2 || !median || !out_len))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    bind_device(ctx);
    for (int p = 0; p < n; ++p)
        if (a[p] < 0 || a[p] >= pool->nseq || b[p] < 0 || b[p] >= pool->nseq) return dos_fail(ctx, POY_ERR_ARG, "pair index out of range");
    poy_status s = ensure_flags(ctx, pool);
    if (s != POY_OK) return s;
    const poy_cm *ident_cm = c2_original ? c2_original : c2_full;
    std::vector<int32_t> idx, xa, xb, lone, lone_p;
    std::vector<int64_t> slot, slot_end;
    int64_t total = 0;
    for (int p = 0; p < n; ++p) {
        const int64_t cap = seq_len(pool, a[p]) + seq_len(pool, b[p]) + 2;
        total = std::max(total, out_off[p] + cap);
        if (pool->h_empty[a[p]] || pool->h_empty[b[p]]) {
            lone.push_back(pool->h_empty[a[p]] ? b[p] : a[p]); lone_p.push_back(p);
            continue;
        }
        idx.push_back(p); xa.push_back(a[p]); xb.push_back(b[p]);
        slot.push_back(out_off[p]); slot_end.push_back(out_off[p] + cap);
    }
    const int m = (int)idx.size(), nl = (int)lone.size();
    if (m > 0) {
        void *v;
        const size_t A = ((size_t)total + 255) & ~(size_t)255;
        if ((s = poy_scratch(ctx, SL_STORE, A + (size_t)m * (8 + 8 + 4 + 4) + 1024, &v)) != POY_OK) return s;
        uint8_t *cur = (uint8_t *)v;
        uint8_t *d_med = cur; cur += A;
        int64_t *d_slot = (int64_t *)cur; cur += 8 * (size_t)m;
        int64_t *d_slot_end = (int64_t *)cur; cur += 8 * (size_t)m;
        int32_t *d_cost = (int32_t *)cur; cur += 4 * (size_t)m;
        int32_t *d_mlen = (int32_t *)cur;
        CK(cudaMemcpyAsync(d_slot, slot.data(), 8 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_slot_end, slot_end.data(), 8 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        bool rj = true;
        s = dos_median_device(ctx, c2_full, pool, m, xa.data(), xb.data(), slot.data(), d_slot, d_slot_end, total, d_cost, d_med,
                              d_mlen, &rj);
        if (s != POY_OK) return s;
        std::vector<int32_t> c((size_t)m), ml((size_t)m);
        CK(cudaMemcpyAsync(c.data(), d_cost, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ml.data(), d_mlen, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int q = 0; q < m; ++q) {
            if (ml[q] < 0) return dos_fail(ctx, POY_ERR_ARG, "median should not be 0");
            cost2[idx[q]] = c[q]; out_len[idx[q]] = ml[q];
            const int64_t from = rj ? slot_end[q] - ml[q] : slot[q];
            CK(cudaMemcpyAsync(median + slot_end[q] - ml[q], d_med + from, (size_t)ml[q], cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (nl > 0) {
        std::vector<int32_t> rc((size_t)nl, 0);
        if (!ident_cm->h.is_identity) {
            if ((s = dos_self_recost(ctx, ident_cm, pool, nl, lone.data(), rc.data())) != POY_OK) return s;
        }
        for (int q = 0; q < nl; ++q) {
            const int p = lone_p[q], keep = lone[q];
            const int64_t len = seq_len(pool, keep), cap = seq_len(pool, a[p]) + seq_len(pool, b[p]) + 2;
            CK(cudaMemcpyAsync(median + out_off[p] + cap - len, pool->d_data + pool->h_off[keep], (size_t)len, cudaMemcpyDeviceToHost,
                               ctx->stream));
            cost2[p] = rc[q]; out_len[p] = (int32_t)len;
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return POY_OK;
}

extern "C" poy_status poy_dos_median(poy_ctx *ctx, const poy_cm *c2_full, const poy_pool *pool, int32_t n, const int32_t *a,
                                     const int32_t *b, const int64_t *out_off, int32_t *cost2, uint8_t *median,
                                     int32_t *out_len) {
    return poy_dos_median2(ctx, c2_full, nullptr, pool, n, a, b, out_off, cost2, median, out_len);
}
