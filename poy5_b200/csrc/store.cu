// Device-resident node store: the sequences of a tree pass -- observed leaves, directional medians,
// the temporaries of a swap round -- live in ONE growing pool in HBM.  DOS.median appends its result to the pool
// without leaving the device (only 12 bytes per pair come back: id, length, cost2), so a downpass level or a chunk
// of SPR candidates never re-uploads a sequence; the per-base gap parameters are computed once per appended
// sequence.  In the reference every node owns an OCaml `Sequence.s` custom block (src/seq.h:52-61) and
// SeqCS.DOS.median allocates a fresh one per call (src/seqCS.ml:985-1084, `create`); the store is that heap, with
// stack discipline (poy_store_truncate) instead of a garbage collector.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

struct poy_store {
    poy_pool *pool;       // owns_data = true, cap_bytes / cap_seqs = what the arrays are allocated for
};

namespace {

poy_status st_fail(poy_ctx *ctx, poy_status s, const char *msg) { return poy_fail(ctx, s, msg); }

// (re)allocate the device arrays for at least need_bytes / need_seqs, keeping the stored sequences
poy_status store_reserve(poy_ctx *ctx, poy_store *st, int64_t need_bytes, int64_t need_seqs) {
    poy_pool *p = st->pool;
    if (need_bytes <= p->cap_bytes && need_seqs <= p->cap_seqs) return POY_OK;
    if (need_seqs > 0x7ffffff0) return st_fail(ctx, POY_ERR_ARG, "node store: too many sequences");
    int64_t nb = p->cap_bytes, ns = p->cap_seqs;
    while (nb < need_bytes) nb += nb / 2 + (1 << 20);
    while (ns < need_seqs) ns += ns / 2 + 1024;
    CK(cudaStreamSynchronize(ctx->stream));
    uint8_t *nd = nullptr; int64_t *no = nullptr;
    size_t cd = 0, co = 0;
    CK(cached_alloc(ctx, (void **)&nd, (size_t)nb, &cd));
    CK(cached_alloc(ctx, (void **)&no, sizeof(int64_t) * ((size_t)ns + 1), &co));
    if (p->nbytes > 0) CK(cudaMemcpyAsync(nd, p->d_data, (size_t)p->nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(no, p->d_off, sizeof(int64_t) * ((size_t)p->nseq + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    cached_free(ctx, p->d_data, p->caps[6]); cached_free(ctx, p->d_off, p->caps[7]);
    p->d_data = nd; p->d_off = no; p->caps[6] = cd; p->caps[7] = co;
    cached_free(ctx, p->d_rowp, p->caps[0]); cached_free(ctx, p->d_colp, p->caps[1]); cached_free(ctx, p->d_rowpk, p->caps[2]);
    cached_free(ctx, p->d_h0, p->caps[3]); cached_free(ctx, p->d_g0, p->caps[4]); cached_free(ctx, p->d_gapfree, p->caps[5]);
    cached_free(ctx, p->d_flags, p->caps[8]);
    p->d_rowp = p->d_colp = nullptr; p->d_rowpk = nullptr; p->d_h0 = p->d_g0 = nullptr; p->d_gapfree = p->d_flags = nullptr;
    poy_status s = pool_alloc(ctx, p, nb, (int32_t)ns);
    if (s != POY_OK) return s;
    p->params_upto = 0;         // the parameter arrays are new
    p->h_gapfree = (uint8_t *)realloc(p->h_gapfree, (size_t)ns + 1);
    p->h_empty = (uint8_t *)realloc(p->h_empty, (size_t)ns + 1);
    p->h_gapcnt = (int32_t *)realloc(p->h_gapcnt, sizeof(int32_t) * ((size_t)ns + 1));
    p->h_off = (int64_t *)realloc(p->h_off, sizeof(int64_t) * ((size_t)ns + 1));
    if (!p->h_gapfree || !p->h_empty || !p->h_gapcnt || !p->h_off) return st_fail(ctx, POY_ERR_NOMEM, "host allocation failed");
    p->cap_bytes = nb; p->cap_seqs = (int32_t)ns;
    return POY_OK;
}

// warp per pair: median q (m bytes, at the start or the end of its slot) -> store bytes [dst[q], dst[q] + m)
__global__ void __launch_bounds__(256) k_gather_medians(int n, const uint8_t *__restrict__ med, const int64_t *__restrict__ slot,
                                                        const int64_t *__restrict__ slot_end, const int *__restrict__ mlen,
                                                        int right_justified, uint8_t *data, const int64_t *__restrict__ dst) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= n) return;
    const int m = mlen[q];
    const uint8_t *src = med + (right_justified ? slot_end[q] - m : slot[q]);
    uint8_t *d = data + dst[q];
    for (int x = lane; x < m; x += 32) d[x] = src[x];
}

}  // namespace

extern "C" poy_status poy_store_create(poy_ctx *ctx, int64_t cap_bytes, int32_t cap_seqs, poy_store **out) {
    bind_device(ctx);
    if (!ctx || !out) return POY_ERR_ARG;
    *out = nullptr;
    if (cap_bytes < (1 << 16)) cap_bytes = 1 << 16;
    if (cap_seqs < 64) cap_seqs = 64;
    const int64_t zero = 0;
    poy_pool *p;
    poy_status s = pool_new(ctx, &zero, 0, cap_seqs, &p);
    if (s != POY_OK) return s;
    p->owns_data = true;
    cudaError_t e = cached_alloc(ctx, (void **)&p->d_data, (size_t)cap_bytes, &p->caps[6]);
    if (e == cudaSuccess) e = cached_alloc(ctx, (void **)&p->d_off, sizeof(int64_t) * ((size_t)cap_seqs + 1), &p->caps[7]);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_off, 0, sizeof(int64_t), ctx->stream);
    if (e != cudaSuccess) { poy_pool_free(ctx, p); return poy_cuda_fail(ctx, e, "poy_store_create"); }
    s = pool_alloc(ctx, p, cap_bytes, cap_seqs);
    if (s != POY_OK) { poy_pool_free(ctx, p); return s; }
    p->cap_bytes = cap_bytes; p->cap_seqs = cap_seqs;
    poy_store *st = new poy_store;
    st->pool = p;
    *out = st;
    return POY_OK;
}

extern "C" void poy_store_free(poy_ctx *ctx, poy_store *st) {
    if (!st) return;
    poy_pool_free(ctx, st->pool);
    delete st;
}

extern "C" const poy_pool *poy_store_pool(const poy_store *st) { return st ? st->pool : nullptr; }
extern "C" int32_t poy_store_count(const poy_store *st) { return st ? st->pool->nseq : 0; }
extern "C" int64_t poy_store_bytes(const poy_store *st) { return st ? st->pool->nbytes : 0; }

extern "C" poy_status poy_store_append(poy_ctx *ctx, poy_store *st, const uint8_t *data, const int64_t *offsets, int32_t nseq,
                                       int32_t *first_id) {
    bind_device(ctx);
    if (!ctx || !st || nseq < 0 || (nseq > 0 && (!data || !offsets))) return POY_ERR_ARG;
    poy_pool *p = st->pool;
    if (first_id) *first_id = p->nseq;
    if (nseq == 0) return POY_OK;
    if (offsets[0] != 0) return st_fail(ctx, POY_ERR_ARG, "offsets must start at 0");
    for (int s = 0; s < nseq; ++s)
        if (offsets[s + 1] <= offsets[s]) return st_fail(ctx, POY_ERR_ARG, "every sequence needs at least its leading gap");
    const int64_t add = offsets[nseq];
    poy_status s = store_reserve(ctx, st, p->nbytes + add, (int64_t)p->nseq + nseq);
    if (s != POY_OK) return s;
    for (int q = 1; q <= nseq; ++q) p->h_off[p->nseq + q] = p->nbytes + offsets[q];
    CK(cudaMemcpyAsync(p->d_data + p->nbytes, data, (size_t)add, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(p->d_off + p->nseq + 1, p->h_off + p->nseq + 1, sizeof(int64_t) * (size_t)nseq, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));     // the caller's buffers may be pageable and reused
    p->nseq += nseq; p->nbytes += add;
    return POY_OK;
}

// stack discipline: forget every sequence with id >= nseq (the temporaries of a finished chunk of candidates)
extern "C" poy_status poy_store_truncate(poy_ctx *ctx, poy_store *st, int32_t nseq) {
    if (!ctx || !st || nseq < 0 || nseq > st->pool->nseq) return POY_ERR_ARG;
    poy_pool *p = st->pool;
    p->nseq = nseq; p->nbytes = p->h_off[nseq];
    p->params_upto = std::min(p->params_upto, nseq);
    p->flags_upto = std::min(p->flags_upto, nseq);
    return POY_OK;
}

extern "C" poy_status poy_store_lengths(const poy_store *st, int32_t n, const int32_t *ids, int32_t *len) {
    if (!st || n < 0 || (n > 0 && (!ids || !len))) return POY_ERR_ARG;
    const poy_pool *p = st->pool;
    for (int q = 0; q < n; ++q) {
        if (ids[q] < 0 || ids[q] >= p->nseq) return POY_ERR_ARG;
        len[q] = (int32_t)(p->h_off[ids[q] + 1] - p->h_off[ids[q]]);
    }
    return POY_OK;
}

// sequences ids[0..n) -> host: sequence q is written to out[out_off[q] ...]
extern "C" poy_status poy_store_read(poy_ctx *ctx, const poy_store *st, int32_t n, const int32_t *ids, const int64_t *out_off,
                                     uint8_t *out) {
    bind_device(ctx);
    if (!ctx || !st || n < 0 || (n > 0 && (!ids || !out_off || !out))) return POY_ERR_ARG;
    const poy_pool *p = st->pool;
    for (int q = 0; q < n; ++q) {
        if (ids[q] < 0 || ids[q] >= p->nseq) return st_fail(ctx, POY_ERR_ARG, "sequence id out of range");
        CK(cudaMemcpyAsync(out + out_off[q], p->d_data + p->h_off[ids[q]], (size_t)(p->h_off[ids[q] + 1] - p->h_off[ids[q]]),
                           cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

// DOS.median over store ids; the medians are APPENDED to the store.  out_id[p] = id of pair p's median -- for a pair
// with an empty child that is the other child's id (sequences are immutable, so no copy is made); out_len[p] its
// length; cost2[p] the alignment cost (0, or the self-recost under a non-identity c2_original, for an empty child).
extern "C" poy_status poy_store_median(poy_ctx *ctx, poy_store *st, const poy_cm *c2_full, const poy_cm *c2_original, int32_t n,
                                       const int32_t *a, const int32_t *b, int32_t *out_id, int32_t *out_len, int32_t *cost2) {
    bind_device(ctx);
    if (!ctx || !st || !c2_full || n < 0 || (n > 0 && (!a || !b || !out_id || !cost2))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    poy_pool *p = st->pool;
    for (int q = 0; q < n; ++q)
        if (a[q] < 0 || a[q] >= p->nseq || b[q] < 0 || b[q] >= p->nseq) return st_fail(ctx, POY_ERR_ARG, "sequence id out of range");
    poy_status s = ensure_flags(ctx, p);
    if (s != POY_OK) return s;
    const poy_cm *ident_cm = c2_original ? c2_original : c2_full;
    std::vector<int32_t> idx, xa, xb, lone, lone_p;
    std::vector<int64_t> slot, slot_end;
    int64_t total = 0;
    for (int q = 0; q < n; ++q) {
        if (p->h_empty[a[q]] || p->h_empty[b[q]]) {
            lone.push_back(p->h_empty[a[q]] ? b[q] : a[q]); lone_p.push_back(q);
            continue;
        }
        const int64_t cap = (p->h_off[a[q] + 1] - p->h_off[a[q]]) + (p->h_off[b[q] + 1] - p->h_off[b[q]]) + 2;
        idx.push_back(q); xa.push_back(a[q]); xb.push_back(b[q]);
        slot.push_back(total); slot_end.push_back(total + cap);
        total += cap;
    }
    const int m = (int)idx.size(), nl = (int)lone.size();
    if (nl > 0) {
        std::vector<int32_t> rc((size_t)nl, 0);
        if (!ident_cm->h.is_identity && (s = dos_self_recost(ctx, ident_cm, p, nl, lone.data(), rc.data())) != POY_OK) return s;
        for (int q = 0; q < nl; ++q) {
            out_id[lone_p[q]] = lone[q]; cost2[lone_p[q]] = rc[q];
            if (out_len) out_len[lone_p[q]] = (int32_t)(p->h_off[lone[q] + 1] - p->h_off[lone[q]]);
        }
    }
    if (m == 0) return POY_OK;
    void *v, *v_pin;
    const size_t A = ((size_t)total + 255) & ~(size_t)255;
    if ((s = poy_scratch(ctx, SL_STORE, A + (size_t)m * (8 + 8 + 8 + 4 + 4) + 1024, &v)) != POY_OK) return s;
    if ((s = poy_pinned(ctx, 5, (size_t)m * (8 + 8 + 8 + 4 + 4) + 64, &v_pin)) != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v;
    uint8_t *d_med = cur; cur += A;
    int64_t *d_slot = (int64_t *)cur; cur += 8 * (size_t)m;
    int64_t *d_slot_end = (int64_t *)cur; cur += 8 * (size_t)m;
    int64_t *d_dst = (int64_t *)cur; cur += 8 * (size_t)m;
    int32_t *d_cost = (int32_t *)cur; cur += 4 * (size_t)m;
    int32_t *d_mlen = (int32_t *)cur;
    int64_t *h_slot = (int64_t *)v_pin, *h_slot_end = h_slot + m, *h_dst = h_slot_end + m;
    int32_t *h_cost = (int32_t *)(h_dst + m), *h_mlen = h_cost + m;
    memcpy(h_slot, slot.data(), 8 * (size_t)m); memcpy(h_slot_end, slot_end.data(), 8 * (size_t)m);
    CK(cudaMemcpyAsync(d_slot, h_slot, 16 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));   // slot + slot_end are contiguous
    bool rj = true;
    s = dos_median_device(ctx, c2_full, p, m, xa.data(), xb.data(), h_slot, d_slot, d_slot_end, total, d_cost, d_med, d_mlen, &rj);
    if (s != POY_OK) return s;
    CK(cudaMemcpyAsync(h_cost, d_cost, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));    // cost + mlen are contiguous
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t add = 0;
    for (int q = 0; q < m; ++q) {
        if (h_mlen[q] <= 0) return st_fail(ctx, POY_ERR_ARG, "median should not be 0");
        add += h_mlen[q];
    }
    if ((s = store_reserve(ctx, st, p->nbytes + add, (int64_t)p->nseq + m)) != POY_OK) return s;
    int64_t at = p->nbytes;
    for (int q = 0; q < m; ++q) {
        h_dst[q] = at;
        at += h_mlen[q];
        p->h_off[p->nseq + q + 1] = at;
        out_id[idx[q]] = p->nseq + q; cost2[idx[q]] = h_cost[q];
        if (out_len) out_len[idx[q]] = h_mlen[q];
    }
    CK(cudaMemcpyAsync(d_dst, h_dst, 8 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(p->d_off + p->nseq + 1, p->h_off + p->nseq + 1, 8 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    k_gather_medians<<<(m + 7) / 8, 256, 0, ctx->stream>>>(m, d_med, d_slot, d_slot_end, d_mlen, rj ? 1 : 0, p->d_data, d_dst);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));      // h_off is pageable and h_dst is reused by the next call
    p->nseq += m; p->nbytes = at;
    return POY_OK;
}

// DOS.distance over store ids (src/seqCS.ml:701-774)
extern "C" poy_status poy_store_distance(poy_ctx *ctx, poy_store *st, const poy_cm *c2_original, int32_t n, const int32_t *a,
                                         const int32_t *b, int32_t missing_distance, int32_t *cost) {
    if (!st) return POY_ERR_ARG;
    return poy_dos_distance(ctx, c2_original, st->pool, n, a, b, missing_distance, cost);
}
