// Host side of libpoy5b200.so: context / pool / cost-model objects, batch orchestration
// (job construction, band-doubling schedule, direction-arena waves) and the C ABI of
// include/poy5_b200.h.  No CPU implementation of any alignment lives here: every
// alignment result comes from the kernels in cost_affine.cu / band_affine.cu.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <chrono>
#include <atomic>
#include <thread>
#include "common.cuh"

// ---- error plumbing ---------------------------------------------------------------------
poy_status poy_fail(poy_ctx *ctx, poy_status s, const char *msg) {
    if (ctx) { strncpy(ctx->err, msg, sizeof(ctx->err) - 1); ctx->err[sizeof(ctx->err) - 1] = 0; }
    return s;
}
poy_status poy_cuda_fail(poy_ctx *ctx, cudaError_t e, const char *where) {
    char buf[400];
    snprintf(buf, sizeof buf, "%s: %s", where, cudaGetErrorString(e));
    return poy_fail(ctx, e == cudaErrorMemoryAllocation ? POY_ERR_NOMEM : POY_ERR_CUDA, buf);
}

// Every entry point binds the calling thread to the context's device first: host threads other than the one that
// created the context start out on device 0.

extern "C" const char *poy_status_string(poy_status s) {
    switch (s) {
        case POY_OK: return "ok";
        case POY_ERR_CUDA: return "CUDA runtime error";
        case POY_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case POY_ERR_ARG: return "bad argument";
        case POY_ERR_ORDER: return "pass the shorter one as first";
        case POY_ERR_COST_RANGE: return "cost model outside the reference's HIGH_NUM domain";
        case POY_ERR_MODEL: return "entry point does not match the cost model type";
        case POY_ERR_NOMEM: return "out of device memory";
    }
    return "unknown";
}
extern "C" const char *poy_last_error(const poy_ctx *ctx) { return ctx ? ctx->err : "no context"; }

// ---- context -------------------------------------------------------------------------------
extern "C" poy_status poy_ctx_create(int device, void *stream, poy_ctx **out) {
    if (!out) return POY_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return POY_ERR_NO_DEVICE;
    poy_ctx *ctx = (poy_ctx *)calloc(1, sizeof(poy_ctx));
    if (!ctx) return POY_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return POY_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { free(ctx); return POY_ERR_NO_DEVICE; }
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->owns_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { free(ctx); return POY_ERR_CUDA; }
        ctx->owns_stream = true;
    }
    for (int a = 0; a < POY_N_AUX; ++a) {
        cudaStreamCreateWithFlags(&ctx->aux[a], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->ev_join[a], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->tb_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fin, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_tb_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_tb_done[1], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_twin_start, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_twin_done, cudaEventDisableTiming);
    {   // direction arena: up to 40% of the free HBM, at most 64 GiB
        size_t fr = 0, tot = 0;
        ctx->arena_limit = 8ull << 30;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
            uint64_t lim = (uint64_t)(fr * 0.4);
            if (lim > (64ull << 30)) lim = 64ull << 30;
            if (lim > ctx->arena_limit) ctx->arena_limit = lim;
        }
    }
    *out = ctx;
    return POY_OK;
}

extern "C" void poy_ctx_destroy(poy_ctx *ctx) {
    bind_device(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->twin) { poy_ctx_destroy(ctx->twin); ctx->twin = nullptr; }
    cudaEventDestroy(ctx->ev_twin_start); cudaEventDestroy(ctx->ev_twin_done);
    cudaStreamSynchronize(ctx->tb_stream);
    cudaStreamDestroy(ctx->tb_stream); cudaEventDestroy(ctx->ev_fin); cudaEventDestroy(ctx->ev_tb_done[0]); cudaEventDestroy(ctx->ev_tb_done[1]);
    for (int s = 0; s < 14; ++s) if (ctx->d_scratch[s]) cudaFree(ctx->d_scratch[s]);
    for (int i = 0; i < ctx->cache_n; ++i) cudaFree(ctx->cache_ptr[i]);
    for (int s = 0; s < 6; ++s) if (ctx->h_pinned[s]) cudaFreeHost(ctx->h_pinned[s]);
    for (int a = 0; a < POY_N_AUX; ++a) { cudaStreamDestroy(ctx->aux[a]); cudaEventDestroy(ctx->ev_join[a]); }
    cudaEventDestroy(ctx->ev_fork);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
}
extern "C" poy_status poy_ctx_set_arena_limit(poy_ctx *ctx, uint64_t bytes) {
    bind_device(ctx);
    if (!ctx || bytes < (1u << 20)) return POY_ERR_ARG;
    ctx->arena_limit = bytes;
    return POY_OK;
}
// give the grow-only device scratch (direction arenas, job arrays, ...) and the block cache back to the driver;
// everything is re-allocated on demand.  For callers that switch from one kind of workload to another.
extern "C" poy_status poy_ctx_trim(poy_ctx *ctx) {
    bind_device(ctx);
    if (!ctx) return POY_ERR_ARG;
    for (poy_ctx *c : { ctx, ctx->twin }) {
        if (!c) continue;
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->tb_stream);
        for (int a = 0; a < POY_N_AUX; ++a) cudaStreamSynchronize(c->aux[a]);
        for (int s = 0; s < 14; ++s)
            if (c->d_scratch[s]) { cudaFree(c->d_scratch[s]); c->d_scratch[s] = nullptr; c->scratch_cap[s] = 0; }
        for (int i = 0; i < c->cache_n; ++i) cudaFree(c->cache_ptr[i]);
        c->cache_n = 0; c->cache_bytes = 0;
    }
    return POY_OK;
}

extern "C" poy_status poy_ctx_synchronize(poy_ctx *ctx) {
    bind_device(ctx);
    if (!ctx) return POY_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}
extern "C" uint64_t poy_ctx_launch_count(const poy_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" poy_status poy_ctx_stats(const poy_ctx *ctx, int64_t out[8]) {
    if (!ctx || !out) return POY_ERR_ARG;
    out[0] = (int64_t)ctx->launches; out[1] = ctx->stat_band_cells; out[2] = ctx->stat_probe; out[3] = ctx->stat_full;
    out[4] = ctx->stat_repeat; out[5] = ctx->stat_rounds; out[6] = ctx->stat_pairs; out[7] = 0;
    return POY_OK;
}

// cached device blocks (all users are ordered on ctx->stream, so stream-ordered reuse is safe)
cudaError_t cached_alloc(poy_ctx *ctx, void **out, size_t bytes, size_t *cap_out) {
    if (bytes < 256) bytes = 256;
    int best = -1;
    for (int i = 0; i < ctx->cache_n; ++i)
        if (ctx->cache_cap[i] >= bytes && ctx->cache_cap[i] <= 4 * bytes && (best < 0 || ctx->cache_cap[i] < ctx->cache_cap[best])) best = i;
    if (best >= 0) {
        *out = ctx->cache_ptr[best]; *cap_out = ctx->cache_cap[best];
        ctx->cache_bytes -= ctx->cache_cap[best];
        ctx->cache_ptr[best] = ctx->cache_ptr[ctx->cache_n - 1]; ctx->cache_cap[best] = ctx->cache_cap[ctx->cache_n - 1];
        --ctx->cache_n;
        return cudaSuccess;
    }
    size_t cap = 256;
    while (cap < bytes) cap += cap / 2 + 256;   // geometric size classes so that blocks get reused
    cudaError_t e = cudaMalloc(out, cap);
    if (e != cudaSuccess) {                      // out of memory: drop the cache and retry with the exact size
        cudaGetLastError();
        for (int i = 0; i < ctx->cache_n; ++i) cudaFree(ctx->cache_ptr[i]);
        ctx->cache_n = 0; ctx->cache_bytes = 0;
        cap = bytes;
        e = cudaMalloc(out, cap);
    }
    *cap_out = cap;
    return e;
}
void cached_free(poy_ctx *ctx, void *p, size_t cap) {
    if (!p) return;
    // bounded by entries AND by total bytes (4 GiB): the cache serves short-lived tree-search pools ...
    bool keep = ctx && ctx->cache_n < 64 && cap <= (512u << 20) && ctx->cache_bytes + cap <= (4ull << 30);
    // ... and the per-base parameter arrays of bulk batches (GBs each: cudaFree of such a block idles the device for
    // 0.1-0.8 s, measured as 0.6-3 s per sweep step of bench.py), but those only while a third of the device stays free
    if (!keep && ctx && ctx->cache_n < 64 && ctx->cache_bytes + cap <= (48ull << 30)) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr >= tot / 3) keep = true;
    }
    if (keep) {
        ctx->cache_ptr[ctx->cache_n] = p; ctx->cache_cap[ctx->cache_n] = cap; ++ctx->cache_n; ctx->cache_bytes += cap;
    } else cudaFree(p);
}

// grow-only scratch slots
poy_status poy_scratch(poy_ctx *ctx, int slot, size_t bytes, void **out) {
    if (ctx->scratch_cap[slot] < bytes) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scratch[slot]) { cudaFree(ctx->d_scratch[slot]); ctx->d_scratch[slot] = nullptr; ctx->scratch_cap[slot] = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ctx->d_scratch[slot], want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&ctx->d_scratch[slot], want); }
        if (e != cudaSuccess && slot != SL_DIR2 && ctx->d_scratch[SL_DIR2]) {   // ... or in the second direction arena
            cudaGetLastError();
            cudaStreamSynchronize(ctx->tb_stream);
            cudaFree(ctx->d_scratch[SL_DIR2]); ctx->d_scratch[SL_DIR2] = nullptr; ctx->scratch_cap[SL_DIR2] = 0;
            e = cudaMalloc(&ctx->d_scratch[slot], want);
        }
        if (e != cudaSuccess) {   // the memory may sit in the block cache of freed pools (this context's or its twin's)
            cudaGetLastError();
            for (poy_ctx *c : { ctx, ctx->twin }) {
                if (!c) continue;
                for (int i = 0; i < c->cache_n; ++i) cudaFree(c->cache_ptr[i]);
                c->cache_n = 0; c->cache_bytes = 0;
            }
            e = cudaMalloc(&ctx->d_scratch[slot], want);
        }
        if (e != cudaSuccess) return poy_cuda_fail(ctx, e, "cudaMalloc(scratch)");
        ctx->scratch_cap[slot] = want;
    }
    *out = ctx->d_scratch[slot];
    return POY_OK;
}
poy_status poy_pinned(poy_ctx *ctx, int slot, size_t bytes, void **out) {
    if (ctx->pinned_cap[slot] < bytes) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_pinned[slot]) { cudaFreeHost(ctx->h_pinned[slot]); ctx->h_pinned[slot] = nullptr; ctx->pinned_cap[slot] = 0; }
        size_t want = bytes + bytes / 4 + 256;
        CK(cudaMallocHost(&ctx->h_pinned[slot], want));
        ctx->pinned_cap[slot] = want;
    }
    *out = ctx->h_pinned[slot];
    return POY_OK;
}

// ---- Cost_matrix.Two_D table construction (src/cost_matrix.ml) --------------------------------
namespace {
const int A_SZ = 5, NCOMB = 31, GAPC = 16;
const int CM_MAX_INT = 0x3fffffff;  // (Int32.max_int) lsr 1, src/cost_matrix.ml:38

inline int &at(int32_t *t, int a, int b) { return t[(a << 5) + b]; }
inline int at(const int32_t *t, int a, int b) { return t[(a << 5) + b]; }

// members of a bitset, highest bit first (BitSet.Int.list_of_packed_max, src/bitSet.ml:343-353)
int members(int v, int *out) {
    int n = 0;
    for (int b = A_SZ - 1; b >= 0; --b) if (v & (1 << b)) out[n++] = 1 << b;
    return n;
}

// fill_best_cost_and_median_for_all_combinations (src/cost_matrix.ml:862-897) with
// test_combinations (:479-506) and cleanup (:700-716).
void fill_all_combinations(poy_cm_host *m) {
    const bool affine = m->cost_model_type == 1;
    for (int i = 1; i <= NCOMB; ++i) {
        int li[A_SZ], ni = members(i, li);
        for (int j = 1; j <= NCOMB; ++j) {
            int lj[A_SZ], nj = members(j, lj);
            int best = 0, c = CM_MAX_INT, w = 0;
            for (int x = 0; x < ni; ++x)
                for (int y = 0; y < nj; ++y) {
                    const int a = li[x], b = lj[y];
                    for (int e = 0; e < A_SZ; ++e) {
                        const int v = 1 << e;
                        const int goa = (affine && v == GAPC && (a & GAPC) && (b & GAPC)) ? m->gap_open : 0;
                        const int tc = at(m->cost, a, v) + at(m->cost, v, b) + goa;
                        if (tc < c) { c = tc; best = v; }
                        else if (tc == c) best |= v;
                    }
                    if (at(m->cost, a, b) > w) w = at(m->cost, a, b);
                }
            if (ni == 1 && nj == 1) m->median[(i << 5) + j] = (uint8_t)(i | j);
            else {
                at(m->cost, i, j) = c;
                int med = best;
                if (affine && med != GAPC && (med & GAPC)) med = GAPC;
                m->median[(i << 5) + j] = (uint8_t)med;
            }
            at(m->worst, i, j) = w;
        }
    }
}

// fill_best_cost_and_median_for_all_combinations_bitwise (src/cost_matrix.ml:721-804)
void fill_bitwise(poy_cm_host *m, bool create_original) {
    int32_t old[1024];
    memcpy(old, m->cost, sizeof old);
    struct Acc { int best, med, worst; };
    auto step = [&](Acc acc, int i, int j) {
        const int cost1 = at(old, i, i) + at(old, i, j), cost2 = at(old, i, j) + at(old, j, j);
        int cij, mij;
        if (cost1 == cost2) { cij = create_original ? cost1 - at(old, i, i) : cost1; mij = i | j; }
        else if (cost1 > cost2) { cij = create_original ? cost2 - at(old, j, j) : cost2; mij = j; }
        else { cij = create_original ? cost1 - at(old, i, i) : cost1; mij = i; }
        if (cij < acc.best) { acc.best = cij; acc.med = mij; }
        else if (cij == acc.best) acc.med |= mij;
        if (cij > acc.worst) acc.worst = cij;
        return acc;
    };
    for (int i = 1; i <= NCOMB; ++i) {
        int li[A_SZ], ni = members(i, li);
        for (int j = 1; j <= NCOMB; ++j) {
            int lj[A_SZ], nj = members(j, lj);
            int median, best, worst;
            if (ni == 1 && nj == 1) {
                if (i == j) {
                    int cii = at(old, i, i);
                    if (!create_original) cii *= 2;
                    median = i; best = worst = cii;
                } else {
                    const int cost1 = at(old, i, j) + at(old, j, j), cost2 = at(old, i, i) + at(old, i, j);
                    int cij;
                    if (cost1 == cost2) { cij = cost1; median = i | j; }
                    else if (cost1 > cost2) { cij = cost2; median = i; }
                    else { cij = cost1; median = j; }
                    if (create_original) cij = at(old, i, j);
                    best = worst = cij;
                }
            } else {
                Acc acc = { CM_MAX_INT, 0, 0 };
                for (int x = 0; x < nj; ++x) for (int y = 0; y < ni; ++y) acc = step(acc, lj[x], li[y]);
                for (int x = 0; x < ni; ++x) for (int y = 0; y < nj; ++y) acc = step(acc, li[x], lj[y]);
                median = acc.med; best = acc.best; worst = acc.worst;
            }
            m->median[(i << 5) + j] = (uint8_t)median;
            at(m->cost, i, j) = best;
            at(m->worst, i, j) = worst;
        }
    }
}

void fill_prepend_tail(poy_cm_host *m) {  // fill_default_prepend_tail, src/cost_matrix.ml:994-1001
    for (int i = 1; i <= NCOMB; ++i) { m->tail[i] = at(m->cost, i, GAPC); m->prepend[i] = at(m->cost, GAPC, i); }
}

// fill_cost_matrix (src/cost_matrix.ml:1140-1189), use_comb = true, level = 0
void fill_cost_matrix(const int32_t single[25], bool create_original, poy_cm_host *m) {
    memset(m, 0, sizeof *m);
    bool pos = true, sym = true, ident = true;
    for (int a = 0; a < A_SZ; ++a)
        for (int b = 0; b < A_SZ; ++b) {
            at(m->cost, 1 << a, 1 << b) = single[a * A_SZ + b];
            if (single[a * A_SZ + b] < 0) pos = false;
            if (single[a * A_SZ + b] != single[b * A_SZ + a]) sym = false;
            if (a == b && single[a * A_SZ + b] != 0) ident = false;
        }
    if (pos && sym && ident) fill_all_combinations(m);
    else fill_bitwise(m, create_original);
    fill_prepend_tail(m);
    m->is_identity = ident ? 1 : 0;          // set_identity / set_metric, src/cost_matrix.ml:1169-1177
    m->is_metric = (pos && sym && ident) ? 1 : 0;
}
}  // namespace

extern "C" poy_status poy_cm_fill(const int32_t single[25], int32_t gap_open, poy_cm_host *full, poy_cm_host *original) {
    if (!single || !full || !original) return POY_ERR_ARG;
    for (int x = 0; x < 25; ++x) if (single[x] < 0) return POY_ERR_COST_RANGE;
    fill_cost_matrix(single, false, full);
    fill_cost_matrix(single, true, original);
    if (gap_open >= 0) {  // set_cost_model (Affine go), src/cost_matrix.ml:1003-1016 via src/data.ml:5937-5964
        poy_cm_host *ms[2] = { full, original };
        for (poy_cm_host *m : ms) {
            m->cost_model_type = 1;
            m->gap_open = gap_open;
            fill_bitwise(m, false);
        }
    }
    return POY_OK;
}

extern "C" int32_t poy_cm_min_non0(const poy_cm_host *cm) {
    int m = 0x3fffffff;  // INT_MAX/2
    for (int x = 0; x < 1024; ++x) if (cm->cost[x] > 0 && cm->cost[x] < m) m = cm->cost[x];
    return m;
}

extern "C" int32_t poy_cm_get_closest(const poy_cm_host *cm, int32_t a, int32_t b) {
    if (a <= 0 || b <= 0 || a > NCOMB || b > NCOMB) return -1;
    if (a == GAPC || b == GAPC) { /* keep b */ }
    else if ((a & GAPC) && (b & GAPC)) b = GAPC;
    else b &= ~GAPC;
    int best = a, cur = CM_MAX_INT;
    for (int e = 0; e < A_SZ; ++e) {  // states_of_code: ascending bit order
        const int x = 1 << e;
        if (!(b & x)) continue;
        const int nc = at(cm->cost, a, x);
        if (nc < cur) { best = x; cur = nc; }
    }
    return best;
}

extern "C" poy_status poy_cm_upload(poy_ctx *ctx, const poy_cm_host *h, poy_cm **out) {
    bind_device(ctx);
    if (!ctx || !h || !out) return POY_ERR_ARG;
    *out = nullptr;
    int max_entry = 0;
    for (int x = 0; x < 1024; ++x) {
        if (h->cost[x] < 0) return poy_fail(ctx, POY_ERR_COST_RANGE, "negative cost entry");
        max_entry = std::max(max_entry, h->cost[x]);
    }
    for (int x = 0; x < 32; ++x) {
        if (h->prepend[x] < 0 || h->tail[x] < 0) return poy_fail(ctx, POY_ERR_COST_RANGE, "negative prepend/tail cost");
        max_entry = std::max(max_entry, std::max(h->prepend[x], h->tail[x]));
    }
    if (h->gap_open < 0 || max_entry >= POY_INF || h->gap_open >= POY_INF)
        return poy_fail(ctx, POY_ERR_COST_RANGE, "cost entry or gap opening >= HIGH_NUM");
    static std::atomic<uint64_t> next_uid{0};   // contexts may be driven from different host threads
    poy_cm *cm = new poy_cm;
    cm->uid = ++next_uid;
    cm->h = *h;
    cm->min_non0 = poy_cm_min_non0(h);
    cm->max_entry = max_entry;
    memset(&cm->sig, 0, sizeof cm->sig);
    for (int a = 0; a < 32; ++a) { cm->sig.prepend[a] = h->prepend[a]; cm->sig.gapext[a] = h->cost[(a << 5) + GAPC]; }
    cm->sig.gap_open = h->gap_open; cm->sig.valid = 1;
    DevCM *img = new DevCM;
    memset(img, 0, sizeof *img);
    for (int a = 0; a < 16; ++a) for (int b = 0; b < 16; ++b) img->cost16[a * 16 + b] = h->cost[(a << 5) + b];
    memcpy(img->cost32, h->cost, sizeof img->cost32);
    memcpy(img->worst32, h->worst, sizeof img->worst32);
    memcpy(img->median32, h->median, sizeof img->median32);
    for (int a = 1; a < 32; ++a) for (int b = 1; b < 32; ++b) img->closest32[(a << 5) + b] = (uint8_t)poy_cm_get_closest(h, a, b);
    memcpy(img->prepend, h->prepend, sizeof img->prepend);
    memcpy(img->tail, h->tail, sizeof img->tail);
    for (int a = 0; a < 32; ++a) img->gapext[a] = h->cost[(a << 5) + GAPC];
    img->gap_open = h->gap_open;
    img->model = h->cost_model_type;
    img->min_non0 = cm->min_non0;
    img->max_entry = max_entry;
    cudaError_t e = cudaMalloc(&cm->d, sizeof(DevCM));
    if (e == cudaSuccess) e = cudaMemcpyAsync(cm->d, img, sizeof(DevCM), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    delete img;
    if (e != cudaSuccess) { delete cm; return poy_cuda_fail(ctx, e, "poy_cm_upload"); }
    *out = cm;
    return POY_OK;
}
extern "C" void poy_cm_free(poy_ctx *ctx, poy_cm *cm) {
    bind_device(ctx);
    if (!cm) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(cm->d);
    delete cm;
}

// ---- pool ----------------------------------------------------------------------------------------
// `nb` / `ns`: bytes and sequences the per-base / per-sequence arrays are allocated for (the node store asks for more
// than it currently holds)
poy_status pool_alloc(poy_ctx *ctx, poy_pool *p, int64_t nb_, int32_t ns_) {
    const size_t nb = (size_t)std::max<int64_t>(nb_, 1), ns = (size_t)std::max(ns_, 1);
    CK(cached_alloc(ctx, (void **)&p->d_rowp, nb * sizeof(int4), &p->caps[0]));
    CK(cached_alloc(ctx, (void **)&p->d_colp, nb * sizeof(int4), &p->caps[1]));
    CK(cached_alloc(ctx, (void **)&p->d_rowpk, nb * sizeof(unsigned), &p->caps[2]));
    CK(cached_alloc(ctx, (void **)&p->d_h0, nb * sizeof(int), &p->caps[3]));
    CK(cached_alloc(ctx, (void **)&p->d_g0, nb * sizeof(int), &p->caps[4]));
    CK(cached_alloc(ctx, (void **)&p->d_gapfree, ns, &p->caps[5]));
    CK(cached_alloc(ctx, (void **)&p->d_flags, ns * sizeof(int2), &p->caps[8]));
    return POY_OK;
}

extern "C" void poy_pool_free(poy_ctx *ctx, poy_pool *p) {
    bind_device(ctx);
    if (!p) return;
    // no device sync: the blocks go back to the context's cache and are only reused in stream order
    if (p->owns_data) { cached_free(ctx, p->d_data, p->caps[6]); cached_free(ctx, p->d_off, p->caps[7]); }
    cached_free(ctx, p->d_rowp, p->caps[0]); cached_free(ctx, p->d_colp, p->caps[1]); cached_free(ctx, p->d_rowpk, p->caps[2]);
    cached_free(ctx, p->d_h0, p->caps[3]); cached_free(ctx, p->d_g0, p->caps[4]); cached_free(ctx, p->d_gapfree, p->caps[5]);
    cached_free(ctx, p->d_flags, p->caps[8]);
    free(p->h_off);
    free(p->h_gapfree);
    free(p->h_empty);
    free(p->h_gapcnt);
    delete p;
}

// host-side bookkeeping of a pool with room for `cap_seqs` sequences
poy_status pool_new(poy_ctx *ctx, const int64_t *h_off, int32_t nseq, int32_t cap_seqs, poy_pool **out) {
    if (nseq < 0 || !h_off || h_off[0] != 0) return poy_fail(ctx, POY_ERR_ARG, "pool offsets must start at 0");
    for (int s = 0; s < nseq; ++s)
        if (h_off[s + 1] <= h_off[s]) return poy_fail(ctx, POY_ERR_ARG, "every pool sequence needs at least its leading gap");
    if (cap_seqs < nseq) cap_seqs = nseq;
    poy_pool *p = new poy_pool;
    memset(p, 0, sizeof *p);
    p->nseq = nseq;
    p->nbytes = h_off[nseq];
    p->h_gapfree = (uint8_t *)calloc((size_t)cap_seqs + 1, 1);
    p->h_empty = (uint8_t *)calloc((size_t)cap_seqs + 1, 1);
    p->h_gapcnt = (int32_t *)calloc((size_t)cap_seqs + 1, sizeof(int32_t));
    p->h_off = (int64_t *)malloc(sizeof(int64_t) * ((size_t)cap_seqs + 1));
    if (!p->h_gapfree || !p->h_empty || !p->h_gapcnt || !p->h_off) { poy_pool_free(nullptr, p); return poy_fail(ctx, POY_ERR_NOMEM, "host allocation failed"); }
    memcpy(p->h_off, h_off, sizeof(int64_t) * (nseq + 1));
    *out = p;
    return POY_OK;
}

extern "C" poy_status poy_pool_upload(poy_ctx *ctx, const uint8_t *data, const int64_t *offsets, int32_t nseq, poy_pool **out) {
    bind_device(ctx);
    if (!ctx || !data || !offsets || !out) return POY_ERR_ARG;
    *out = nullptr;
    poy_pool *p;
    poy_status s = pool_new(ctx, offsets, nseq, nseq, &p);
    if (s != POY_OK) return s;
    p->owns_data = true;
    cudaError_t e = cached_alloc(ctx, (void **)&p->d_data, (size_t)std::max<int64_t>(p->nbytes, 1), &p->caps[6]);
    if (e == cudaSuccess) e = cached_alloc(ctx, (void **)&p->d_off, sizeof(int64_t) * (nseq + 1), &p->caps[7]);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_data, data, (size_t)p->nbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_off, offsets, sizeof(int64_t) * (nseq + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { poy_pool_free(ctx, p); return poy_cuda_fail(ctx, e, "poy_pool_upload"); }
    s = pool_alloc(ctx, p, p->nbytes, nseq);
    if (s != POY_OK) { poy_pool_free(ctx, p); return s; }
    *out = p;
    return POY_OK;
}

extern "C" poy_status poy_pool_from_device(poy_ctx *ctx, const uint8_t *d_data, const int64_t *d_offsets,
                                           const int64_t *h_offsets, int32_t nseq, poy_pool **out) {
    bind_device(ctx);
    if (!ctx || !d_data || !d_offsets || !h_offsets || !out) return POY_ERR_ARG;
    *out = nullptr;
    poy_pool *p;
    poy_status s = pool_new(ctx, h_offsets, nseq, nseq, &p);
    if (s != POY_OK) return s;
    p->owns_data = false;
    p->d_data = const_cast<uint8_t *>(d_data);
    p->d_off = const_cast<int64_t *>(d_offsets);
    s = pool_alloc(ctx, p, p->nbytes, nseq);
    if (s != POY_OK) { poy_pool_free(ctx, p); return s; }
    *out = p;
    return POY_OK;
}

// Per-base gap parameters of the sequences that do not have them yet for this cost model (all of them when the
// model's parameter signature differs from the one they were computed for).
poy_status ensure_params(poy_ctx *ctx, const poy_cm *cm, const poy_pool *cpool) {
    poy_pool *pool = const_cast<poy_pool *>(cpool);
    if (!pool->sig.valid || memcmp(&pool->sig, &cm->sig, sizeof(ParamSig)) != 0) { pool->sig = cm->sig; pool->params_upto = 0; }
    if (pool->params_upto >= pool->nseq) return POY_OK;
    const int s0 = pool->params_upto, s1 = pool->nseq;
    CK(launch_params(ctx, cm, pool, s0, s1));
    CK(cudaMemcpyAsync(pool->h_gapfree + s0, pool->d_gapfree + s0, (size_t)(s1 - s0), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    pool->params_upto = s1;
    return POY_OK;
}

// Sequence.is_empty / Sequence.count_gaps of the sequences that do not have them yet (device kernel + 8 bytes per
// sequence read back; pools made by poy_pool_from_device and the node store's medians have no host copy)
poy_status ensure_flags(poy_ctx *ctx, const poy_pool *cpool) {
    poy_pool *pool = const_cast<poy_pool *>(cpool);
    if (pool->flags_upto >= pool->nseq) return POY_OK;
    bind_device(ctx);
    const int s0 = pool->flags_upto, s1 = pool->nseq, n = s1 - s0;
    void *v_pin;
    poy_status s = poy_pinned(ctx, 4, sizeof(int2) * (size_t)n, &v_pin);
    if (s != POY_OK) return s;
    int2 *hf = (int2 *)v_pin;
    CK(launch_seq_flags(ctx, pool, s0, s1, (int2 *)pool->d_flags + s0));
    CK(cudaMemcpyAsync(hf, (int2 *)pool->d_flags + s0, sizeof(int2) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < n; ++q) { pool->h_empty[s0 + q] = (uint8_t)hf[q].x; pool->h_gapcnt[s0 + q] = hf[q].y; }
    pool->flags_upto = s1;
    return POY_OK;
}

static int64_t max_len(const poy_pool *pool) {
    int64_t m = 0;
    for (int s = 0; s < pool->nseq; ++s) m = std::max(m, pool->h_off[s + 1] - pool->h_off[s]);
    return m;
}

// The reference is only defined while path costs stay below HIGH_NUM (src/algn.c:37).
static poy_status domain_check(poy_ctx *ctx, const poy_cm *cm, int64_t len_sum) {
    const int64_t per_step = (int64_t)cm->max_entry + 2 * (int64_t)cm->h.gap_open;
    if (per_step * len_sum + cm->h.gap_open >= POY_INF)
        return poy_fail(ctx, POY_ERR_COST_RANGE, "sequence lengths x costs can reach HIGH_NUM; the reference is undefined there");
    return POY_OK;
}

// ---- batch cost-only affine --------------------------------------------------------------------------
extern "C" poy_status poy_batch_cost_affine_dev(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                                const int32_t *d_a, const int32_t *d_b, int32_t *d_cost) {
    bind_device(ctx);
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!d_a || !d_b || !d_cost))) return POY_ERR_ARG;
    if (cm->h.cost_model_type != 1) return poy_fail(ctx, POY_ERR_MODEL, "cost_affine needs an affine cost model");
    if (n == 0) return POY_OK;
    // `ml`: longest sequence among the submitted pairs when the caller knows it (host entry point), else of the pool
    const int64_t ml = ctx->hint_max_len > 0 ? ctx->hint_max_len : max_len(pool);
    ctx->hint_max_len = 0;
    poy_status s = domain_check(ctx, cm, 2 * ml);
    if (s != POY_OK) return s;
    s = ensure_params(ctx, cm, pool);
    if (s != POY_OK) return s;
    void *jobs, *misc, *bound;
    if ((s = poy_scratch(ctx, SL_JOBS, sizeof(CostJob) * (size_t)n, &jobs)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_MISC, 64, &misc)) != POY_OK) return s;
    const int blocks = ctx->sm_count * 4;
    const size_t bound_stride = (size_t)ml + 2;
    if ((s = poy_scratch(ctx, SL_BOUND, sizeof(int4) * 2 * bound_stride * (size_t)blocks * 4, &bound)) != POY_OK) return s;
    int *counts = (int *)misc;  // [0],[1] = work counters; [2],[3] = job counts
    CK(cudaMemsetAsync(counts, 0, 16, ctx->stream));
    CK(launch_build_cost_jobs(ctx, pool, n, d_a, d_b, (CostJob *)jobs, counts + 2));
    CK(launch_cost_affine(ctx, cm, pool, (CostJob *)jobs, n, counts, (int4 *)bound, bound_stride, blocks, d_cost, ml >= 1500));
    return POY_OK;
}

extern "C" poy_status poy_batch_cost_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                            const int32_t *a, const int32_t *b, int32_t *cost) {
    bind_device(ctx);
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!a || !b || !cost))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    int64_t ml = 1;
    for (int p = 0; p < n; ++p) {
        if (a[p] < 0 || a[p] >= pool->nseq || b[p] < 0 || b[p] >= pool->nseq) return poy_fail(ctx, POY_ERR_ARG, "pair index out of range");
        ml = std::max(ml, std::max(pool->h_off[a[p] + 1] - pool->h_off[a[p]], pool->h_off[b[p] + 1] - pool->h_off[b[p]]));
    }
    void *dbuf;
    poy_status s = poy_scratch(ctx, SL_STATE, sizeof(int32_t) * 3 * (size_t)n, &dbuf);
    if (s != POY_OK) return s;
    int32_t *d_a = (int32_t *)dbuf, *d_b = d_a + n, *d_cost = d_b + n;
    CK(cudaMemcpyAsync(d_a, a, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_b, b, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->hint_max_len = ml;    // boundary scratch and domain check from the submitted pairs, not from the whole pool
    s = poy_batch_cost_affine_dev(ctx, cm, pool, n, d_a, d_b, d_cost);
    if (s != POY_OK) return s;
    CK(cudaMemcpyAsync(cost, d_cost, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

// ---- batch banded affine with traceback ------------------------------------------------------------------
namespace {
struct HostPair {
    int lasti, lastj, T, k, dclass, stride, gapfree, fullplane;
    int probe;      // this round's fill writes no direction bytes (see the round loop of align_impl)
    int want_dirs;  // the previous verdict asked for a fill with direction bytes
    int repeat;     // ... at the same threshold (not a new iteration of the reference's loop)
    int64_t off_i, off_j, eb_off, dir_bytes, work;
    int iterations;
    int64_t cells;
    bool done;
    bool eh00_inf;     // some earlier fill of this pair had min(k, lasti) >= 2: it left EH[0][0] = INF behind (k_band2)
    bool state_dirty;  // speculative rounds moved T / EH[0][0] on without the device copy of the pair's state
};

// number of band cells of one fill (rows 1..lasti, columns max(i-k,0)..min(i+delta+k,lastj))
int64_t band_cells(int lasti, int lastj, int k) {
    if (lasti <= 0) return 0;
    const int64_t delta = lastj - lasti;
    int64_t hi = 0, lo = 0;
    const int64_t n1 = std::max<int64_t>(0, std::min<int64_t>(lasti, (int64_t)lastj - 1 - delta - k));  // rows with i+delta+k <= lastj-1
    hi += n1 * (delta + k) + n1 * (n1 + 1) / 2;
    hi += (lasti - n1) * (int64_t)lastj;
    if (lasti > k) { const int64_t m = lasti - k; lo = m * (m + 1) / 2; }
    return hi - lo + lasti;
}
}  // namespace

extern "C" poy_status poy_batch_align_affine_dev(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                                 const int32_t *d_si, const int32_t *d_sj, const uint8_t *d_swaped,
                                                 const int32_t *h_si, const int32_t *h_sj, const int64_t *d_out_off,
                                                 int32_t *d_cost, uint8_t *d_median, uint8_t *d_medianwg,
                                                 uint8_t *d_resi, uint8_t *d_resj, int32_t *d_out_len, int32_t *d_stats);

// `h_deltawh` != NULL selects the linear-gap path (algn_CAML_simple_2 / align_2d): d_resi / d_resj then receive the
// two aligned rows (capacity len1+len2 each, src/sequence.ml:1019-1033), d_median / d_medianwg are unused and
// d_out_len has 2 entries per pair.
static poy_status align_impl(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n, const int32_t *h_si,
                             const int32_t *h_sj, const uint8_t *h_swaped, const int64_t *d_out_off, int32_t *d_cost,
                             uint8_t *d_median, uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj,
                             int32_t *d_out_len, int32_t *h_stats, const int32_t *h_deltawh = nullptr) {
    const bool linear = h_deltawh != nullptr;
    if (!linear && cm->h.cost_model_type != 1) return poy_fail(ctx, POY_ERR_MODEL, "align_affine needs an affine cost model");
    if (linear && cm->h.cost_model_type == 1) return poy_fail(ctx, POY_ERR_MODEL, "the linear-gap entry points need a non-affine cost model");
    if (n == 0) return POY_OK;
    const bool want_trace = d_median || d_medianwg || d_resi || d_resj || d_out_len;
    if (want_trace && !d_out_off) return poy_fail(ctx, POY_ERR_ARG, "out_off is required when any traceback output is requested");
    // Speculative threshold doublings (latency-bound rounds only, see the round loop): up to SPEC_EXTRA extra fills per
    // round, each a clone of its pair at a later threshold, appended to hp behind the n real pairs.  Every per-pair
    // array below is sized for n + SPEC_EXTRA entries.
    const int SPEC_EXTRA = 1024;
    const size_t nx = (size_t)n + SPEC_EXTRA;
    std::vector<HostPair> hp;
    hp.reserve(nx);
    hp.resize((size_t)n);
    int64_t eb_total = 0, maxsum = 0;
    for (int p = 0; p < n; ++p) {
        const int a = h_si[p], b = h_sj[p];
        if (a < 0 || a >= pool->nseq || b < 0 || b >= pool->nseq) return poy_fail(ctx, POY_ERR_ARG, "pair index out of range");
        HostPair &h = hp[p];
        h.off_i = pool->h_off[a]; h.off_j = pool->h_off[b];
        h.lasti = (int)(pool->h_off[a + 1] - h.off_i) - 1;
        h.lastj = (int)(pool->h_off[b + 1] - h.off_j) - 1;
        if (h.lastj < h.lasti) return poy_fail(ctx, POY_ERR_ORDER, "pass the shorter one as first");
        h.T = (h.lastj - h.lasti + 1) * cm->min_non0;  // algn_fill_plane_3_aff, src/algn.c:2348-2349
        h.iterations = 0; h.cells = 0; h.done = false; h.fullplane = 0; h.probe = 0; h.want_dirs = 0; h.repeat = 0;
        h.eh00_inf = false; h.state_dirty = false;
        if (linear) {   // algn_nw_limit / algn_fill_plane_2: full plane or Ukkonen band (src/algn.c:2963, 1141-1176)
            const int lenX = h.lasti + 1, lenY = h.lastj + 1;
            int height = (lenX - lenY) + 50 + h_deltawh[p];
            if (height > lenX) height = lenX;
            if ((float)lenX >= 1.5f * (float)lenY) h.fullplane = 1;
            else if (!((2 * height) < lenX) && 8 >= (lenX - height)) h.fullplane = 1;
        }
        maxsum = std::max<int64_t>(maxsum, (int64_t)h.lasti + h.lastj + 2);
    }
    poy_status s = domain_check(ctx, cm, maxsum);
    if (s != POY_OK) return s;
    if ((s = ensure_params(ctx, cm, pool)) != POY_OK) return s;
    // POY_FORCE_GENERIC=1 routes every pair through the any-width fallback kernels (test hook)
    const char *fg = getenv("POY_FORCE_GENERIC");
    const bool force_generic = fg && fg[0] == '1';
    for (int p = 0; p < n; ++p) {
        hp[p].gapfree = linear ? 1 : (pool->h_gapfree[h_si[p]] && pool->h_gapfree[h_sj[p]]);
        // the stale EB row exists only for pairs with gap-bit symbols (and for the fallback kernel, which runs
        // every pair through the 4-state code)
        const bool generic = force_generic || (int64_t)hp[p].lasti + hp[p].lastj + 2 >= 65535;
        hp[p].eb_off = eb_total;
        if (!linear && (!hp[p].gapfree || generic)) eb_total += hp[p].lastj + 1;
    }
    // Threshold doublings that provably cannot stop are skipped.  Every path from (0,0) to (leni,lenj) has
    // (#insertions - #deletions) = delta, and the gap counters are maxima over tie paths, so gap_num >= delta;
    // the stop rule (affine: gap_num < p, linear: gap_num + 1 < p) therefore fails while p <= delta (+1), and
    // unless the "band spans the matrix" clause fires the reference just doubles T.  A skipped fill leaves no
    // trace in the result -- except through the stale EB row / EH[0][0] of the affine path, which only pairs
    // with gap-bit symbols can observe, so those pairs run every fill.
    for (int p = 0; p < n; ++p) {
        HostPair &h = hp[p];
        if (h.lasti == 0 || h.fullplane || (!linear && !h.gapfree)) continue;
        const int delta = h.lastj - h.lasti;
        for (;;) {
            const int pp = (h.T - delta) / 2, newp = (2 * h.T - delta) / 2;
            const bool spans = linear ? (newp - (h.lastj + 1) + 1 >= 0) : (newp - h.lastj + 1 >= 0);
            const bool cannot_stop = linear ? (pp <= delta + 1) : (pp <= delta);
            if (spans || !cannot_stop || h.T > (1 << 28)) break;
            h.T *= 2;
            h.iterations++;
        }
    }

    void *v_state, *v_eb, *v_jobs, *v_jobs_b, *v_misc, *v_pin, *v_pin2;
    const size_t nx4 = (nx + 3) & ~(size_t)3;
    if ((s = poy_scratch(ctx, SL_STATE, sizeof(PairState) * nx + 3 * nx4 + sizeof(int) * nx, &v_state)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_EBROW, sizeof(int) * 2 * (size_t)eb_total + 16, &v_eb)) != POY_OK) return s;   // rows + their snapshots
    if ((s = poy_scratch(ctx, SL_JOBS, sizeof(BandJob) * (nx + SPEC_EXTRA), &v_jobs)) != POY_OK) return s;     // fills + the traceback jobs of a speculative round
    if ((s = poy_scratch(ctx, SL_JOBSB, sizeof(BandJob) * (nx + SPEC_EXTRA), &v_jobs_b)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_MISC, 256, &v_misc)) != POY_OK) return s;
    if ((s = poy_pinned(ctx, 0, std::max(sizeof(BandJob), sizeof(PairState)) * (nx + SPEC_EXTRA), &v_pin)) != POY_OK) return s;
    if ((s = poy_pinned(ctx, 1, nx + sizeof(int32_t) * nx, &v_pin2)) != POY_OK) return s;
    PairState *d_state = (PairState *)v_state;
    uint8_t *d_done = (uint8_t *)(d_state + nx);
    // the verdicts as the traceback of a round must see them: the next round's stop rule overwrites d_done while the
    // traceback of this one may still be running
    uint8_t *d_done_tb[2] = { d_done + nx4, d_done + 2 * nx4 };
    int *d_prog = (int *)(d_done + 3 * nx4);     // progress words of speculative fills, one per hp entry
    int *d_eb = (int *)v_eb, *d_eb_snap = d_eb + eb_total;
    BandJob *d_jobs_ab[2] = { (BandJob *)v_jobs, (BandJob *)v_jobs_b };
    int *d_counter = (int *)v_misc;
    // POY_ASYNC_TB=0: the traceback of a round is waited for before the next round starts (test hook)
    const char *atb = getenv("POY_ASYNC_TB");
    const bool async_tb = !(atb && atb[0] == '0');
    uint8_t *h_done = (uint8_t *)v_pin2;

    {   // initial per-pair state: EH[0][0] = go, EB row 0 = INF (initialize_matrices_affine, src/algn.c:1875-1898)
        PairState *hs = (PairState *)v_pin;
        for (int p = 0; p < n; ++p) { hs[p].T = hp[p].T; hs[p].eh00 = cm->h.gap_open; hs[p].cost = 0; hs[p].gapnum = 0; hs[p].iterations = 0; hs[p].done = 0; hs[p].cells = 0; }
        CK(cudaMemcpyAsync(d_state, hs, sizeof(PairState) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(d_done, 0, nx, ctx->stream));
    }
    CK(launch_fill_int(ctx, d_eb, eb_total, POY_INF));
    CK(cudaStreamSynchronize(ctx->stream));  // pinned staging is reused below

    std::vector<int> active((size_t)n);
    for (int p = 0; p < n; ++p) active[p] = p;
    std::vector<int> order, rjobs;
    std::vector<uint64_t> keys;
    struct Clone { int real, prev; };   // the pair a speculative fill belongs to, the hp entry of the previous doubling
    std::vector<Clone> clones;
    const int gen_blocks = ctx->sm_count * 2;

    const char *tr = getenv("POY_TRACE");
    const bool trace = tr && (tr[0] == '1' || tr[0] == '2');
    const bool trace_rounds = tr && tr[0] == '2';   // POY_TRACE=2: one line per threshold-doubling round
    double t_prep = 0, t_wait = 0; int rounds = 0, waves = 0;
    long long n_repeat = 0, n_probe = 0, n_full = 0;
    const char *pe = getenv("POY_PROBE");   // POY_PROBE=0 turns the probe fills off, 2 forces them on small batches too (test hook)
    // small batches are latency bound (a repeated threshold is one more round trip), probes only pay when the
    // fills saturate the GPU
    // POY_PROBE_MIN / POY_PROBE_MIN4: smallest batch in which gap-free / 4-state pairs are probed (tuning hooks)
    const char *pm = getenv("POY_PROBE_MIN"), *pm4 = getenv("POY_PROBE_MIN4");
    const int probe_min = pm ? atoi(pm) : 1024, probe_min4 = pm4 ? atoi(pm4) : 1024;
    const bool forced = pe && pe[0] == '2';
    const bool use_probes = !(pe && pe[0] == '0') && (n >= std::min(probe_min, probe_min4) || forced);
    const bool probe_gf = forced || n >= probe_min, probe_4s = forced || n >= probe_min4;
    const char *spe = getenv("POY_SPEC");   // POY_SPEC=0: one fill per pair and round, always
    const bool spec_allowed = !linear && !use_probes && !force_generic && !(spe && spe[0] == '0');
    const char *spt = getenv("POY_SPEC_TEST");
    const bool spec_test = spt && spt[0] == '1';
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = now();
    // POY_LOWLAT=0 turns the low-latency kernel shapes off, 2 forces them (test hook); default: rounds with at most
    // two pairs per SM, where the round lasts as long as one pair's wavefront (see launch_band2)
    const char *ll = getenv("POY_LOWLAT");
    while (!active.empty()) {
        ++rounds;
        // band geometry of this round (algn_newkk_increaseT_aff / algn_newkk_test_aff, src/algn.c:2311-2336, 2195-2196)
        auto geometry = [&](HostPair &h) {
            const int delta = h.lastj - h.lasti;
            const int pp = (h.T - delta) / 2;
            if (!linear) h.k = pp >= h.lasti ? h.lasti - 1 : pp;           // src/algn.c:2195-2196 (lenX = last index)
            else h.k = h.fullplane ? h.lasti : (pp >= h.lasti + 1 ? h.lasti : pp);  // :1070-1071 (lenX = length)
            if (h.lasti == 0) h.k = 0;
            const int64_t B = (int64_t)delta + 2 * (int64_t)h.k + 1;
            // the packed 16x2 gap counters of k_band2 are exact while len_i + len_j < 65535
            h.dclass = h.lasti == 0 ? 64 : ((int64_t)h.lasti + h.lastj + 2 >= 65535 || force_generic) ? 0 : band2_class_for(B);
            h.stride = h.dclass ? band2_stride_for(h.dclass, B) : (int)(((B + 1) / 2 + 31) & ~31ll);
            h.work = h.lasti == 0 ? 64 : ((int64_t)h.lasti + h.lastj + 2) * h.stride;
            // Probe fills.  A fill that will not be the last one needs no direction bytes, and without them a cell
            // costs less than half.  Gap-free pairs are therefore filled without directions until a verdict says
            // "the stop rule fired" (2: repeat this threshold with directions) or "the next fill should stop" (3).
            // For pairs with gap-bit symbols a probe first snapshots the stale EB row / EH[0][0] it is about to
            // change, and the repeated fill starts from the snapshot.  Without traceback outputs nobody needs
            // directions at all.
            h.probe = 0;
            if (!linear && h.dclass != 0 && h.lasti != 0 && use_probes && (h.gapfree ? probe_gf : probe_4s)) {
                if (!want_trace) h.probe = 1;
                else if (!h.want_dirs) {
                    const int newp = (2 * h.T - delta) / 2;
                    h.probe = !(newp - h.lastj + 1 >= 0);   // a band that spans the matrix always stops
                }
            }
            h.dir_bytes = h.probe ? 0 : h.work;
        };
        hp.resize((size_t)n);            // (drops the clones of the previous round)
        for (int p : active) {
            HostPair &h = hp[p];
            geometry(h);
            if (h.probe) ++n_probe; else ++n_full;
            if (!h.repeat) h.iterations++;
            h.cells += h.fullplane ? (int64_t)h.lasti * (h.lastj + 1) : band_cells(h.lasti, h.lastj, h.k);
        }
        // A round is latency bound when all its fills are resident at once: then it lasts as long as ONE pair's wavefront
        // and the shapes with the fewest diagonals per thread are the right ones.  Resident warps: every shape keeps at
        // least 16 warps on an SM (launch bounds of k_band2), a fill takes 1 ... 8 warps per CTA, cluster shapes 2 or 4 CTAs.
        // (2-warp CTAs are bound by shared memory -- six per SM -- and count as 3)
        auto warps_of = [](int cls) { return cls <= 64 ? 1 : cls <= 128 ? 3 : cls <= 256 ? 4 : cls <= 1024 ? 8 : cls <= 2048 ? 16 : 32; };
        const int warp_capacity = ctx->sm_count * 16 - 256;
        int64_t warp_demand = 0;
        for (int p : active) warp_demand += warps_of(hp[p].dclass);
        // (the demand rule only for batches that may speculate -- small, probe-free batches, i.e. tree passes: there it is
        // measured; the tail rounds of large batches keep the two-pairs-per-SM rule they were tuned with)
        const bool lowlat = !(ll && ll[0] == '0') && ((ll && ll[0] == '2') || (int64_t)active.size() <= 2 * (int64_t)ctx->sm_count ||
                                                      (spec_allowed && warp_demand <= warp_capacity && (int)active.size() <= SPEC_EXTRA));
        // Speculative doublings.  In a latency-bound round most SMs idle while every pair waits for ONE wavefront, and a
        // pair then goes through 4-8 such rounds.  So the next doublings of each pair are filled in the same round, as
        // clones of the pair at 2T, 4T, ...: each fill needs from its predecessor only the stale EB entries of columns
        // 0 .. delta + k (written by the predecessor's two leftmost diagonals within its first delta + k + k' rows) and
        // EH[0][0] (known beforehand), so it starts a few rows behind it and stays behind (k_band2: progress words).
        // The first fill of the chain whose stop rule fires is the result, later ones are discarded; the stale row is
        // shared and written in the order of the doublings, so whatever comes next starts from the same state as in
        // the one-fill-per-round schedule.  The chain ends at a fill that certainly stops (its doubled band spans the
        // matrix), when the CTAs no longer fit the GPU at once (the fills wait for each other: all must be resident),
        // or when the direction bytes no longer fit the arena.  POY_SPEC=0 turns this off.
        clones.clear();
        bool spec_round = false;
        if (spec_allowed && lowlat) {
            // every fill of the round must be resident at once (they wait for each other): the budget is counted in
            // resident warps; POY_SPEC_CTAS overrides it, in CTAs of 8 warps (tuning / test hook)
            const char *sb = getenv("POY_SPEC_CTAS");
            int64_t budget = (sb ? (int64_t)atoi(sb) * 8 : (int64_t)warp_capacity) - warp_demand;
            int64_t arena_used = 0;
            bool ok = true;
            for (int p : active) arena_used += (hp[p].dir_bytes + 255) & ~255ll;
            if (budget < 0 || (uint64_t)arena_used * 4 > ctx->arena_limit || (int)active.size() > SPEC_EXTRA) ok = false;
            std::vector<int> tail(active.begin(), active.end());
            std::vector<char> open(active.size(), 1);
            for (int lvl = 1; ok && lvl < 12; ++lvl) {   // (a 10 kb pair of distant sequences takes up to ~12 doublings)
                bool any = false;
                for (size_t qi = 0; qi < active.size(); ++qi) {
                    if (!open[qi]) continue;
                    open[qi] = 0;
                    const HostPair prev = hp[tail[qi]];
                    const int delta = prev.lastj - prev.lasti;
                    if (prev.lasti == 0 || prev.dclass == 0 || prev.T > (1 << 27)) continue;
                    if ((2 * prev.T - delta) / 2 - prev.lastj + 1 >= 0) continue;      // prev always stops
                    HostPair c = prev;
                    c.T = prev.T * 2; c.want_dirs = 0; c.repeat = 0;
                    geometry(c);
                    if (c.dclass == 0) continue;
                    const int64_t bytes = (c.dir_bytes + 255) & ~255ll;
                    if (warps_of(c.dclass) > budget || (uint64_t)(arena_used + bytes) * 4 > ctx->arena_limit ||
                        (uint64_t)(arena_used + bytes) > (1ull << 30) || (int)clones.size() >= SPEC_EXTRA) continue;
                    budget -= warps_of(c.dclass); arena_used += bytes;
                    clones.push_back({ active[qi], tail[qi] });
                    hp.push_back(c);
                    tail[qi] = (int)hp.size() - 1;
                    open[qi] = 1; any = true;
                }
                if (!any) break;
            }
            spec_round = !clones.empty();
        }
        if (!spec_round)
            for (int p : active)
                if (hp[p].state_dirty) {     // back to one fill per round: the device copy of T / EH[0][0] catches up
                    const int v[2] = { hp[p].T, hp[p].eh00_inf ? POY_INF : cm->h.gap_open };
                    CK(cudaMemcpyAsync(&d_state[p].T, v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
                    hp[p].state_dirty = false;
                }
        rjobs.assign(active.begin(), active.end());
        for (size_t c = 0; c < clones.size(); ++c) rjobs.push_back(n + (int)c);
        n_full += (long long)clones.size();
        // order: by kernel class, 4-state pairs before gap-free ones, then by size (largest first, in 1 KiB steps) for
        // load balance; packed into one integer key so the sort touches no other memory
        // (speculative rounds: narrow classes first, the wider fills of a chain wait for them)
        keys.resize(rjobs.size());
        for (size_t q = 0; q < rjobs.size(); ++q) {
            const HostPair &h = hp[rjobs[q]];
            const uint64_t size_q = (uint64_t)std::min<int64_t>(h.work >> 10, (1 << 17) - 1);
            keys[q] = ((uint64_t)(spec_round ? h.dclass : 4096 - h.dclass) << 51) | ((uint64_t)(h.gapfree ? 1 : 0) << 50) | ((uint64_t)(h.probe ? 1 : 0) << 49) |
                      ((((uint64_t)1 << 17) - 1 - size_q) << 32) | (uint32_t)rjobs[q];
        }
        std::sort(keys.begin(), keys.end());
        order.resize(rjobs.size());
        for (size_t q = 0; q < rjobs.size(); ++q) order[q] = (int)(uint32_t)keys[q];
        size_t pos = 0;
        int w_par = 0; bool w_single = false; BandJob *w_jobs = nullptr; uint8_t *w_dir = nullptr;   // the last wave (speculative rounds have one)
        while (pos < order.size()) {
            // one wave: as many pairs as fit in the direction arena
            int64_t used = 0;
            size_t end = pos;
            while (end < order.size()) {
                const int64_t need = (hp[order[end]].dir_bytes + 255) & ~255ll;
                if (end > pos && used + need > (int64_t)ctx->arena_limit) break;
                used += need;
                ++end;
            }
            // Rounds that fit the arena in one wave alternate between two arenas / job arrays, so that the traceback
            // of the pairs that stop in this round (own stream) overlaps the fills of the next one.
            // (the second arena stays small -- at most 1 GiB and a quarter of the limit -- so that a context never outgrows
            // what it was allowed before: large rounds are throughput bound and gain nothing from the overlap)
            const bool single_wave = async_tb && pos == 0 && end == order.size() && (uint64_t)used * 4 <= ctx->arena_limit &&
                                     (uint64_t)used <= (1ull << 30);
            const int par = single_wave ? (rounds & 1) : 0;
            if (!single_wave) {
                for (int q = 0; q < 2; ++q)
                    if (ctx->tb_pending[q]) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tb_done[q], 0)); ctx->tb_pending[q] = false; }
            } else if (ctx->tb_pending[par]) {
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tb_done[par], 0)); ctx->tb_pending[par] = false;
            }
            BandJob *d_jobs = d_jobs_ab[par];
            void *v_dir;
            if ((s = poy_scratch(ctx, par ? SL_DIR2 : SL_DIR, (size_t)used, &v_dir)) != POY_OK) return s;
            uint8_t *d_dir = (uint8_t *)v_dir;
            BandJob *hj = (BandJob *)v_pin;
            const int nj = (int)(end - pos);
            int64_t doff = 0;
            int64_t gen_width = 0;
            for (int q = 0; q < nj; ++q) {
                const int p = order[pos + q];
                const HostPair &h = hp[p];
                BandJob &j = hj[q];
                j.off_i = h.off_i; j.off_j = h.off_j; j.lasti = h.lasti; j.lastj = h.lastj; j.k = h.k; j.pair = p;
                j.swaped = (h_swaped ? (h_swaped[p < n ? p : clones[p - n].real] ? 1 : 0) : 0) | (h.fullplane ? 2 : 0) | ((!linear && h.gapfree) ? 4 : 0) |
                           (h.probe ? 8 : 0) | (want_trace ? 0 : 16) | (h.repeat ? 32 : 0) |
                           ((!linear && h.dclass != 0) ? 64 : 0);
                j.stride = h.stride; j.dir_off = doff; j.eb_off = h.eb_off;
                j.dep = -1; j.need = 0; j.T = h.T; j.eh00 = 0;
                if (spec_round && h.dclass != 0 && h.lasti != 0) {
                    j.swaped |= 128;
                    // EH[0][0] as the doublings before this one leave it
                    bool inf = hp[p < n ? p : clones[p - n].real].eh00_inf;
                    for (int x = p; x >= n; ) { x = clones[x - n].prev; if (std::min(hp[x].k, hp[x].lasti) >= 2) inf = true; }
                    j.eh00 = inf ? POY_INF : cm->h.gap_open;
                    if (p >= n && !h.gapfree) {
                        const HostPair &pv = hp[clones[p - n].prev];
                        j.dep = clones[p - n].prev;
                        j.need = (h.lastj - h.lasti) + h.k + pv.k + 2;
                        if (j.need + 20 >= h.lasti) j.need = SPEC_DONE;     // (the progress word moves in steps of 16 rows)
                        // POY_SPEC_TEST=1 (test hook): every third speculative fill waits for a row its predecessor never
                        // reports, gives up after the bounded wait and is run again in the next round
                        if (spec_test && (p - n) % 3 == 1) j.need = SPEC_ABORT + 1;
                    }
                }
                doff += (h.dir_bytes + 255) & ~255ll;
                if (h.dclass == 0) gen_width = std::max<int64_t>(gen_width, (int64_t)(h.lastj - h.lasti) + 2 * h.k + 1);
            }
            CK(cudaMemcpyAsync(d_jobs, hj, sizeof(BandJob) * (size_t)nj, cudaMemcpyHostToDevice, ctx->stream));
            if (spec_round) CK(cudaMemsetAsync(d_prog, 0, sizeof(int) * nx, ctx->stream));
            if (!linear && eb_total > 0 && want_trace && use_probes) CK(launch_stale_snapshot(ctx, d_jobs, nj, d_state, d_eb, d_eb_snap));
            // launch per band class (contiguous after the sort).  The launches of a wave are independent: they go to
            // round-robin auxiliary streams (fork / join with events) and each gets its own work counter.
            int q0 = 0, nlaunch = 0;
            cudaStream_t main_stream = ctx->stream;
            CK(cudaEventRecord(ctx->ev_fork, main_stream));
            unsigned used_aux = 0;
            while (q0 < nj) {
                const int ax = nlaunch & (POY_N_AUX - 1);
                ctx->stream = ctx->aux[ax];
                if (!(used_aux & (1u << ax))) { cudaStreamWaitEvent(ctx->stream, ctx->ev_fork, 0); used_aux |= 1u << ax; }
                const int cls = hp[order[pos + q0]].dclass, gf = hp[order[pos + q0]].gapfree, pr = hp[order[pos + q0]].probe;
                int q1 = q0;
                // (the fallback kernels take both kinds of pair in one launch: they share one work buffer)
                while (q1 < nj && hp[order[pos + q1]].dclass == cls &&
                       (linear || cls == 0 || (hp[order[pos + q1]].gapfree == gf && hp[order[pos + q1]].probe == pr))) ++q1;
                if (cls != 0) {
                    cudaError_t le;
                    if (linear) le = launch_band_lin(ctx, cm, pool, d_jobs + q0, q1 - q0, cls, d_counter + (nlaunch & 15), d_state, d_dir);
                    else le = launch_band2(ctx, cm, pool, d_jobs + q0, q1 - q0, cls, gf != 0, pr != 0, d_counter + (nlaunch & 15), d_state, d_eb, d_dir, lowlat, d_prog);
                    if (le != cudaSuccess) { ctx->stream = main_stream; return poy_cuda_fail(ctx, le, "band fill launch"); }
                } else {
                    void *v_work;
                    const size_t wstride = 6 * (size_t)((gen_width + 31) & ~31ll);
                    const int blocks = std::min(gen_blocks, q1 - q0);
                    ctx->stream = main_stream;
                    if ((s = poy_scratch(ctx, SL_WORK, sizeof(int) * wstride * (size_t)blocks, &v_work)) != POY_OK) return s;
                    ctx->stream = ctx->aux[ax];
                    cudaError_t le;
                    if (linear) le = launch_band_lin_generic(ctx, cm, pool, d_jobs + q0, q1 - q0, d_state, d_dir, (int *)v_work, wstride, blocks);
                    else le = launch_band_generic(ctx, cm, pool, d_jobs + q0, q1 - q0, d_state, d_eb, d_dir, (int *)v_work, wstride, blocks);
                    if (le != cudaSuccess) { ctx->stream = main_stream; return poy_cuda_fail(ctx, le, "generic band fill launch"); }
                }
                ++nlaunch;
                q0 = q1;
            }
            ctx->stream = main_stream;
            for (int ax = 0; ax < POY_N_AUX; ++ax)
                if (used_aux & (1u << ax)) {
                    CK(cudaEventRecord(ctx->ev_join[ax], ctx->aux[ax]));
                    CK(cudaStreamWaitEvent(main_stream, ctx->ev_join[ax], 0));
                }
            // stop rule on the device, then traceback of the pairs that stopped (others return immediately)
            if (linear) CK(launch_lin_finish(ctx, pool, d_jobs, nj, d_state, d_done));
            else CK(launch_band_finish(ctx, d_jobs, nj, d_state, d_done, pool->d_g0, cm->h.gap_open, d_prog));
            if (want_trace && !spec_round) {
                const uint8_t *tb_done = d_done;
                if (single_wave) {
                    CK(cudaMemcpyAsync(d_done_tb[par], d_done, (size_t)n, cudaMemcpyDeviceToDevice, main_stream));
                    tb_done = d_done_tb[par];
                    CK(cudaEventRecord(ctx->ev_fin, main_stream));
                    CK(cudaStreamWaitEvent(ctx->tb_stream, ctx->ev_fin, 0));
                    ctx->stream = ctx->tb_stream;
                }
                cudaError_t te;
                if (linear) te = launch_traceback_lin(ctx, pool, d_jobs, nj, tb_done, d_dir, d_out_off, d_resi, d_resj, d_out_len);
                else te = launch_traceback(ctx, cm, pool, d_jobs, nj, tb_done, d_dir, d_out_off, d_median, d_medianwg, d_resi,
                                           d_resj, d_out_len);
                ctx->stream = main_stream;
                if (te != cudaSuccess) return poy_cuda_fail(ctx, te, "traceback launch");
                if (single_wave) { CK(cudaEventRecord(ctx->ev_tb_done[par], ctx->tb_stream)); ctx->tb_pending[par] = true; }
            }
            { const double t = now(); t_prep += t - t_mark; t_mark = t; }
            CK(cudaStreamSynchronize(ctx->stream));  // the pinned job staging and the arena are reused by the next wave
            { const double t = now(); t_wait += t - t_mark; t_mark = t; }
            ++waves;
            pos = end;
            w_par = par; w_single = single_wave; w_jobs = d_jobs; w_dir = d_dir;
        }
        CK(cudaMemcpyAsync(h_done, d_done, spec_round ? nx : (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (trace_rounds) {
            int hist[4097 / 64 + 1] = {0}, ngf = 0, npr = 0;
            for (int p : rjobs) { hist[hp[p].dclass / 64]++; ngf += hp[p].gapfree; npr += hp[p].probe; }
            fprintf(stderr, "[poy5_b200]   round %d: %zu active + %zu speculative (%d gap-free, %d probes), %.2f ms since batch start; classes:", rounds,
                    active.size(), clones.size(), ngf, npr, (t_prep + t_wait) * 1e3);
            for (int c = 0; c <= 4096 / 64; ++c) if (hist[c]) fprintf(stderr, " %d:%d", c * 64, hist[c]);
            fprintf(stderr, "\n");
        }
        std::vector<int> next;
        next.reserve(active.size());
        if (spec_round) {
            // per pair: walk its chain in the order of the doublings; the first fill that stopped is the result
            std::vector<int> nxt(hp.size(), -1), posq(hp.size(), -1);
            for (size_t c = 0; c < clones.size(); ++c) nxt[clones[c].prev] = n + (int)c;
            for (size_t q = 0; q < order.size(); ++q) posq[order[q]] = (int)q;
            BandJob *hj = (BandJob *)v_pin, *htb = hj + nx;
            int ntb = 0;
            for (int p : active) {
                HostPair &h = hp[p];
                if (h.dclass == 0 || h.lasti == 0) {          // not part of the speculation: the plain verdicts
                    if (std::min(h.k, h.lasti) >= 2) h.eh00_inf = true;
                    const int verdict = h_done[p];
                    if (verdict == 1) { h.done = true; htb[ntb++] = hj[posq[p]]; continue; }
                    h.T *= 2; h.want_dirs = 0; h.repeat = 0;
                    next.push_back(p);
                    continue;
                }
                int winner = -1, nextT = h.T;
                bool first = true;
                for (int x = p; x >= 0; x = nxt[x]) {
                    const int verdict = h_done[x];
                    if (verdict >= 5) { ++n_repeat; break; }        // this fill gave up on its predecessor: next round starts here
                    if (!first) { h.iterations++; h.cells += band_cells(hp[x].lasti, hp[x].lastj, hp[x].k); }
                    first = false;
                    if (std::min(hp[x].k, hp[x].lasti) >= 2) h.eh00_inf = true;
                    if (verdict == 1) { winner = x; break; }
                    nextT = hp[x].T * 2;
                }
                if (winner < 0) { h.T = nextT; h.want_dirs = 0; h.repeat = 0; h.state_dirty = true; next.push_back(p); continue; }
                h.done = true;
                h.T = hp[winner].T; h.k = hp[winner].k; h.dclass = hp[winner].dclass; h.stride = hp[winner].stride;
                if (winner != p) CK(cudaMemcpyAsync(d_state + p, d_state + winner, sizeof(PairState), cudaMemcpyDeviceToDevice, ctx->stream));
                htb[ntb] = hj[posq[winner]];
                htb[ntb].pair = p;
                ++ntb;
            }
            if (want_trace && ntb > 0) {
                BandJob *d_tb = w_jobs + nx;
                CK(cudaMemcpyAsync(d_tb, htb, sizeof(BandJob) * (size_t)ntb, cudaMemcpyHostToDevice, ctx->stream));
                cudaStream_t main_stream = ctx->stream;
                if (w_single) {
                    CK(cudaEventRecord(ctx->ev_fin, main_stream));
                    CK(cudaStreamWaitEvent(ctx->tb_stream, ctx->ev_fin, 0));
                    ctx->stream = ctx->tb_stream;
                }
                cudaError_t te = launch_traceback(ctx, cm, pool, d_tb, ntb, nullptr, w_dir, d_out_off, d_median, d_medianwg, d_resi, d_resj, d_out_len);
                ctx->stream = main_stream;
                if (te != cudaSuccess) return poy_cuda_fail(ctx, te, "traceback launch");
                if (w_single) { CK(cudaEventRecord(ctx->ev_tb_done[w_par], ctx->tb_stream)); ctx->tb_pending[w_par] = true; }
            }
            active.swap(next);
            continue;
        }
        for (int p : active) {
            // (EH[0][0] as this fill leaves it, for a later speculative round: k_band2 / k_band_generic set it to INF
            // under this condition, probe fills and repeated fills alike)
            if (std::min(hp[p].k, hp[p].lasti) >= 2) hp[p].eh00_inf = true;
            const int verdict = h_done[p];   // k_band_finish / k_lin_finish
            if (verdict == 1) { hp[p].done = true; continue; }
            if (verdict == 2) { hp[p].want_dirs = 1; hp[p].repeat = 1; ++n_repeat; }   // same threshold again, with directions
            else { hp[p].T *= 2; hp[p].want_dirs = (verdict == 3); hp[p].repeat = 0; }
            next.push_back(p);
        }
        active.swap(next);
    }
    for (int q = 0; q < 2; ++q)      // the caller's stream order must see the traced-back outputs
        if (ctx->tb_pending[q]) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tb_done[q], 0)); ctx->tb_pending[q] = false; }
    if (d_cost) {
        CK(launch_gather_cost(ctx, d_state, n, d_cost));
    }
    {
        int64_t cells = 0;
        for (int p = 0; p < n; ++p) cells += hp[p].cells;
        ctx->stat_band_cells += cells; ctx->stat_probe += n_probe; ctx->stat_full += n_full; ctx->stat_repeat += n_repeat;
        ctx->stat_rounds += rounds; ctx->stat_pairs += n;
    }
    if (trace) fprintf(stderr, "[poy5_b200] align n=%d rounds=%d waves=%d host prep %.1f ms, device wait %.1f ms; fills: %lld probe, %lld full, %lld repeated\n",
                       n, rounds, waves, t_prep * 1e3, t_wait * 1e3, n_probe, n_full, n_repeat);
    if (h_stats)
        for (int p = 0; p < n; ++p) {
            h_stats[4 * p + 0] = hp[p].iterations; h_stats[4 * p + 1] = hp[p].T; h_stats[4 * p + 2] = hp[p].k;
            h_stats[4 * p + 3] = (int32_t)std::min<int64_t>(hp[p].cells / 1024, 0x7fffffff);
        }
    return POY_OK;
}

// Large batches are cut in two halves that run align_impl concurrently, the second one on the context's twin
// (own streams, scratch and host thread).  Every per-pair array is indexed by the pair's position in the batch, so
// the second half simply gets the pointers advanced by n0.  POY_SPLIT=0 turns this off.
poy_status align_split(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n, const int32_t *h_si,
                       const int32_t *h_sj, const uint8_t *h_swaped, const int64_t *d_out_off, int32_t *d_cost,
                       uint8_t *d_median, uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj,
                       int32_t *d_out_len, int32_t *h_stats, const int32_t *h_deltawh) {
    const bool linear = h_deltawh != nullptr;
    const char *se = getenv("POY_SPLIT"), *sm = getenv("POY_SPLIT_MIN");   // POY_SPLIT_MIN: smallest batch that is split (test hook)
    const int split_min = sm ? std::max(2, atoi(sm)) : 4096;
    const bool model_ok = linear ? cm->h.cost_model_type != 1 : cm->h.cost_model_type == 1;
    if (n < split_min || ctx->is_twin || !model_ok || (se && se[0] == '0'))
        return align_impl(ctx, cm, pool, n, h_si, h_sj, h_swaped, d_out_off, d_cost, d_median, d_medianwg, d_resi, d_resj,
                          d_out_len, h_stats, h_deltawh);
    if (!ctx->twin) {
        poy_status cs = poy_ctx_create(ctx->device, nullptr, &ctx->twin);
        if (cs != POY_OK) return poy_fail(ctx, cs, "second lane: context creation failed");
        ctx->twin->is_twin = true;
    }
    poy_ctx *tw = ctx->twin;
    poy_status s = ensure_params(ctx, cm, pool);     // once, before the lanes diverge
    if (s != POY_OK) return s;
    CK(cudaEventRecord(ctx->ev_twin_start, ctx->stream));
    CK(cudaStreamWaitEvent(tw->stream, ctx->ev_twin_start, 0));
    // the two lanes share the direction arena half and half; restored on every exit path
    struct ArenaGuard {
        poy_ctx *c; uint64_t full;
        ~ArenaGuard() { c->arena_limit = full; }
    } guard{ ctx, ctx->arena_limit };
    ctx->arena_limit = tw->arena_limit = std::max<uint64_t>(guard.full / 2, 1ull << 20);
    const int32_t n0 = n / 2, n1 = n - n0;
    const int per = linear ? 2 : 4;
    poy_status s1 = POY_OK;
    const uint64_t tw_launches0 = tw->launches;
    auto second = [&] {
        cudaSetDevice(ctx->device);
        s1 = align_impl(tw, cm, pool, n1, h_si + n0, h_sj + n0, h_swaped ? h_swaped + n0 : nullptr,
                        d_out_off ? d_out_off + n0 : nullptr, d_cost ? d_cost + n0 : nullptr, d_median, d_medianwg, d_resi, d_resj,
                        d_out_len ? d_out_len + (size_t)per * n0 : nullptr, h_stats ? h_stats + 4 * (size_t)n0 : nullptr,
                        h_deltawh ? h_deltawh + n0 : nullptr);
    };
    std::thread lane;
    bool threaded = true;
    try { lane = std::thread(second); } catch (...) { threaded = false; }   // no C++ exception may cross the C ABI
    const poy_status s0 = align_impl(ctx, cm, pool, n0, h_si, h_sj, h_swaped, d_out_off, d_cost, d_median, d_medianwg, d_resi,
                                     d_resj, d_out_len, h_stats, h_deltawh);
    if (threaded) lane.join(); else second();   // thread creation failed: the second half runs after the first
    ctx->launches += tw->launches - tw_launches0;
    ctx->stat_band_cells += tw->stat_band_cells; ctx->stat_probe += tw->stat_probe; ctx->stat_full += tw->stat_full;
    ctx->stat_repeat += tw->stat_repeat; ctx->stat_rounds += tw->stat_rounds; ctx->stat_pairs += tw->stat_pairs;
    tw->stat_band_cells = tw->stat_probe = tw->stat_full = tw->stat_repeat = tw->stat_rounds = tw->stat_pairs = 0;
    cudaEventRecord(ctx->ev_twin_done, tw->stream);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_twin_done, 0);
    if (s0 != POY_OK && s1 != POY_OK) {
        char both[sizeof ctx->err];
        snprintf(both, sizeof both, "%.240s; second lane: %.240s", ctx->err, tw->err);
        snprintf(ctx->err, sizeof ctx->err, "%s", both);
        return s0;
    }
    if (s0 != POY_OK) return s0;
    if (s1 != POY_OK) { snprintf(ctx->err, sizeof ctx->err, "%s", tw->err); return s1; }
    return POY_OK;
}

extern "C" poy_status poy_batch_align_affine_dev(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                                 const int32_t *d_si, const int32_t *d_sj, const uint8_t *d_swaped,
                                                 const int32_t *h_si, const int32_t *h_sj, const int64_t *d_out_off,
                                                 int32_t *d_cost, uint8_t *d_median, uint8_t *d_medianwg,
                                                 uint8_t *d_resi, uint8_t *d_resj, int32_t *d_out_len, int32_t *d_stats) {
    bind_device(ctx);
    (void)d_si; (void)d_sj;
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!h_si || !h_sj))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    // the band schedule is driven from the host, so the swaped flags are needed there as well
    std::vector<uint8_t> sw;
    if (d_swaped) {
        sw.resize((size_t)n);
        CK(cudaMemcpyAsync(sw.data(), d_swaped, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int32_t> stats;
    if (d_stats) stats.resize(4 * (size_t)n);
    poy_status s = align_split(ctx, cm, pool, n, h_si, h_sj, d_swaped ? sw.data() : nullptr, d_out_off, d_cost, d_median,
                              d_medianwg, d_resi, d_resj, d_out_len, d_stats ? stats.data() : nullptr);
    if (s != POY_OK) return s;
    if (d_stats) {
        CK(cudaMemcpyAsync(d_stats, stats.data(), sizeof(int32_t) * 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return POY_OK;
}

extern "C" poy_status poy_batch_align_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                             const int32_t *si, const int32_t *sj, const uint8_t *swaped,
                                             const int64_t *out_off, int32_t *cost, uint8_t *median, uint8_t *medianwg,
                                             uint8_t *resi, uint8_t *resj, int32_t *out_len, int32_t *stats) {
    bind_device(ctx);
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!si || !sj))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    const bool want_trace = median || medianwg || resi || resj || out_len;
    if (want_trace && !out_off) return poy_fail(ctx, POY_ERR_ARG, "out_off is required when any traceback output is requested");
    // total output bytes = end of the last slot
    int64_t total = 0;
    if (want_trace)
        for (int p = 0; p < n; ++p) {
            if (si[p] < 0 || si[p] >= pool->nseq || sj[p] < 0 || sj[p] >= pool->nseq) return poy_fail(ctx, POY_ERR_ARG, "pair index out of range");
            const int64_t cap = (pool->h_off[si[p] + 1] - pool->h_off[si[p]]) + (pool->h_off[sj[p] + 1] - pool->h_off[sj[p]]) + 2;
            total = std::max(total, out_off[p] + cap);
        }
    const int nout = (median ? 1 : 0) + (medianwg ? 1 : 0) + (resi ? 1 : 0) + (resj ? 1 : 0);
    const size_t al_total = ((size_t)total + 255) & ~(size_t)255;
    void *v_out;
    const size_t need = sizeof(int64_t) * (size_t)n + sizeof(int32_t) * 5 * (size_t)n + al_total * (size_t)nout + 1024;
    poy_status s = poy_scratch(ctx, SL_JOBS2, need, &v_out);
    if (s != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v_out;
    int64_t *d_out_off = (int64_t *)cur; cur += sizeof(int64_t) * (size_t)n;
    int32_t *d_cost = (int32_t *)cur; cur += sizeof(int32_t) * (size_t)n;
    int32_t *d_len = (int32_t *)cur; cur += sizeof(int32_t) * 4 * (size_t)n;
    cur = (uint8_t *)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
    uint8_t *d_m = nullptr, *d_w = nullptr, *d_i = nullptr, *d_j = nullptr;
    if (median) { d_m = cur; cur += al_total; }
    if (medianwg) { d_w = cur; cur += al_total; }
    if (resi) { d_i = cur; cur += al_total; }
    if (resj) { d_j = cur; cur += al_total; }
    if (want_trace) CK(cudaMemcpyAsync(d_out_off, out_off, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    s = align_split(ctx, cm, pool, n, si, sj, swaped, want_trace ? d_out_off : nullptr, d_cost, d_m, d_w, d_i, d_j,
                   want_trace ? d_len : nullptr, stats);
    if (s != POY_OK) return s;
    if (cost) CK(cudaMemcpyAsync(cost, d_cost, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_len) CK(cudaMemcpyAsync(out_len, d_len, sizeof(int32_t) * 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (median) CK(cudaMemcpyAsync(median, d_m, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    if (medianwg) CK(cudaMemcpyAsync(medianwg, d_w, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    if (resi) CK(cudaMemcpyAsync(resi, d_i, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    if (resj) CK(cudaMemcpyAsync(resj, d_j, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

// ---- batch twins of algn_CAML_simple_2 / algn_CAML_align_2d (linear gap) ---------------------------------------
extern "C" poy_status poy_batch_align_linear(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                             const int32_t *s1, const int32_t *s2, const int32_t *deltawh,
                                             const uint8_t *swaped, const int64_t *out_off, int32_t *cost, uint8_t *r1,
                                             uint8_t *r2, int32_t *out_len, int32_t *stats) {
    bind_device(ctx);
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!s1 || !s2 || !deltawh))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    const bool want_trace = r1 || r2 || out_len;
    if (want_trace && !out_off) return poy_fail(ctx, POY_ERR_ARG, "out_off is required when any traceback output is requested");
    int64_t total = 0;
    for (int p = 0; p < n; ++p) {
        if (s1[p] < 0 || s1[p] >= pool->nseq || s2[p] < 0 || s2[p] >= pool->nseq) return poy_fail(ctx, POY_ERR_ARG, "pair index out of range");
        if (want_trace) {
            const int64_t cap = (pool->h_off[s1[p] + 1] - pool->h_off[s1[p]]) + (pool->h_off[s2[p] + 1] - pool->h_off[s2[p]]);
            total = std::max(total, out_off[p] + cap);
        }
    }
    const size_t al_total = ((size_t)total + 255) & ~(size_t)255;
    void *v_out;
    const size_t need = sizeof(int64_t) * (size_t)n + sizeof(int32_t) * 3 * (size_t)n + al_total * 2 + 1024;
    poy_status s = poy_scratch(ctx, SL_JOBS2, need, &v_out);
    if (s != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v_out;
    int64_t *d_out_off = (int64_t *)cur; cur += sizeof(int64_t) * (size_t)n;
    int32_t *d_cost = (int32_t *)cur; cur += sizeof(int32_t) * (size_t)n;
    int32_t *d_len = (int32_t *)cur; cur += sizeof(int32_t) * 2 * (size_t)n;
    cur = (uint8_t *)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
    uint8_t *d_1 = nullptr, *d_2 = nullptr;
    if (r1) { d_1 = cur; cur += al_total; }
    if (r2) { d_2 = cur; cur += al_total; }
    if (want_trace) CK(cudaMemcpyAsync(d_out_off, out_off, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    s = align_split(ctx, cm, pool, n, s1, s2, swaped, want_trace ? d_out_off : nullptr, d_cost, nullptr, nullptr, d_1, d_2,
                   want_trace ? d_len : nullptr, stats, deltawh);
    if (s != POY_OK) return s;
    if (cost) CK(cudaMemcpyAsync(cost, d_cost, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_len) CK(cudaMemcpyAsync(out_len, d_len, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (r1) CK(cudaMemcpyAsync(r1, d_1, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    if (r2) CK(cudaMemcpyAsync(r2, d_2, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

extern "C" poy_status poy_batch_cost_linear(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                            const int32_t *s1, const int32_t *s2, const int32_t *deltawh, int32_t *cost) {
    bind_device(ctx);
    return poy_batch_align_linear(ctx, cm, pool, n, s1, s2, deltawh, nullptr, nullptr, cost, nullptr, nullptr, nullptr, nullptr);
}

// ---- column-wise helpers over aligned rows ----------------------------------------------------------------------
namespace {
struct RowsOnDevice { uint8_t *a, *b, *out; int64_t *off, *out_off; int *len, *res; int64_t total, out_total; };

// stages rows_a / rows_b (+ offsets, lengths and, if given, output offsets with `extra` bytes of slack per pair)
poy_status stage_rows(poy_ctx *ctx, int n, const uint8_t *rows_a, const uint8_t *rows_b, const int64_t *off, const int32_t *len,
                      const int64_t *out_off, int extra, RowsOnDevice *r) {
    r->total = 0; r->out_total = 0;
    for (int p = 0; p < n; ++p) {
        if (off[p] < 0 || len[p] < 0) return poy_fail(ctx, POY_ERR_ARG, "negative offset or length");
        r->total = std::max(r->total, off[p] + len[p]);
        if (out_off) r->out_total = std::max(r->out_total, out_off[p] + len[p] + extra);
    }
    const size_t A = ((size_t)r->total + 255) & ~(size_t)255, O = ((size_t)r->out_total + 255) & ~(size_t)255;
    void *v;
    poy_status s = poy_scratch(ctx, SL_JOBS2, 2 * A + O + (size_t)n * (8 + 8 + 4 + 4) + 2048, &v);
    if (s != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v;
    r->a = cur; cur += A; r->b = cur; cur += A; r->out = cur; cur += O;
    cur = (uint8_t *)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
    r->off = (int64_t *)cur; cur += 8 * (size_t)n; r->out_off = (int64_t *)cur; cur += 8 * (size_t)n;
    r->len = (int *)cur; cur += 4 * (size_t)n; r->res = (int *)cur;
    CK(cudaMemcpyAsync(r->a, rows_a, (size_t)r->total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r->b, rows_b, (size_t)r->total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r->off, off, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r->len, len, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (out_off) CK(cudaMemcpyAsync(r->out_off, out_off, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return POY_OK;
}
}  // namespace

extern "C" poy_status poy_batch_median_2(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                         const int64_t *off, const int32_t *len, int32_t with_gaps, const int64_t *out_off,
                                         uint8_t *out, int32_t *out_len) {
    bind_device(ctx);
    if (!ctx || !cm || n < 0 || (n > 0 && (!rows_a || !rows_b || !off || !len || !out_off || !out || !out_len))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    RowsOnDevice r;
    poy_status s = stage_rows(ctx, n, rows_a, rows_b, off, len, out_off, 1, &r);
    if (s != POY_OK) return s;
    CK(launch_median_2(ctx, cm, n, r.a, r.b, r.off, r.len, with_gaps, r.out_off, r.out, r.res));
    CK(cudaMemcpyAsync(out, r.out, (size_t)r.out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, r.res, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

// Sequence.Align.closest, the column-wise half (src/sequence.ml:1217-1228): rows_parent / rows_mine are the two
// aligned rows, every column becomes get_closest cm parent.(i) mine.(i), gaps are squeezed out and the leading gap
// is put back (remove_gaps2, src/sequence.ml:209-222).
extern "C" poy_status poy_batch_closest(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_parent, const uint8_t *rows_mine,
                                        const int64_t *off, const int32_t *len, const int64_t *out_off, uint8_t *out, int32_t *out_len) {
    bind_device(ctx);
    if (!ctx || !cm || n < 0 || (n > 0 && (!rows_parent || !rows_mine || !off || !len || !out_off || !out || !out_len))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    RowsOnDevice r;
    poy_status s = stage_rows(ctx, n, rows_parent, rows_mine, off, len, out_off, 1, &r);
    if (s != POY_OK) return s;
    CK(launch_median_2(ctx, cm, n, r.a, r.b, r.off, r.len, 2, r.out_off, r.out, r.res));
    CK(cudaMemcpyAsync(out, r.out, (size_t)r.out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, r.res, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

extern "C" poy_status poy_batch_union(poy_ctx *ctx, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b, const int64_t *off,
                                      const int32_t *len, uint8_t *out) {
    bind_device(ctx);
    if (!ctx || n < 0 || (n > 0 && (!rows_a || !rows_b || !off || !len || !out))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    RowsOnDevice r;
    poy_status s = stage_rows(ctx, n, rows_a, rows_b, off, len, off, 0, &r);
    if (s != POY_OK) return s;
    CK(launch_union(ctx, r.total, r.a, r.b, r.out));
    CK(cudaMemcpyAsync(out, r.out, (size_t)r.total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

extern "C" poy_status poy_batch_aligned_cost(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                             const int64_t *off, const int32_t *len, int32_t use_worst, int32_t *cost) {
    bind_device(ctx);
    if (!ctx || !cm || n < 0 || (n > 0 && (!rows_a || !rows_b || !off || !len || !cost))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    RowsOnDevice r;
    poy_status s = stage_rows(ctx, n, rows_a, rows_b, off, len, nullptr, 0, &r);
    if (s != POY_OK) return s;
    CK(launch_aligned_cost(ctx, cm, n, r.a, r.b, r.off, r.len, use_worst, r.res));
    CK(cudaMemcpyAsync(cost, r.res, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}

extern "C" poy_status poy_batch_ancestor_2(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                           const int64_t *off, const int32_t *len, const int64_t *out_off, uint8_t *out,
                                           int32_t *out_len) {
    bind_device(ctx);
    if (!ctx || !cm || n < 0 || (n > 0 && (!rows_a || !rows_b || !off || !len || !out_off || !out || !out_len))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    RowsOnDevice r;
    poy_status s = stage_rows(ctx, n, rows_a, rows_b, off, len, out_off, 1, &r);
    if (s != POY_OK) return s;
    CK(launch_ancestor_2(ctx, cm, n, r.a, r.b, r.off, r.len, r.out_off, r.out, r.res));
    CK(cudaMemcpyAsync(out, r.out, (size_t)r.out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, r.res, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < n; ++p) if (out_len[p] < 0) return poy_fail(ctx, POY_ERR_ARG, "median should not be 0");
    return POY_OK;
}

// ---- micro-benchmark ---------------------------------------------------------------------------------------
extern "C" poy_status poy_microbench_int(poy_ctx *ctx, int32_t kind, double *ops_per_second, double *sm_clock_mhz) {
    bind_device(ctx);
    if (!ctx || !ops_per_second) return POY_ERR_ARG;
    void *v;
    poy_status s = poy_scratch(ctx, SL_MISC, 1 << 20, &v);
    if (s != POY_OK) return s;
    float ms = 0;
    double ops = 0;
    CK(launch_microbench(ctx, kind, 4096, (unsigned long long *)v, (int *)v + 1024, &ms, &ops));
    *ops_per_second = ops / (ms * 1e-3);
    if (sm_clock_mhz) {
        unsigned long long cyc[2] = { 0, 1 };
        CK(cudaMemcpy(cyc, v, sizeof cyc, cudaMemcpyDeviceToHost));
        *sm_clock_mhz = cyc[1] ? (double)cyc[0] / (double)cyc[1] * 1e3 : 0.0;  // cycles per ns -> MHz
    }
    return POY_OK;
}
