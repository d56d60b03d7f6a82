// Three-dimensional cost model: Cost_matrix.Three_D.of_two_dim_comb restated in C++ (src/cost_matrix.ml:1605-1652,
// bit-indexed alphabets with combinations, the DNA case), `struct cm_3d` tables on the device (src/cm.h:253-280,
// cm_set_val_3d src/cm.c:704) and the column-wise three-way median that the live 3-D path applies to three aligned
// rows (Sequence.Align.align_3_powell_inter, src/sequence.ml:1342-1369: Three_D.median per column, medianWgap keeps
// every column, median drops the pure-gap ones and gets the leading gap back).
// (algn_get_median_3d, src/algn.c:3641-3657, never advances its three pointers and has no caller; not built.)
#include <string.h>
#include "common.cuh"

struct poy_cm3d {
    uint8_t *d_median;     // 32^3
    int32_t *d_cost;       // 32^3
    poy_cm3d_host h;
};

extern "C" poy_status poy_cm3d_fill(const poy_cm_host *m, poy_cm3d_host *out) {
    if (!m || !out) return POY_ERR_ARG;
    const int alph = 31, lcm = 5, gap = 16;
    const int32_t MAXI = 0x3fffffff;          // max_int of cost_matrix.ml (src/cost_matrix.ml:38)
    memset(out, 0, sizeof *out);
    for (int i = 1; i <= alph; ++i)
        for (int j = 1; j <= alph; ++j)
            for (int k = 1; k <= alph; ++k) {
                const int pos = (((i << lcm) + j) << lcm) + k;      // cost_position, src/cost_matrix.ml:1532-1533
                int32_t best = MAXI; int med = 0;
                for (int l = 0; l < lcm; ++l) {
                    const int inter = 1 << l;
                    const int shared = ((i & inter) ? 1 : 0) + ((j & inter) ? 1 : 0) + ((k & inter) ? 1 : 0);
                    int32_t c;
                    if (m->is_metric || shared >= 2 || inter != gap)
                        c = m->cost[(inter << 5) + i] + m->cost[(inter << 5) + j] + m->cost[(inter << 5) + k];
                    else c = MAXI;
                    if (c < best) { best = c; med = inter; }
                    else if (c == best) med |= inter;
                }
                int pick = 1;                                        // pick_bit 1: the lowest set bit
                while (pick < (1 << lcm) && !(pick & med)) pick <<= 1;
                out->cost[pos] = best;
                out->median[pos] = (uint8_t)pick;
            }
    return POY_OK;
}

extern "C" poy_status poy_cm3d_upload(poy_ctx *ctx, const poy_cm3d_host *h, poy_cm3d **out) {
    bind_device(ctx);
    if (!ctx || !h || !out) return POY_ERR_ARG;
    *out = nullptr;
    poy_cm3d *c = new poy_cm3d;
    c->h = *h; c->d_median = nullptr; c->d_cost = nullptr;
    cudaError_t e = cudaMalloc(&c->d_median, sizeof h->median);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cost, sizeof h->cost);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_median, h->median, sizeof h->median, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_cost, h->cost, sizeof h->cost, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(c->d_median); cudaFree(c->d_cost); delete c; return poy_cuda_fail(ctx, e, "poy_cm3d_upload"); }
    *out = c;
    return POY_OK;
}

extern "C" void poy_cm3d_free(poy_ctx *ctx, poy_cm3d *c) {
    bind_device(ctx);
    if (!c) return;
    cudaFree(c->d_median); cudaFree(c->d_cost);
    delete c;
}

namespace {
// warp per triple of aligned rows; medianwg = one symbol per column, median = the non-gap ones after a leading gap;
// cost3 = sum of Three_D.cost over the columns
__global__ void __launch_bounds__(128)
k_median_3(const uint8_t *__restrict__ med3, const int32_t *__restrict__ cost3, int n, const uint8_t *__restrict__ a,
           const uint8_t *__restrict__ b, const uint8_t *__restrict__ c, const int64_t *__restrict__ off, const int *__restrict__ len,
           const int64_t *__restrict__ out_off, uint8_t *median, uint8_t *medianwg, int *out_len, int *cost) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= n) return;
    const uint8_t *ra = a + off[p], *rb = b + off[p], *rc = c + off[p];
    uint8_t *om = median + out_off[p], *ow = medianwg + out_off[p];
    const int L = len[p];
    int w = 1, csum = 0;
    if (lane == 0) om[0] = POY_GAP;
    for (int x0 = 0; x0 < L; x0 += 32) {
        const int x = x0 + lane;
        int m = 0; bool keep = false;
        if (x < L) {
            const int pos = ((((ra[x] & 31) << 5) + (rb[x] & 31)) << 5) + (rc[x] & 31);
            m = med3[pos]; csum += cost3[pos];
            ow[x] = (uint8_t)m;
            keep = m != POY_GAP;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) om[w + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)m;
        w += __popc(mask);
    }
    for (int o = 16; o; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    if (lane == 0) { out_len[p] = w; if (cost) cost[p] = csum; }
}
}  // namespace

// rows_a/b/c: packed aligned rows (HOST), triple p = rows_x[off[p] .. off[p] + len[p]); median slot of triple p starts at
// out_off[p] (capacity len[p] + 1), medianwg at out_off[p] (len[p] bytes)
extern "C" poy_status poy_batch_median_3(poy_ctx *ctx, const poy_cm3d *cm3, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                         const uint8_t *rows_c, const int64_t *off, const int32_t *len, const int64_t *out_off,
                                         uint8_t *median, uint8_t *medianwg, int32_t *out_len, int32_t *cost3) {
    bind_device(ctx);
    if (!ctx || !cm3 || n < 0 || (n > 0 && (!rows_a || !rows_b || !rows_c || !off || !len || !out_off || !median || !medianwg || !out_len)))
        return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    int64_t total = 0, out_total = 0;
    for (int p = 0; p < n; ++p) {
        if (len[p] < 0) return poy_fail(ctx, POY_ERR_ARG, "negative row length");
        total = std::max<int64_t>(total, off[p] + len[p]);
        out_total = std::max<int64_t>(out_total, out_off[p] + len[p] + 1);
    }
    const size_t A = ((size_t)total + 255) & ~(size_t)255, O = ((size_t)out_total + 255) & ~(size_t)255;
    void *v;
    poy_status s = poy_scratch(ctx, SL_JOBS2, 3 * A + 2 * O + (size_t)n * (8 + 8 + 4 + 4 + 4) + 1024, &v);
    if (s != POY_OK) return s;
    uint8_t *cur = (uint8_t *)v;
    uint8_t *d_a = cur; cur += A; uint8_t *d_b = cur; cur += A; uint8_t *d_c = cur; cur += A;
    uint8_t *d_m = cur; cur += O; uint8_t *d_w = cur; cur += O;
    int64_t *d_off = (int64_t *)cur; cur += 8 * (size_t)n; int64_t *d_oo = (int64_t *)cur; cur += 8 * (size_t)n;
    int *d_len = (int *)cur; cur += 4 * (size_t)n; int *d_ol = (int *)cur; cur += 4 * (size_t)n; int *d_cost = (int *)cur;
    CK(cudaMemcpyAsync(d_a, rows_a, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_b, rows_b, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_c, rows_c, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off, off, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_oo, out_off, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_len, len, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    k_median_3<<<(n + 3) / 4, 128, 0, ctx->stream>>>(cm3->d_median, cm3->d_cost, n, d_a, d_b, d_c, d_off, d_len, d_oo, d_m, d_w, d_ol, d_cost);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(median, d_m, (size_t)out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(medianwg, d_w, (size_t)out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, d_ol, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (cost3) CK(cudaMemcpyAsync(cost3, d_cost, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}
