// Sequence.NewkkAlign on the device: the diagonal-storage Ukkonen alignment of src/newkkonen.c (affine model:
// newkkonen_CAML_algn_affine :1472-1495 -> newkk_algn :1360-1444 -> increaseT :1155-1171 -> ukktest :1077-1153 ->
// update_internal_cell :680-870 / assign_best_cost_and_direction :600-677; newkkonen_CAML_backtrace_affine ->
// backtrace_affine :1666-1763, trivial_algn / trivial_backtrace :1351-1356, 1498-1521).
//
// The reference keeps one matrix in diagonal-major storage across threshold doublings and re-evaluates only the cells
// whose neighbours changed (two queues per row).  A cell is a pure function of its three neighbours and its border
// status, and every cell whose inputs or border status changed is re-evaluated, so the matrix after a doubling equals
// a FRESH fill of the new band -- which is what a GPU wants: here one CTA owns a pair, every thread owns band
// diagonals (the reference's storage unit), the latest cell of each diagonal lives in shared memory (global scratch
// for bands wider than NK_SMEM_DIAGS) and the band is swept by anti-diagonals.  The whole threshold-doubling loop of
// increaseT runs inside ONE launch (no host round trip per doubling).  Direction words (the reference's
// DIRECTION_MATRIX, 11 bits) are only written by a second fill of the final band, band-only and anti-diagonal major
// (cell (i,j) at dir[(i+j) * wh + ((j-i+k) >> 1)]), followed by the traceback, one thread per pair.
// The non-affine entry point is not offered: update_internal_cell never sets costDiag there (:796, :855-856), every
// interior cell gets cost 0 and the traceback leaves the band and raises (tests/test_newkk.py shows it on the compiled reference).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

#define NK_INF 0x3fffffff
#define NK_MUCH_LONGER 100
#define NK_ALIGN 1
#define NK_DO_DELETE 16
#define NK_DO_INSERT 32
#define NK_END_INSERT 64
#define NK_END_DELETE 128
#define NK_END_DIAG 256
#define NK_DO_DIAG 512
#define NK_INS_EQ_DEL 1024
#define NK_THREADS 256
#define NK_SMEM_DIAGS 4096          // diagonals whose state fits the dynamic shared memory (6 planes x 4 B x 4096 = 96 KB)

struct NkJob {
    int64_t off1, off2;   // pool offsets of the shorter / longer sequence
    int len1, len2;       // lengths INCLUDING the leading gap
    int swaped, pair;
    int k;                // final half band (second pass)
    int wh;               // direction words per anti-diagonal
    int64_t dir_off;      // offset (in words) of this pair's direction block
};
struct NkResult { int cost, k, iterations, trivial; long long cells; };

namespace {

__device__ __forceinline__ int nk_add(int a, int b) { return (a >= NK_INF || b >= NK_INF) ? NK_INF : a + b; }
__device__ __forceinline__ int nk_go(int base, int prev, int idx, int go) {
    if (idx == 1 && (base & 16)) return 0;
    if (idx > 1 && !(prev & 16) && (base & 16)) return 0;
    return go;
}

// One fill of the band with half width k.  Planes (cost, P, Q, ED, CD, G = g1 | g2 << 16) hold the latest cell of every
// band diagonal slot ds = j - i + k.  Returns (through the planes) the state of cell (len1-1, len2-1) in slot delta + k.
template <bool DIR>
__device__ void nk_fill(const int *s_cost, const uint8_t *__restrict__ s1, const uint8_t *__restrict__ s2, int len1, int len2,
                        int k, int realgo, const int *row0, int *pl, int ws, uint16_t *dir, int wh) {
    const int delta = len2 - len1, W = delta + 2 * k + 1, bb = delta + 1;
    int *pC = pl, *pP = pl + ws, *pQ = pP + ws, *pE = pQ + ws, *pD = pE + ws;
    unsigned *pG = (unsigned *)(pD + ws);
    // row 0 (newkk_algn :1390-1431; cells beyond the base band come out of update_internal_cell with only a left
    // neighbour: the same prefix sum)
    for (int ds = threadIdx.x; ds < W; ds += blockDim.x) {
        const int j = ds - k;
        if (j >= 0 && j < len2) {
            if (j == 0) { pC[ds] = 0; pP[ds] = realgo; pQ[ds] = realgo; pE[ds] = NK_INF; pD[ds] = 0; pG[ds] = 0u; }
            else { pC[ds] = row0[j]; pP[ds] = NK_INF; pQ[ds] = row0[j]; pE[ds] = NK_INF; pD[ds] = NK_INF; pG[ds] = (unsigned)j & 0xFFFFu; }
            if (DIR) dir[(size_t)j * wh + (ds >> 1)] = j == 0 ? (uint16_t)0 :
                         (uint16_t)(NK_DO_INSERT | NK_END_DELETE | NK_END_DIAG | ((j < bb || j == 1) ? NK_END_INSERT : 0));
        }
    }
    __syncthreads();
    const int a_end = len1 + len2 - 2;
    for (int a = 1; a <= a_end; ++a) {
        const int par = (a + k) & 1;
        for (int ds = 2 * threadIdx.x + par; ds < W; ds += 2 * blockDim.x) {
            const int d = ds - k, i = (a - d) >> 1, j = (a + d) >> 1;
            if (i < 1 || i >= len1 || j < 0 || j >= len2) continue;
            const int b1 = s1[i], b2 = s2[j], p1 = s1[i - 1], p2 = j > 0 ? s2[j - 1] : 0;
            const int go1 = nk_go(b1, p1, i, realgo), go2 = nk_go(b2, p2, j, realgo);
            const int xg1 = ((p1 & 16) && !(b1 & 16)) ? realgo : 0, xg2 = ((p2 & 16) && !(b2 & 16)) ? realgo : 0;
            int thisP = NK_INF, thisQ = NK_INF, thisED = NK_INF, thisCD = NK_INF;
            int costL, extL, openL, costR, extR, openR, costM, costD, extD, openD;
            int g1L = 0, g2L = 0, g1R = 0, g2R = 0, g1M = 0, g2M = 0;
            if (ds == 0 || j == 0) costL = extL = openL = NK_INF;          // left border of the band / of the matrix
            else {
                const int add = s_cost[(b2 << 5) + 16];
                extL = nk_add(pQ[ds - 1], add + xg2); openL = nk_add(pD[ds - 1], add + go2);
                costL = min(openL, extL); thisQ = costL;
                const unsigned g = pG[ds - 1]; g1L = g & 0xFFFF; g2L = g >> 16;
            }
            if (ds == W - 1) costR = extR = openR = NK_INF;                // right border (i >= 1 here)
            else {
                const int add = s_cost[(b1 << 5) + 16];
                extR = nk_add(pP[ds + 1], add + xg1); openR = nk_add(pD[ds + 1], add + go1);
                costR = min(openR, extR); thisP = costR;
                const unsigned g = pG[ds + 1]; g1R = g & 0xFFFF; g2R = g >> 16;
            }
            if (j == 0) costM = costD = extD = openD = NK_INF;
            else {
                const int b1n = b1 & 15, b2n = b2 & 15;
                const int add = s_cost[(b1n << 5) + b2n];
                const int mP = pP[ds], mQ = pQ[ds], mE = pE[ds], mD = pD[ds];
                const int fromR = mP + s_cost[(b1n << 5) + b2] + (b1n == b1 ? 0 : realgo);
                const int fromL = mQ + s_cost[(b1 << 5) + b2n] + (b2n == b2 ? 0 : realgo);
                thisCD = nk_add(mD, add);
                thisCD = min(thisCD, min(fromR, fromL));
                thisCD = min(thisCD, nk_add(mE, go1 + go2 + add));
                openD = nk_add(mD, go1 + go2);
                extD = nk_add(mE, ((b1 & 16) && (b2 & 16)) ? 0 : NK_INF);
                thisED = min(extD, openD);
                costM = thisCD; costD = thisED;
                const unsigned g = pG[ds]; g1M = g & 0xFFFF; g2M = g >> 16;
            }
            int best = costL, dw = NK_DO_INSERT, r1 = g1L + 1, r2 = g2L;
            if (costR <= best) {
                if (costR < best) { best = costR; dw = NK_DO_DELETE; r1 = g1R; r2 = g2R + 1; }
                else { dw |= NK_DO_DELETE; r1 = max(r1, g1R); r2 = max(r2, g2R + 1); }
            }
            if (costM <= best) {
                if (costM < best) { best = thisCD; dw = NK_ALIGN; r1 = g1M; r2 = g2M; }
                else { dw |= NK_ALIGN; r1 = max(r1, g1M); r2 = max(r2, g2M); }
            }
            if (costD <= best) {
                if (costD < best) { best = costD; dw = NK_DO_DIAG; r1 = g1M; r2 = g2M; }
                else { dw |= NK_DO_DIAG; r1 = max(r1, g1M); r2 = max(r2, g2M); }
            }
            if (DIR) {
                if (extR >= openR) dw |= NK_END_DELETE;
                if (extL >= openL) dw |= NK_END_INSERT;
                if (extD >= openD) dw |= NK_END_DIAG;
                if (extR == extL && extR == best) dw |= NK_INS_EQ_DEL;
                dir[(size_t)a * wh + (ds >> 1)] = (uint16_t)dw;
            }
            pC[ds] = best; pP[ds] = thisP; pQ[ds] = thisQ; pE[ds] = thisED; pD[ds] = thisCD;
            pG[ds] = ((unsigned)r1 & 0xFFFFu) | ((unsigned)r2 << 16);
        }
        __syncthreads();
    }
}

// FINAL = false: the threshold-doubling loop of increaseT without direction words -> NkResult (cost, final k, ...)
// FINAL = true:  one fill of the band J.k WITH direction words
template <bool FINAL>
__global__ void __launch_bounds__(NK_THREADS)
k_newkk(const DevCM *__restrict__ cm, const uint8_t *__restrict__ data, const NkJob *__restrict__ jobs, int njobs, int *counter,
        int *work, size_t work_stride, int ws_global, NkResult *res, uint16_t *dir) {
    extern __shared__ int s_dyn[];
    __shared__ int s_cost[1024];
    __shared__ int s_job, s_stop;
    __shared__ long long s_sum;
    // cm_calc_cost on the calloc'd table: row / column 0 are never set (src/cm.c:627)
    for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_cost[x] = ((x >> 5) == 0 || (x & 31) == 0) ? 0 : cm->cost32[x];
    const int realgo = cm->gap_open > 0 ? cm->gap_open : 0;
    const int delta_cost = cm->min_non0;
    int *gwork = work + (size_t)blockIdx.x * work_stride;
    int *row0 = gwork;                       // len2 ints
    int *gplanes = gwork + ws_global;        // 6 planes of ws_global ints
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_job = atomicAdd(counter, 1); s_sum = 0; }
        __syncthreads();
        const int job = s_job;
        if (job >= njobs) break;
        const NkJob J = jobs[job];
        const uint8_t *s1 = data + J.off1, *s2 = data + J.off2;
        const int len1 = J.len1, len2 = J.len2, delta = len2 - len1;
        if ((long long)len1 * NK_MUCH_LONGER < len2) {          // trivial_algn
            if (!FINAL) {
                long long part = 0;
                for (int x = threadIdx.x; x < len1 + len2; x += blockDim.x) {
                    const int c = x < len1 ? s1[x] : s2[x - len1];
                    part += s_cost[((c & 15) << 5) + 16];
                }
                atomicAdd((unsigned long long *)&s_sum, (unsigned long long)part);
                __syncthreads();
                if (threadIdx.x == 0) { NkResult r; r.cost = (int)s_sum; r.k = 0; r.iterations = 0; r.trivial = 1; r.cells = 0; res[J.pair] = r; }
            }
            continue;
        }
        // row 0: Q[0][j] = Q[0][j-1] + cost(s2[j], gap) + extgo(j); first step from (0,0) as in newkk_algn
        if (threadIdx.x == 0) {
            int q = 0;
            row0[0] = 0;
            for (int j = 1; j < len2; ++j) {
                const int b = s2[j], pb = s2[j - 1], add = s_cost[(b << 5) + 16];
                if (j == 1) q = add + ((b & 16) ? 0 : realgo);
                else q = q + add + (((pb & 16) && !(b & 16)) ? realgo : 0);
                row0[j] = q;
            }
        }
        __syncthreads();
        if (FINAL) {
            const int W = delta + 2 * J.k + 1;
            int *pl = W <= NK_SMEM_DIAGS ? s_dyn : gplanes;
            const int ws = W <= NK_SMEM_DIAGS ? NK_SMEM_DIAGS : ws_global;
            nk_fill<true>(s_cost, s1, s2, len1, len2, J.k, realgo, row0, pl, ws, dir + J.dir_off, J.wh);
            continue;
        }
        int T = (delta + 1) * delta_cost, iters = 0, k = 0, cost = 0;
        long long cells = 0;
        for (;;) {
            const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
            k = p >= len1 ? len1 - 1 : p;
            const int W = delta + 2 * k + 1;
            int *pl = W <= NK_SMEM_DIAGS ? s_dyn : gplanes;
            const int ws = W <= NK_SMEM_DIAGS ? NK_SMEM_DIAGS : ws_global;
            nk_fill<false>(s_cost, s1, s2, len1, len2, k, realgo, row0, pl, ws, nullptr, 0);
            ++iters;
            if (threadIdx.x == 0) {
                const unsigned g = ((unsigned *)(pl + 5 * (size_t)ws))[delta + k];
                const int gn = max((int)(g & 0xFFFFu), (int)(g >> 16));
                cost = pl[delta + k];
                s_stop = (p > gn || newp - len2 + 1 >= 0) ? 1 : 0;
            }
            __syncthreads();
            const int stop = s_stop;
            __syncthreads();
            if (stop) break;
            T *= 2;
        }
        if (threadIdx.x == 0) { NkResult r; r.cost = cost; r.k = k; r.iterations = iters; r.trivial = 0; r.cells = cells; res[J.pair] = r; }
    }
}

// backtrace_affine / trivial_backtrace: one thread per pair; rows are written right to left into the pair's slot
// [out_off, out_off + len1 + len2) exactly like my_prepend fills the capacity-(sz1+sz2) sequences of get_alignment
// (src/sequence.ml:1862-1877)
__global__ void k_newkk_traceback(const uint8_t *__restrict__ data, const NkJob *__restrict__ jobs, int njobs,
                                  const uint16_t *__restrict__ dir, const int64_t *__restrict__ out_off, uint8_t *r1, uint8_t *r2,
                                  int *out_len) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= njobs) return;
    const NkJob J = jobs[t];
    const uint8_t *s1 = data + J.off1, *s2 = data + J.off2;
    const int len1 = J.len1, len2 = J.len2, cap = len1 + len2;
    uint8_t *e1 = r1 + out_off[J.pair] + cap, *e2 = r2 + out_off[J.pair] + cap;
    int n = 0;
    if ((long long)len1 * NK_MUCH_LONGER < len2) {
        // prepended in reading order of the inputs, so both rows come out reversed (reference behaviour)
        for (int i = 0; i < len1; ++i) { ++n; e1[-n] = s1[i]; e2[-n] = 16; }
        for (int j = 0; j < len2; ++j) { ++n; e1[-n] = 16; e2[-n] = s2[j]; }
        out_len[2 * J.pair] = out_len[2 * J.pair + 1] = n;
        return;
    }
    const uint16_t *db = dir + J.dir_off;
    const int k = J.k, wh = J.wh, swaped = J.swaped;
    int i = len1 - 1, j = len2 - 1, mode = 0;     // 0 todo, 1 delete, 2 insert, 3 diagonal, 4 align
    while (i >= 0 && j >= 0) {
        const int dw = db[(size_t)(i + j) * wh + ((j - i + k) >> 1)];
        if (dw == 0) { ++n; e1[-n] = 16; e2[-n] = 16; --i; --j; continue; }
        if (mode == 0) {
            const int hi = dw & NK_DO_INSERT, hd = dw & NK_DO_DELETE, ha = dw & NK_ALIGN, hg = dw & NK_DO_DIAG;
            if (!swaped) mode = hd ? 1 : hi ? 2 : hg ? 3 : ha ? 4 : -1;
            else mode = hi ? 2 : hd ? 1 : hg ? 3 : ha ? 4 : -1;
            if (mode < 0) break;                   // "invalid dir": cannot happen for a filled cell
        } else if (mode == 1) {
            ++n; e1[-n] = s1[i]; e2[-n] = 16; --i;
            if (dw & (NK_END_DELETE | NK_INS_EQ_DEL)) mode = 0;
        } else if (mode == 2) {
            ++n; e1[-n] = 16; e2[-n] = s2[j]; --j;
            if (dw & (NK_END_INSERT | NK_INS_EQ_DEL)) mode = 0;
        } else if (mode == 3) {
            if (dw & NK_END_DIAG) mode = 0;
            ++n; e1[-n] = s1[i]; e2[-n] = s2[j]; --i; --j;
        } else {
            ++n; e1[-n] = s1[i] & 15; e2[-n] = s2[j] & 15; --i; --j; mode = 0;
        }
    }
    out_len[2 * J.pair] = out_len[2 * J.pair + 1] = n;
}

}  // namespace

// ---- batch twin of Sequence.NewkkAlign.align_2 / cost_2 (src/sequence.ml:1879-1990) --------------------------------
extern "C" poy_status poy_batch_newkk_align(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n, const int32_t *s1,
                                            const int32_t *s2, const uint8_t *swaped, const int64_t *out_off, int32_t *cost,
                                            uint8_t *r1, uint8_t *r2, int32_t *out_len, int32_t *stats) {
    bind_device(ctx);
    if (!ctx || !cm || !pool || n < 0 || (n > 0 && (!s1 || !s2))) return POY_ERR_ARG;
    if (n == 0) return POY_OK;
    if (cm->h.cost_model_type != 1)
        return poy_fail(ctx, POY_ERR_MODEL, "NewkkAlign: only the affine entry point is defined (the reference's non-affine newkkonen "
                                            "leaves costDiag unset: cost 0, traceback raises Failure)");
    const bool want_rows = r1 || r2 || out_len;
    if (want_rows && (!r1 || !r2 || !out_len || !out_off)) return poy_fail(ctx, POY_ERR_ARG, "r1, r2, out_len and out_off go together");
    std::vector<NkJob> hj((size_t)n);
    int maxlen2 = 1;
    int64_t out_total = 0;
    for (int p = 0; p < n; ++p) {
        const int a = s1[p], b = s2[p];
        if (a < 0 || a >= pool->nseq || b < 0 || b >= pool->nseq) return poy_fail(ctx, POY_ERR_ARG, "pair index out of range");
        NkJob &j = hj[p];
        j.off1 = pool->h_off[a]; j.off2 = pool->h_off[b];
        j.len1 = (int)(pool->h_off[a + 1] - j.off1); j.len2 = (int)(pool->h_off[b + 1] - j.off2);
        if (j.len1 > j.len2) return poy_fail(ctx, POY_ERR_ORDER, "ERROR: newkkonen.newkk_algn, s1 len > s2 len");
        if ((int64_t)j.len1 + j.len2 >= 65535) return poy_fail(ctx, POY_ERR_ARG, "NewkkAlign: gap counters are unsigned short (len1 + len2 < 65535)");
        j.swaped = swaped ? (swaped[p] ? 1 : 0) : 0; j.pair = p; j.k = 0; j.wh = 0; j.dir_off = 0;
        maxlen2 = std::max(maxlen2, j.len2);
        if (want_rows) out_total = std::max<int64_t>(out_total, out_off[p] + j.len1 + j.len2);
    }
    // my_add clamps at INT_MAX/2; plain sums stay far below it
    if ((int64_t)(cm->max_entry + 2 * (int64_t)std::max(cm->h.gap_open, 0)) * 2 * maxlen2 >= (1 << 29))
        return poy_fail(ctx, POY_ERR_COST_RANGE, "sequence lengths x costs can reach INT_MAX/2");
    const int blocks = std::min(n, ctx->sm_count * 2);
    const int ws_global = (2 * maxlen2 + 63) & ~63;              // W <= len1 + len2 - 1
    const size_t work_stride = (size_t)ws_global * 7;
    void *v_jobs, *v_work, *v_res, *v_misc;
    poy_status s;
    if ((s = poy_scratch(ctx, SL_JOBS, sizeof(NkJob) * (size_t)n, &v_jobs)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_WORK, sizeof(int) * work_stride * (size_t)blocks, &v_work)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_STATE, sizeof(NkResult) * (size_t)n + sizeof(int64_t) * (size_t)n + sizeof(int32_t) * 2 * (size_t)n, &v_res)) != POY_OK) return s;
    if ((s = poy_scratch(ctx, SL_MISC, 256, &v_misc)) != POY_OK) return s;
    NkJob *d_jobs = (NkJob *)v_jobs;
    NkResult *d_res = (NkResult *)v_res;
    int64_t *d_out_off = (int64_t *)(d_res + n);
    int32_t *d_out_len = (int32_t *)(d_out_off + n);
    int *d_counter = (int *)v_misc;
    static const size_t smem = sizeof(int) * 6 * NK_SMEM_DIAGS;
    CK(cudaFuncSetAttribute(k_newkk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_newkk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemcpyAsync(d_jobs, hj.data(), sizeof(NkJob) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream));
    k_newkk<false><<<blocks, NK_THREADS, smem, ctx->stream>>>(cm->d, pool->d_data, d_jobs, n, d_counter, (int *)v_work, work_stride,
                                                               ws_global, d_res, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<NkResult> hr((size_t)n);
    CK(cudaMemcpyAsync(hr.data(), d_res, sizeof(NkResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < n; ++p) {
        if (cost) cost[p] = hr[p].cost;
        if (stats) { stats[4 * p] = hr[p].iterations; stats[4 * p + 1] = 0; stats[4 * p + 2] = hr[p].k; stats[4 * p + 3] = hr[p].trivial; }
    }
    if (!want_rows) return POY_OK;
    // second pass: the final band again, with direction words, in waves that fit the arena; then the traceback
    uint8_t *d_r1, *d_r2;
    void *v_out;
    const size_t A = ((size_t)out_total + 255) & ~(size_t)255;
    if ((s = poy_scratch(ctx, SL_STORE2, 2 * A, &v_out)) != POY_OK) return s;
    d_r1 = (uint8_t *)v_out; d_r2 = d_r1 + A;
    CK(cudaMemcpyAsync(d_out_off, out_off, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int pos = 0;
    while (pos < n) {
        int64_t used = 0;
        int end = pos;
        while (end < n) {
            NkJob &j = hj[end];
            int64_t need = 0;
            if (!hr[end].trivial) {
                j.k = hr[end].k;
                const int W = (j.len2 - j.len1) + 2 * j.k + 1;
                j.wh = (W + 2) / 2;
                need = ((int64_t)(j.len1 + j.len2 - 1) * j.wh + 127) & ~127ll;
            }
            if (end > pos && (used + need) * 2 > (int64_t)ctx->arena_limit) break;
            j.dir_off = used;
            used += need;
            ++end;
        }
        void *v_dir;
        if ((s = poy_scratch(ctx, SL_DIR, (size_t)used * 2 + 256, &v_dir)) != POY_OK) return s;
        const int nj = end - pos;
        CK(cudaMemcpyAsync(d_jobs + pos, hj.data() + pos, sizeof(NkJob) * (size_t)nj, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream));
        k_newkk<true><<<std::min(nj, blocks), NK_THREADS, smem, ctx->stream>>>(cm->d, pool->d_data, d_jobs + pos, nj, d_counter, (int *)v_work,
                                                                              work_stride, ws_global, d_res, (uint16_t *)v_dir);
        ctx->launches++;
        CK(cudaGetLastError());
        k_newkk_traceback<<<(nj + 63) / 64, 64, 0, ctx->stream>>>(pool->d_data, d_jobs + pos, nj, (const uint16_t *)v_dir, d_out_off, d_r1, d_r2,
                                                                  d_out_len);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));      // the arena and hj are reused by the next wave
        pos = end;
    }
    CK(cudaMemcpyAsync(r1, d_r1, (size_t)out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(r2, d_r2, (size_t)out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, d_out_len, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POY_OK;
}
