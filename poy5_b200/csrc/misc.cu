// Small helper kernels: job construction, stop rule, gathers, and the INT32/DPX issue-rate
// micro-benchmark that supplies the roofline denominator.
#include "common.cuh"

// ---- cost-only job construction (device side, so pair lists may stay resident in HBM) -----------
// Rows are the shorter sequence (algn_CAML_cost_affine_3 swaps internally, src/algn.c:2496-2513).
// One list: general (4-state) pairs are appended from the front, gap-free pairs from the back, so the slower
// general pairs are started first and the gap-free ones fill the tail.
__global__ void k_build_cost_jobs(const int64_t *__restrict__ off, const uint8_t *__restrict__ gapfree, int n,
                                  const int *__restrict__ a, const int *__restrict__ b, CostJob *jobs, int *counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int sa = a[p], sb = b[p];
    const int64_t oa = off[sa], ob = off[sb];
    const int la = (int)(off[sa + 1] - oa), lb = (int)(off[sb + 1] - ob);
    CostJob j;
    if (la <= lb) { j.off_i = oa; j.lasti = la - 1; j.off_j = ob; j.lastj = lb - 1; }
    else { j.off_i = ob; j.lasti = lb - 1; j.off_j = oa; j.lastj = la - 1; }
    j.out = p;
    // tiny pairs take the exact flat-layout emulation inside the general kernel (cost_affine.cu, TINY_L)
    j.gapfree = gapfree[sa] && gapfree[sb] && (j.lastj + 1 > 8);
    if (j.gapfree) jobs[n - 1 - atomicAdd(counts, 1)] = j;
    else jobs[atomicAdd(counts + 1, 1)] = j;
}

cudaError_t launch_build_cost_jobs(poy_ctx *ctx, const poy_pool *pool, int n, const int *d_a, const int *d_b,
                                   CostJob *d_jobs, int *d_counts) {
    k_build_cost_jobs<<<(n + 255) / 256, 256, 0, ctx->stream>>>(pool->d_off, pool->d_gapfree, n, d_a, d_b, d_jobs, d_counts);
    ctx->launches++;
    return cudaGetLastError();
}

// ---- stop rule of algn_newkk_increaseT_aff (src/algn.c:2319-2335) -----------------------------------
__global__ void k_band_finish(const BandJob *__restrict__ jobs, int njobs, PairState *state, uint8_t *done,
                              const int *__restrict__ g0, int gap_open, const int *prog) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= njobs) return;
    const BandJob J = jobs[t];
    PairState *st = state + J.pair;
    if (J.swaped & 128) {
        // speculative fill: threshold from the job, no state carried on the device; 5 = the fill did not run
        const int pv = prog[J.pair];
        if (pv != SPEC_DONE && pv != SPEC_STOP) { done[J.pair] = pv == SPEC_ABORT ? 6 : 5; return; }
        const int delta = J.lastj - J.lasti, T = J.T;
        const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
        done[J.pair] = ((st->gapnum < p) || (newp - J.lastj + 1 >= 0)) ? 1 : 0;
        return;
    }
    int fin;
    if (J.lasti == 0) {
        // no rows: final_cost_matrix keeps the value initialize_matrices_affine gave it (src/algn.c:1879-1897);
        // the reference keeps doubling T with gap_num = 65535 until the band spans the matrix -- same result.
        st->cost = J.lastj >= 1 ? gap_open + g0[J.off_j + J.lastj] : 0;
        fin = 1;
    } else {
        const int delta = J.lastj - J.lasti, T = st->T;
        const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
        fin = (st->gapnum < p) || (newp - J.lastj + 1 >= 0);
    }
    st->iterations++;
    // verdict for the host: 1 = final, 2 = the stop rule fired on a probe fill (repeat this threshold with
    // direction bytes), 0 / 3 = double the threshold; 3 predicts that the next fill stops (gap_num already below its p)
    int verdict;
    if (fin) {
        verdict = ((J.swaped & 8) && !(J.swaped & 16) && J.lasti != 0) ? 2 : 1;
    } else {
        st->T *= 2;
        const int delta = J.lastj - J.lasti, T = st->T;
        const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
        verdict = ((st->gapnum < p) || (newp - J.lastj + 1 >= 0)) ? 3 : 0;
    }
    st->done = (verdict == 1);
    done[J.pair] = (uint8_t)verdict;
}

cudaError_t launch_band_finish(poy_ctx *ctx, const BandJob *d_jobs, int njobs, PairState *d_state, uint8_t *d_done,
                               const int *d_g0, int gap_open, const int *d_prog) {
    if (njobs <= 0) return cudaSuccess;
    k_band_finish<<<(njobs + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, njobs, d_state, d_done, d_g0, gap_open, d_prog);
    ctx->launches++;
    return cudaGetLastError();
}

// ---- stale state across threshold doublings (pairs with gap-bit symbols) -----------------------------------
// A probe fill changes the EB row / EH[0][0] the next fill inherits; if its threshold has to be repeated with
// direction bytes, the repeat must start from what the probe started from.  One warp per job: probe jobs save,
// repeat jobs restore, everything else returns.
__global__ void __launch_bounds__(256) k_stale_snapshot(const BandJob *__restrict__ jobs, int njobs, PairState *state, int *eb, int *snap) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= njobs) return;
    const BandJob J = jobs[t];
    if (J.swaped & 4) return;                       // gap-free pairs have no stale state
    const bool save = (J.swaped & 8) && !(J.swaped & 16), restore = (J.swaped & 32) != 0;
    if (!save && !restore) return;
    int *a = eb + J.eb_off, *b = snap + J.eb_off;
    PairState *st = state + J.pair;
    if (save) {
        for (int j = lane; j <= J.lastj; j += 32) b[j] = a[j];
        if (lane == 0) st->eh00_snap = st->eh00;
    } else {
        for (int j = lane; j <= J.lastj; j += 32) a[j] = b[j];
        if (lane == 0) st->eh00 = st->eh00_snap;
    }
}
cudaError_t launch_stale_snapshot(poy_ctx *ctx, const BandJob *d_jobs, int njobs, PairState *d_state, int *d_eb, int *d_eb_snap) {
    if (njobs <= 0) return cudaSuccess;
    k_stale_snapshot<<<(njobs + 7) / 8, 256, 0, ctx->stream>>>(d_jobs, njobs, d_state, d_eb, d_eb_snap);
    ctx->launches++;
    return cudaGetLastError();
}

__global__ void k_gather_cost(const PairState *__restrict__ state, int n, int *cost) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) cost[p] = state[p].cost;
}
cudaError_t launch_gather_cost(poy_ctx *ctx, const PairState *d_state, int n, int *d_cost) {
    k_gather_cost<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_state, n, d_cost);
    ctx->launches++;
    return cudaGetLastError();
}

__global__ void k_fill_int(int *d, int64_t n, int v) {
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < n; x += (int64_t)gridDim.x * blockDim.x) d[x] = v;
}
cudaError_t launch_fill_int(poy_ctx *ctx, int *d, int64_t n, int v) {
    if (n <= 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_fill_int<<<blocks, 256, 0, ctx->stream>>>(d, n, v);
    ctx->launches++;
    return cudaGetLastError();
}

// ---- INT32 / DPX issue-rate micro-benchmark --------------------------------------------------------------
// 8 independent dependency chains per thread, `iters` rounds, 1024 threads per SM x 2 CTAs.
template <int KIND>
__global__ void __launch_bounds__(512) k_microbench(int iters, int seed, unsigned long long *cycles, int *sink) {
    int x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = seed + threadIdx.x * 8 + c;
    const int y = seed | 1, z = seed ^ 0x55;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (KIND == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            else if (KIND == 1) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[c]) : "r"(z + c));
            else if (KIND == 2) x[c] = __viaddmin_s32(x[c], y, z + c + it);
            else if (KIND == 3) x[c] = __vimin3_s32(x[c], y + it, z + c);
            else {  // the gap-free cost-only cell: 3 adds, 2 add-min, 1 min3 per cell
                const int cb = x[c] + y;
                const int eh = __viaddmin_s32(cb, z, x[(c + 1) & 7]) + it;
                const int ev = __viaddmin_s32(x[(c + 2) & 7], z, cb) + c;
                x[c] = __vimin3_s32(cb, eh, ev);
            }
        }
    }
    const unsigned long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    int acc = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc ^= x[c];
    if (acc == 0x7fffffff) sink[threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = g1 - g0; }  // SM cycles, nanoseconds
}

cudaError_t launch_microbench(poy_ctx *ctx, int kind, int iters, unsigned long long *d_cycles, int *d_sink, float *ms,
                              double *ops) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = ctx->sm_count * 4, threads = 512;
    const int per_iter[5] = { 8, 8, 8, 8, 8 * 6 };
    for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
        cudaEventRecord(e0, ctx->stream);
        switch (kind) {
            case 0: k_microbench<0><<<blocks, threads, 0, ctx->stream>>>(iters, 12345, d_cycles, d_sink); break;
            case 1: k_microbench<1><<<blocks, threads, 0, ctx->stream>>>(iters, 12345, d_cycles, d_sink); break;
            case 2: k_microbench<2><<<blocks, threads, 0, ctx->stream>>>(iters, 12345, d_cycles, d_sink); break;
            case 3: k_microbench<3><<<blocks, threads, 0, ctx->stream>>>(iters, 12345, d_cycles, d_sink); break;
            default: k_microbench<4><<<blocks, threads, 0, ctx->stream>>>(iters, 12345, d_cycles, d_sink); break;
        }
        cudaEventRecord(e1, ctx->stream);
        ctx->launches++;
    }
    cudaError_t e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (kind < 0 || kind > 4) kind = 4;
    *ops = (double)blocks * threads * (double)iters * per_iter[kind];
    return e != cudaSuccess ? e : cudaGetLastError();
}
