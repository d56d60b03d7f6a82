// Linear-gap DO alignment on the GPU: batch twin of algn_CAML_simple_2 / algn_CAML_align_2d
// (src/algn.c:3134, 3500) = algn_nw -> algn_fill_plane_2 (:2954-2975, 1134-1177), i.e. the full
// plane algn_fill_plane (:927-973) or the Ukkonen band algn_newkk_increaseT / algn_newkk_test /
// algn_newkk_fill_a_row (:1008-1130) with the row kernels algn_fill_row, algn_fill_ukk_*_cell,
// algn_fill_last_column, algn_fill_first_cell (:458-730), and backtrace_2d (:3277-3327).
//
// One state per cell (the reference's `mm`), three candidates:
//   DELETE = up + cost(s1_i, gap)      (absent on the right border; the last column of a row without
//                                       right border additionally offers up + tail_cost[s1_i])
//   INSERT = left + cost(gap, s2_j)    (absent on the left border)
//   ALIGN  = diag + cost(s1_i, s2_j)
// every minimal candidate is recorded (3-bit mask).  Column 0: up + tail_cost[s1_i] in the band
// (alg_row[0] = tail) or up + cost(s1_i, gap) in the full plane.  Row 0: prefix sums of
// prepend_cost (the pool's g0 row).
//
// The band's stop rule needs max(#insertions, #deletions) of the ALIGN > INSERT > DELETE traceback
// (backtrace_2d_gaps, :978-1005).  The traceback follows, from every cell, that cell's own preferred
// predecessor, so the two counts obey a forward recurrence and are carried through the fill as a
// packed 16x2 register like the affine gap counters -- no per-iteration traceback is needed.
//
// The machine mapping is that of band2.cu: anti-diagonal wavefront, thread t owns D diagonals in
// registers, NW warps cooperate on one pair through shared memory, direction bytes are stored
// anti-diagonal major.  The full plane is the band with k = lenX-1 (no borders).  Direction byte:
// bits 0-1 = move of the real traceback (0 align, 1 insert, 2 delete) under the per-pair `swaped`
// priority (follow_insertion_or_deletion, :108-124).
#include <type_traits>
#include "common.cuh"

namespace {

template <int N, class F>
__device__ __forceinline__ void sfor(F &&f) {
    if constexpr (N > 0) {
        sfor<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

struct LRow { int cdel, tail, off; };  // cost(s1_i,gap), tail_cost[s1_i], byte offset of the cost32 row
struct LCol { int gapc, off; };        // cost(gap,s2_j), byte offset of the column

#define LIN_FULLPLANE 1

// returns the direction byte; MM / G hold the diagonal predecessor on entry, the new cell on exit
__device__ __forceinline__ unsigned lin_cell(int &MM, unsigned &G, int lMM, unsigned lG, int uMM, unsigned uG, const LRow r,
                                             const LCol c, const char *s_tab, bool lb, bool rb, bool jzero, bool lastcol,
                                             bool fullplane, bool swaped) {
    const int alg = *(const int *)(s_tab + r.off + c.off);
    int t1 = uMM + r.cdel;
    if (lastcol && !rb) t1 = min(t1, uMM + r.tail);
    const int t2 = lMM + c.gapc;
    const int t3 = MM + alg;
    int m;
    bool eD, eI, eA;
    if (jzero) {                       // first cell of a row that starts at column 0
        m = uMM + (fullplane ? r.cdel : r.tail);
        eD = true; eI = false; eA = false;
    } else {
        const int big = 0x7fffffff;
        const int c1 = rb ? big : t1, c2 = lb ? big : t2;
        m = min(min(c1, c2), t3);
        eD = (c1 == m); eI = (c2 == m); eA = (t3 == m);
    }
    // gap counts of the ALIGN > INSERT > DELETE traceback
    G = eA ? G : (eI ? lG + 1u : uG + 0x10000u);
    MM = m;
    unsigned mv;
    if (eA) mv = 0u;
    else if (swaped) mv = eI ? 1u : 2u;
    else mv = eD ? 2u : 1u;
    return mv;
}

template <int H>
__device__ __forceinline__ void store_dir(uint8_t *p, unsigned long long packed) {
    if (H == 1) *p = (uint8_t)packed;
    else if (H == 2) *(uint16_t *)p = (uint16_t)packed;
    else if (H == 4) *(uint32_t *)p = (uint32_t)packed;
    else *(unsigned long long *)p = packed;
}

}  // namespace

template <int D, int NW, int WPB>
__global__ void __launch_bounds__(WPB * 32)
k_band_lin(const DevCM *__restrict__ cm, const uint8_t *__restrict__ data, const int *__restrict__ g0v,
           const BandJob *__restrict__ jobs, int njobs, int *counter, PairState *state, uint8_t *dir) {
    constexpr int H = D / 2;
    static_assert(NW == 1 || WPB == NW, "cooperating warps fill the whole CTA");
    __shared__ int s_cost[1024];
    __shared__ int s_tailv[32];
    __shared__ int s_job;
    __shared__ int s_xe[NW][2], s_xo[NW][2];
    const int lane = threadIdx.x & 31, warp = (NW == 1) ? 0 : (threadIdx.x >> 5);
    const int tid = (NW == 1) ? lane : (int)threadIdx.x;
    for (int x = threadIdx.x; x < 1024; x += WPB * 32) s_cost[x] = cm->cost32[x];
    if (threadIdx.x < 32) s_tailv[threadIdx.x] = cm->tail[threadIdx.x];
    __syncthreads();
    const char *s_tab = (const char *)s_cost;

    for (;;) {
        int job;
        if (NW == 1) {
            job = 0;
            if (lane == 0) job = atomicAdd(counter, 1);
            job = __shfl_sync(0xffffffffu, job, 0);
        } else {
            __syncthreads();
            if (tid == 0) s_job = atomicAdd(counter, 1);
            __syncthreads();
            job = s_job;
        }
        if (job >= njobs) break;
        const BandJob J = jobs[job];
        const int lasti = J.lasti, lastj = J.lastj, k = J.k;
        const bool swaped = (J.swaped & 1) != 0, fullplane = (J.swaped & 2) != 0;
        if (lasti == 0) continue;
        const int delta = lastj - lasti, B = delta + 2 * k + 1;
        const uint8_t *s1 = data + J.off_i, *s2 = data + J.off_j;
        const int *g0 = g0v + J.off_j;
        PairState *st = state + J.pair;
        uint8_t *dbase = dir + J.dir_off;
        const int stride = J.stride;
        const int d0 = tid * D;
        const int rbslot = fullplane ? -1 : (B - 1) - d0;

        int MM[D];
        unsigned G[D];
        sfor<D>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int d = d0 + u, j0 = d - k;
            if (d < B && j0 >= 0 && j0 <= lastj) { MM[u] = g0[j0]; G[u] = (unsigned)j0 & 0xFFFFu; }  // row 0: j0 insertions
            else { MM[u] = POY_INF; G[u] = 0u; }
        });

        int a = k & 1;
        int i0 = (a - d0 + k) >> 1, j0 = a - i0;
        LRow R[H];
        LCol C[H + 1];
        auto load_row = [&](int i) {
            i = i < 0 ? 0 : (i > lasti ? lasti : i);
            const int sym = s1[i] & 31;
            LRow e; e.cdel = s_cost[(sym << 5) + POY_GAP]; e.tail = s_tailv[sym]; e.off = sym << 7;
            return e;
        };
        auto load_col = [&](int j) {
            j = j < 0 ? 0 : (j > lastj ? lastj : j);
            const int sym = s2[j] & 31;
            LCol e; e.gapc = s_cost[(POY_GAP << 5) + sym]; e.off = sym << 2;
            return e;
        };
        sfor<H>([&](auto hc) { constexpr int h = decltype(hc)::value; R[h] = load_row(i0 - h); });
        sfor<H + 1>([&](auto hc) { constexpr int h = decltype(hc)::value; C[h] = load_col(j0 + h); });

        if (NW > 1) {
            if (lane == 0) { s_xe[warp][0] = MM[0]; s_xe[warp][1] = (int)G[0]; }
            if (lane == 31) { s_xo[warp][0] = MM[D - 1]; s_xo[warp][1] = (int)G[D - 1]; }
            __syncthreads();
        } else {
            __syncwarp();
        }

        const int a_end = lasti + lastj;
        int a_main = delta + k + 2;
        if ((a_main ^ a) & 1) ++a_main;
        const bool warp_in_band = (NW == 1) || (warp * 32 * D < B);

        auto iteration = [&](auto edge_c) {
            constexpr bool EDGE = decltype(edge_c)::value;
            uint8_t nsym1 = 0, nsym2 = 0;
            if (warp_in_band) {
                int ri = i0 + 1, cj = j0 + 1 + H;
                ri = ri < 0 ? 0 : (ri > lasti ? lasti : ri);
                cj = cj < 0 ? 0 : (cj > lastj ? lastj : cj);
                nsym1 = s1[ri]; nsym2 = s2[cj];
            }
            {   // even diagonals
                int sMM = __shfl_up_sync(0xffffffffu, MM[D - 1], 1);
                unsigned sG = __shfl_up_sync(0xffffffffu, G[D - 1], 1);
                if (NW > 1 && lane == 0 && warp > 0) { sMM = s_xo[warp - 1][0]; sG = (unsigned)s_xo[warp - 1][1]; }
                unsigned long long packed = 0;
                if (warp_in_band)
                sfor<H>([&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr int u = 2 * h;
                    const int i = i0 - h, j = j0 + h, d = d0 + u;
                    bool valid = true;
                    if (EDGE) valid = (d < B) && (i >= 1) && (i <= lasti) && (j >= 0) && (j <= lastj);
                    if (valid) {
                        int lMM; unsigned lG;
                        if constexpr (u == 0) { lMM = sMM; lG = sG; }
                        else { lMM = MM[u > 0 ? u - 1 : 0]; lG = G[u > 0 ? u - 1 : 0]; }
                        bool lb = false;
                        if constexpr (u == 0) lb = (tid == 0) && !fullplane;
                        const unsigned b = lin_cell(MM[u], G[u], lMM, lG, MM[u + 1], G[u + 1], R[h], C[h], s_tab, lb, u == rbslot,
                                                    EDGE && j == 0, j == lastj, fullplane, swaped);
                        packed |= (unsigned long long)b << (8 * h);
                    }
                });
                if (warp_in_band) store_dir<H>(dbase + (size_t)a * stride + tid * H, packed);
                if (NW > 1) {
                    if (lane == 0) { s_xe[warp][0] = MM[0]; s_xe[warp][1] = (int)G[0]; }
                    __syncthreads();
                }
            }
            {   // odd diagonals
                int sMM = __shfl_down_sync(0xffffffffu, MM[0], 1);
                unsigned sG = __shfl_down_sync(0xffffffffu, G[0], 1);
                if (NW > 1 && lane == 31 && warp < NW - 1) { sMM = s_xe[warp + 1][0]; sG = (unsigned)s_xe[warp + 1][1]; }
                unsigned long long packed = 0;
                if (warp_in_band)
                sfor<H>([&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr int u = 2 * h + 1;
                    const int i = i0 - h, j = j0 + h + 1, d = d0 + u;
                    bool valid = true;
                    if (EDGE) valid = (d < B) && (i >= 1) && (i <= lasti) && (j >= 0) && (j <= lastj);
                    if (valid) {
                        int uMM; unsigned uG;
                        if constexpr (u == D - 1) { uMM = sMM; uG = sG; }
                        else { constexpr int uu = u < D - 1 ? u + 1 : u; uMM = MM[uu]; uG = G[uu]; }
                        const unsigned b = lin_cell(MM[u], G[u], MM[u - 1], G[u - 1], uMM, uG, R[h], C[h + 1], s_tab, false, u == rbslot,
                                                    EDGE && j == 0, j == lastj, fullplane, swaped);
                        packed |= (unsigned long long)b << (8 * h);
                    }
                });
                if (warp_in_band) store_dir<H>(dbase + (size_t)(a + 1) * stride + tid * H, packed);
                if (NW > 1) {
                    if (lane == 31) { s_xo[warp][0] = MM[D - 1]; s_xo[warp][1] = (int)G[D - 1]; }
                    __syncthreads();
                }
            }
            sfor<H - 1>([&](auto hc) { constexpr int h = H - 1 - decltype(hc)::value; R[h] = R[h - 1]; });
            sfor<H>([&](auto hc) { constexpr int h = decltype(hc)::value; C[h] = C[h + 1]; });
            ++i0; ++j0;
            {
                const int a1 = nsym1 & 31, a2 = nsym2 & 31;
                R[0].cdel = s_cost[(a1 << 5) + POY_GAP]; R[0].tail = s_tailv[a1]; R[0].off = a1 << 7;
                C[H].gapc = s_cost[(POY_GAP << 5) + a2]; C[H].off = a2 << 2;
            }
        };

        for (; a <= a_end && a < a_main; a += 2) iteration(std::true_type{});
        for (; a <= a_end; a += 2) iteration(std::false_type{});

        const int dstar = delta + k;
        if (tid == dstar / D) {
            sfor<D>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                if (u == dstar % D) {
                    st->cost = MM[u];
                    st->gapnum = max((int)(G[u] & 0xFFFFu), (int)(G[u] >> 16));
                }
            });
        }
    }
}

// ---- generic fallback (any band width): one CTA per pair, diagonal state in global memory -----------
__global__ void __launch_bounds__(256)
k_band_lin_generic(const DevCM *__restrict__ cm, const uint8_t *__restrict__ data, const int *__restrict__ g0v,
                   const BandJob *__restrict__ jobs, int njobs, PairState *state, uint8_t *dir, int *work, size_t work_stride) {
    __shared__ int s_cost[1024];
    __shared__ int s_tailv[32];
    for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_cost[x] = cm->cost32[x];
    if (threadIdx.x < 32) s_tailv[threadIdx.x] = cm->tail[threadIdx.x];
    __syncthreads();
    const char *s_tab = (const char *)s_cost;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const BandJob J = jobs[job];
        const int lasti = J.lasti, lastj = J.lastj, k = J.k;
        const bool swaped = (J.swaped & 1) != 0, fullplane = (J.swaped & 2) != 0;
        if (lasti == 0) continue;
        const int delta = lastj - lasti, B = delta + 2 * k + 1;
        const uint8_t *s1 = data + J.off_i, *s2 = data + J.off_j;
        const int *g0 = g0v + J.off_j;
        PairState *st = state + J.pair;
        uint8_t *dbase = dir + J.dir_off;
        const int stride = J.stride;
        const size_t ws = work_stride / 2;
        int *wMM = work + (size_t)blockIdx.x * work_stride;
        unsigned *wG = (unsigned *)(wMM + ws);
        for (int d = threadIdx.x; d < B; d += blockDim.x) {
            const int j0 = d - k;
            if (j0 >= 0 && j0 <= lastj) { wMM[d] = g0[j0]; wG[d] = (unsigned)j0 & 0xFFFFu; }
            else { wMM[d] = POY_INF; wG[d] = 0u; }
        }
        __syncthreads();
        const int a_end = lasti + lastj;
        for (int a = 1; a <= a_end; ++a) {
            const int par = (a + k) & 1;
            for (int d = 2 * threadIdx.x + par; d < B; d += 2 * blockDim.x) {
                const int i = (a - d + k) >> 1, j = a - i;
                if (i >= 1 && i <= lasti && j >= 0 && j <= lastj) {
                    const int a1 = s1[i] & 31, a2 = s2[j] & 31;
                    LRow r; r.cdel = s_cost[(a1 << 5) + POY_GAP]; r.tail = s_tailv[a1]; r.off = a1 << 7;
                    LCol c; c.gapc = s_cost[(POY_GAP << 5) + a2]; c.off = a2 << 2;
                    int mm = wMM[d]; unsigned g = wG[d];
                    const int lMM = d > 0 ? wMM[d - 1] : POY_INF; const unsigned lG = d > 0 ? wG[d - 1] : 0u;
                    const int uMM = d + 1 < B ? wMM[d + 1] : POY_INF; const unsigned uG = d + 1 < B ? wG[d + 1] : 0u;
                    const unsigned b = lin_cell(mm, g, lMM, lG, uMM, uG, r, c, s_tab, !fullplane && d == 0, !fullplane && d == B - 1,
                                                j == 0, j == lastj, fullplane, swaped);
                    wMM[d] = mm; wG[d] = g;
                    dbase[(size_t)a * stride + (d >> 1)] = (uint8_t)b;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int d = delta + k;
            st->cost = wMM[d];
            st->gapnum = max((int)(wG[d] & 0xFFFFu), (int)(wG[d] >> 16));
        }
        __syncthreads();
    }
}

// ---- stop rule of algn_newkk_increaseT (src/algn.c:1117-1130); lenX = lasti+1, lenY = lastj+1 ---------------
__global__ void k_lin_finish(const BandJob *__restrict__ jobs, int njobs, PairState *state, uint8_t *done,
                             const int *__restrict__ g0) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= njobs) return;
    const BandJob J = jobs[t];
    PairState *st = state + J.pair;
    int fin;
    if (J.lasti == 0) {                 // no rows: the cost is the end of row 0
        st->cost = g0[J.off_j + J.lastj];
        fin = 1;
    } else if (J.swaped & 2) {          // full plane: a single fill
        fin = 1;
    } else {
        const int delta = J.lastj - J.lasti, T = st->T, lenY = J.lastj + 1;
        const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
        fin = ((st->gapnum + 1) < p) || (newp - lenY + 1 >= 0);
    }
    st->iterations++;
    if (!fin) st->T *= 2;
    st->done = fin;
    done[J.pair] = (uint8_t)fin;
}

// ---- backtrace_2d (src/algn.c:3277-3327): warp per pair, shared-memory tile of direction bytes --------------
#define TB_ROWS 64
#define TB_COLS 32
__global__ void __launch_bounds__(128)
k_traceback_lin(const uint8_t *__restrict__ data, const BandJob *__restrict__ jobs, int njobs, const uint8_t *__restrict__ done,
                const uint8_t *__restrict__ dir, const int64_t *__restrict__ out_off, uint8_t *r1, uint8_t *r2, int *out_len) {
    __shared__ __align__(16) uint8_t s_tile[4][TB_ROWS * TB_COLS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 4 + w;
    if (t >= njobs) return;
    const BandJob J = jobs[t];
    if (done && !done[J.pair]) return;
    const uint8_t *s1 = data + J.off_i, *s2 = data + J.off_j;
    const int k = J.k;
    const int B = (J.lastj - J.lasti) + 2 * k + 1;
    const uint8_t *db = dir + J.dir_off;
    const int stride = J.stride;
    const int cap = J.lasti + J.lastj + 2;  // len1 + len2
    const int64_t base = out_off ? out_off[J.pair] : 0;
    uint8_t *p1 = r1 ? r1 + base + cap : nullptr;
    uint8_t *p2 = r2 ? r2 + base + cap : nullptr;
    int n = 0;
#define PUT2(a, b) do { ++n; if (p1) *(--p1) = (uint8_t)(a); if (p2) *(--p2) = (uint8_t)(b); } while (0)
    int i = J.lasti, j = J.lastj;
    uint8_t *tile = s_tile[w];
    for (;;) {
        i = __shfl_sync(0xffffffffu, i, 0);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (i == 0 || j == 0) break;
        const int a_hi = i + j;
        int dcur = j - i + k;
        dcur = dcur < 0 ? 0 : (dcur >= B ? B - 1 : dcur);
        int c0 = ((dcur >> 1) - 12) & ~15;
        if (c0 > stride - TB_COLS) c0 = stride - TB_COLS;
        if (c0 < 0) c0 = 0;
        for (int q = lane; q < TB_ROWS * 2; q += 32) {
            const int r = q >> 1, half = q & 1;
            const int a = a_hi - r;
            if (a >= 0 && c0 + half * 16 + 16 <= stride)
                *(uint4 *)(tile + r * TB_COLS + half * 16) = *(const uint4 *)(db + (size_t)a * stride + c0 + half * 16);
        }
        __syncwarp();
        if (lane == 0) {
            while (i != 0 && j != 0) {
                const int r = a_hi - (i + j);
                int d = j - i + k;
                d = d < 0 ? 0 : (d >= B ? B - 1 : d);
                const int cb = (d >> 1) - c0;
                if (r >= TB_ROWS || cb < 0 || cb >= TB_COLS) break;
                const unsigned mv = tile[r * TB_COLS + cb] & 3u;
                if (mv == 0) { PUT2(s1[i], s2[j]); --i; --j; }
                else if (mv == 1) { PUT2(POY_GAP, s2[j]); --j; }
                else { PUT2(s1[i], POY_GAP); --i; }
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        // row 0 is all INSERT, column 0 all DELETE, and cell (0,0) aligns the two leading gaps
        while (j != 0) { PUT2(POY_GAP, s2[j]); --j; }
        while (i != 0) { PUT2(s1[i], POY_GAP); --i; }
        PUT2(s1[0], s2[0]);
        if (out_len) { out_len[2 * J.pair + 0] = n; out_len[2 * J.pair + 1] = n; }
    }
#undef PUT2
}

template <int D, int NW>
static cudaError_t launch_lin_one(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                  int *d_counter, PairState *d_state, uint8_t *d_dir) {
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
    if (e != cudaSuccess) return e;
    constexpr int WPB = NW == 1 ? 8 : NW;
    const int groups_per_block = WPB / NW;
    int blocks = (njobs + groups_per_block - 1) / groups_per_block;
    const int cap = ctx->sm_count * 6;
    if (blocks > cap) blocks = cap;
    k_band_lin<D, NW, WPB><<<blocks, WPB * 32, 0, ctx->stream>>>(cm->d, pool->d_data, pool->d_g0, d_jobs, njobs, d_counter, d_state, d_dir);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_band_lin(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs, int cls,
                            int *d_counter, PairState *d_state, uint8_t *d_dir) {
    if (njobs <= 0) return cudaSuccess;
#define LL(DD, WW) return launch_lin_one<DD, WW>(ctx, cm, pool, d_jobs, njobs, d_counter, d_state, d_dir)
    switch (cls) {
        case 64: LL(2, 1);
        case 128: LL(4, 1);
        case 256: LL(8, 1);
        case 512: LL(8, 2);
        case 768:
        case 1024: LL(8, 4);
        case 1280:
        case 1536:
        case 2048: LL(8, 8);
        case 4096: LL(8, 16);
    }
#undef LL
    return cudaErrorInvalidValue;
}

cudaError_t launch_band_lin_generic(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                    PairState *d_state, uint8_t *d_dir, int *d_work, size_t work_stride, int blocks) {
    if (njobs <= 0) return cudaSuccess;
    k_band_lin_generic<<<blocks, 256, 0, ctx->stream>>>(cm->d, pool->d_data, pool->d_g0, d_jobs, njobs, d_state, d_dir, d_work, work_stride);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_lin_finish(poy_ctx *ctx, const poy_pool *pool, const BandJob *d_jobs, int njobs, PairState *d_state, uint8_t *d_done) {
    if (njobs <= 0) return cudaSuccess;
    k_lin_finish<<<(njobs + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, njobs, d_state, d_done, pool->d_g0);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_traceback_lin(poy_ctx *ctx, const poy_pool *pool, const BandJob *d_jobs, int njobs, const uint8_t *d_done,
                                 const uint8_t *d_dir, const int64_t *d_out_off, uint8_t *d_r1, uint8_t *d_r2, int *d_out_len) {
    if (njobs <= 0) return cudaSuccess;
    k_traceback_lin<<<(njobs + 3) / 4, 128, 0, ctx->stream>>>(pool->d_data, d_jobs, njobs, d_done, d_dir, d_out_off, d_r1, d_r2, d_out_len);
    ctx->launches++;
    return cudaGetLastError();
}
