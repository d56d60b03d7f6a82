// Cost-only affine DO alignment for gap-free pairs, second generation: k_cost_gf<C>.
// Batch twin of algn_CAML_cost_affine_3 -> algn_fill_plane_3_aff_nobt (src/algn.c:2457-2515,
// 1822-1863, 1987-2110) restricted to pairs in which no symbol carries the gap bit (all observed /
// leaf DNA): there the EB state can neither win nor tie a minimum (EB >= CB in every cell while table
// entries stay <= HIGH_NUM), go_i = go_j = GO, hext_j = ge_j, vext_i = ge_i, so a cell is
//     CB = min3(CB', EV', EH') + diag          (the min3 of the diagonal cell is carried as M)
//     EH = min(EH[j-1], CB[j-1] + GO) + ge_j
//     EV = min(EV[i-1], CB[i-1] + GO) + ge_i
//     M  = min3(CB, EH, EV)
// = 3 DPX instructions (VIADDMNMX x2, VIMNMX3), 3 adds and one table lookup.
//
// Mapping (see cost_affine.cu for the general scheme): one warp per pair, column blocks of 32*C
// columns, lane t owns C columns, skewed wavefront with neighbour exchange by __shfl_up.  New here:
//  * columns are RIGHT-aligned: column lastj is always slot C-1 of lane 31 of the last block, so the
//    result and the reference's even-row/last-column EV quirk touch one fixed register.
//    The padding on the left of block 0 consists of replicas of column 0 (ge = 0, table entry INF):
//    they reproduce EH = INF, EV[i][0] and M[i][0] exactly, and their CB stays >= INF;
//  * row parameters are not loaded per lane: the warp loads 32 packed rows at once every 32 steps,
//    lane 0 picks its row by shuffle and every row then travels down the lanes with the DP values;
//  * the 16x17 cost table is replicated once per shared-memory bank (conflict-free lookups).
#include "common.cuh"

#define GF_TAB_COLS 17                       // 16 symbols + 1 "padding" column whose entries are INF
#define GF_ROW_BYTES (GF_TAB_COLS * 128)     // bytes between consecutive table rows (32 replicas x 4 B)

template <int C>
__global__ void __launch_bounds__(128)
k_cost_gf(const DevCM *__restrict__ cm, const unsigned *__restrict__ rowpk, const int4 *__restrict__ colp,
          const int *__restrict__ g0v, const CostJob *__restrict__ jobs, const int *__restrict__ njobs_ptr, int *counter,
          int4 *bound, size_t bound_stride, int *__restrict__ cost_out) {
    constexpr int W = 32 * C;
    __shared__ int s_tab_i[16 * GF_TAB_COLS * 32];
    for (int x = threadIdx.x; x < 16 * GF_TAB_COLS * 32; x += blockDim.x) {
        const int e = x >> 5, a = e / GF_TAB_COLS, b = e % GF_TAB_COLS;
        s_tab_i[x] = b < 16 ? cm->cost16[a * 16 + b] : POY_INF;
    }
    __syncthreads();
    const char *s_tab = (const char *)s_tab_i;
    const int GO = cm->gap_open;
    const int njobs = *njobs_ptr;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int4 *bnd0 = bound + (size_t)warp_global * 2 * bound_stride;
    int4 *bnd1 = bnd0 + bound_stride;

    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= njobs) break;
        const CostJob J = jobs[job];
        const int lasti = J.lasti, lastj = J.lastj;
        const unsigned *rp = rowpk + J.off_i;
        const int4 *cp = colp + J.off_j;
        const int *g0 = g0v + J.off_j;
        if (lasti == 0) {  // no rows: minimum over row 0 at the last column (src/algn.c:2105-2109)
            if (lane == 0) cost_out[J.out] = lastj >= 1 ? min(GO + g0[lastj], POY_INF) : 0;
            continue;
        }
        const int nb = (lastj + W - 1) / W;
        const int pad = nb * W - lastj;
        for (int b = 0; b < nb; ++b) {
            const int jb = b * W + lane * C - pad;  // slot c <-> column jb + c + 1 (<= 0: replica of column 0)
            const int4 *bin = (b & 1) ? bnd0 : bnd1;
            int4 *bout = (b & 1) ? bnd1 : bnd0;
            const bool last_block = (b == nb - 1);
            int c_ge[C], c_off[C], CBu[C], EVu[C], Mu[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jb + c + 1;
                if (j >= 1) {
                    const int4 v = cp[j];
                    c_ge[c] = v.x;
                    c_off[c] = ((v.w & 15) << 7) + (lane << 2);
                    CBu[c] = POY_INF; EVu[c] = POY_INF;
                    Mu[c] = min(GO + g0[j], POY_INF);       // min3(INF, INF, EH[0][j])
                } else {
                    c_ge[c] = 0;
                    c_off[c] = (16 << 7) + (lane << 2);
                    CBu[c] = 0; EVu[c] = GO;                  // CB[0][0], EV[0][0]
                    Mu[c] = min(0, GO);
                }
            }
            // cell (0, jb): row-0 neighbour to the left of slot 0 (diagonal predecessor of row 1)
            int dM;
            if (jb >= 1) dM = min(GO + g0[jb], POY_INF); else dM = min(0, GO);
            int ev_col0 = GO;                                 // EV[i][0] = GO + sum ge_r (src/algn.c:2066-2070)
            int oCB = POY_INF, oEH = POY_INF, oM = POY_INF;
            unsigned rk = 0, win = 0;
            int4 bnext = make_int4(0, 0, 0, 0);
            if (b > 0 && lane == 0) bnext = bin[1];

            const int nsteps = lasti + 31;
            for (int s = 0; s < nsteps; ++s) {
                if ((s & 31) == 0) {                          // next 32 packed rows, one coalesced load
                    const int r = s + lane + 1;
                    win = rp[r <= lasti ? r : lasti];
                }
                const int i = s - lane + 1;
                int lCB = __shfl_up_sync(0xffffffffu, oCB, 1);
                int lEH = __shfl_up_sync(0xffffffffu, oEH, 1);
                int lM = __shfl_up_sync(0xffffffffu, oM, 1);
                unsigned rprev = __shfl_up_sync(0xffffffffu, rk, 1);
                const unsigned rfirst = __shfl_sync(0xffffffffu, win, s & 31);
                rk = lane == 0 ? rfirst : rprev;              // row i's parameters travel down the lanes
                if (i >= 1) {
                    const int ge_i = (int)(rk >> 16);
                    const char *rowbase = s_tab + (rk & 0xFFFFu);
                    if (lane == 0) {
                        if (b == 0) {
                            ev_col0 += ge_i;
                            lCB = POY_INF; lEH = POY_INF; lM = ev_col0;
                        } else {
                            lCB = bnext.x; lEH = bnext.y; lM = bnext.z;
                            bnext = bin[i < lasti ? i + 1 : lasti];
                        }
                    }
                    int cbL = lCB, ehL = lEH, mD = dM;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int diag = *(const int *)(rowbase + c_off[c]);
                        const int cb = mD + diag;
                        const int eh = __viaddmin_s32(cbL, GO, ehL) + c_ge[c];
                        const int ev = __viaddmin_s32(CBu[c], GO, EVu[c]) + ge_i;
                        mD = Mu[c];
                        Mu[c] = __vimin3_s32(cb, eh, ev);
                        CBu[c] = cb; EVu[c] = ev;
                        cbL = cb; ehL = eh;
                    }
                    // F5: EV at the last column of an even row comes from clobbered predecessors
                    if (last_block && lane == 31 && !(i & 1)) EVu[C - 1] = POY_INF + ge_i;
                    dM = lM;
                    oCB = cbL; oEH = ehL; oM = Mu[C - 1];
                    if (lane == 31 && !last_block && i <= lasti) bout[i] = make_int4(oCB, oEH, oM, 0);
                }
            }
            if (last_block && lane == 31)                    // lane 31 finished row lasti in the last step
                cost_out[J.out] = __vimin3_s32(oCB, oEH, EVu[C - 1]);
            __syncwarp();
        }
    }
}

cudaError_t launch_cost_gf(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const CostJob *d_jobs,
                           const int *d_count, int *d_counter, int4 *d_bound, size_t bound_stride, int blocks, int *d_cost) {
    k_cost_gf<16><<<blocks, 128, 0, ctx->stream>>>(cm->d, pool->d_rowpk, pool->d_colp, pool->d_g0, d_jobs, d_count, d_counter,
                                                   d_bound, bound_stride, d_cost);
    ctx->launches++;
    return cudaGetLastError();
}
