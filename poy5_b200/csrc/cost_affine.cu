// Batched cost-only affine DO alignment: the batch twin of algn_CAML_cost_affine_3 ->
// algn_fill_plane_3_aff_nobt (src/algn.c:2457-2515, 1822-1863, 1987-2110).
//
// One WARP per pair, persistent kernel, jobs pulled from one list.  The full leni x lenj plane is swept in
// column blocks of W = 32*C columns; inside a block lane t owns C adjacent columns and keeps the previous row of
// its states in registers.  Sequences wider than W hand one boundary column (16 bytes per row) from block to
// block through a per-warp scratch in global memory.
//
// Recurrences per cell (i,j), ' = cell (i-1,j-1)   (src/algn.c:1260-1433):
//   EH = min(EH[i][j-1] + hext_j, CB[i][j-1] + go_j + ge_j)
//   EV = min(EV[i-1][j] + vext_i, CB[i-1][j] + go_i + ge_i)
//   EB = min(EB' + (both gaps ? 0 : INF), CB' + (both ? (clean ? 0 : 2GO) : INF))
//   CB = diag + min(CB', EV' + [ic has gap]go_j, EH' + [jc has gap]go_i, EB' + max(go_i,go_j))
// evaluated with the DPX fused add-min instructions (__viaddmin_s32, __vimin3_s32).
//
// Two code paths:
//  * gap-free pairs (neither sequence contains a gap-bit symbol: all leaf / observed DNA) run a SKEWED wavefront
//    (at step s lane t computes row s - t + 1, the left neighbour's value arrives one step later by __shfl_up),
//    drop the EB state -- provably EB >= CB in every cell, so it never wins a minimum -- fold min3(CB,EV,EH) of
//    the diagonal cell into one carried value M and work in a shifted domain that moves both gap extensions
//    into the cost table: 3 DPX instructions, 1 add and one table lookup per cell (cost_pair_gf);
//  * pairs with gap-bit symbols keep all four states and go ROW BY ROW, all lanes on the same row, so that the
//    rows without a gap bit (most of them) can skip EB and the flag logic warp-uniformly (cost_pair_rows).
//
// The reference's row-buffer aliasing is reproduced in closed
// form: for lenj >= 3 its only observable effect is that on every even row i>=2
// EV[i][lastj] is computed from clobbered predecessors (both INF), i.e.
// EV[i][lastj] = INF + min(vext_i, go_i+ge_i).  Pairs with lenj <= TINY_L are run
// through an exact emulation of the reference's flat scratch layout instead.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

#define TINY_L 8
#define GF_TAB_COLS 17                       // 16 symbols + 1 "padding" column whose entries are INF

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// compile-time loop N-1 .. 0 (register arrays stay scalar-replaced)
template <int N, class F>
__device__ __forceinline__ void sfor_down(F &&f) {
    if constexpr (N > 0) {
        f(std::integral_constant<int, N - 1>{});
        sfor_down<N - 1>(f);
    }
}

// ---- exact emulation for tiny pairs (lane 0 only) -------------------------------
__device__ int cost_affine_tiny(const DevCM *cm, const int *s_cost16, const int4 *rp, const int4 *cp, const int *g0,
                                int lasti, int lastj) {
    const int L = lastj + 1, stride = lastj + 2, go = cm->gap_open;
    int M[9 * TINY_L + 16];  // CB, EB, EV, EH row pairs at 0, 2L, 4L, 6L (src/algn.c:2489-2492)
    for (int x = 0; x < 9 * TINY_L + 16; ++x) M[x] = 0;
    int *cb0 = M, *eb0 = M + 2 * L, *ev0 = M + 4 * L, *eh0 = M + 6 * L;
    cb0[0] = 0; eb0[0] = POY_INF; eh0[0] = go; ev0[0] = go;
    for (int j = 1; j <= lastj; ++j) {
        eh0[j] = go + g0[j];
        cb0[j] = POY_INF; eb0[j] = POY_INF; ev0[j] = POY_INF;
    }
    for (int i = 1; i <= lasti; ++i) {
        const int cur = (i & 1) * stride, prv = ((i & 1) ^ 1) * stride;
        int *CB = cb0 + cur, *EB = eb0 + cur, *EV = ev0 + cur, *EH = eh0 + cur;
        const int *pCB = cb0 + prv, *pEB = eb0 + prv, *pEV = ev0 + prv, *pEH = eh0 + prv;
        const int4 r = rp[i];
        EH[0] = POY_INF;
        const int r0 = pEV[0] + r.x;
        EH[0] = POY_INF; CB[0] = r0; EB[0] = POY_INF; EV[0] = r0; CB[0] = POY_INF;
        for (int j = 1; j <= lastj; ++j) {
            const int4 c = cp[j];
            int ext = EH[j - 1] + c.x, opn = CB[j - 1] + c.y;
            EH[j] = ext < opn ? ext : opn;
            ext = pEV[j] + r.x; opn = pCB[j] + r.y;
            EV[j] = ext < opn ? ext : opn;
            const bool both = (r.w & c.w & PF_HASGAP) != 0, clean = ((r.w | c.w) & PF_PREVGAP) == 0;
            ext = pEB[j - 1] + (both ? 0 : POY_INF);
            opn = pCB[j - 1] + (both ? (clean ? 0 : 2 * go) : POY_INF);
            EB[j] = ext < opn ? ext : opn;
            const int diag = s_cost16[(r.w & 15) * GF_TAB_COLS + (c.w & 15)];
            int a = pCB[j - 1] + diag;
            const int v = pEV[j - 1] + diag + ((r.w & PF_HASGAP) ? c.z : 0);
            const int h = pEH[j - 1] + diag + ((c.w & PF_HASGAP) ? r.z : 0);
            const int d = pEB[j - 1] + diag + (c.z < r.z ? r.z : c.z);
            if (a > v) a = v;
            if (a > h) a = h;
            if (a > d) a = d;
            CB[j] = a;
        }
    }
    const int cur = lasti >= 1 ? (lasti & 1) * stride : 0;
    int res = eh0[cur + lastj];
    res = imin(res, ev0[cur + lastj]);
    res = imin(res, eb0[cur + lastj]);
    res = imin(res, cb0[cur + lastj]);
    return res;
}


// ---- general pairs: 4 states, one row at a time --------------------------------------------------------------
// All 32 lanes work on the SAME row, so everything that depends on the row symbol is warp uniform -- in
// particular "this row and the one above carry no gap bit", which holds for most rows of an interior-node
// sequence and removes the EB state and all flag logic from the cell (fast rows below).  The price is that the
// in-row EH dependency crosses lanes: with E = EH - S_j (S_j = prefix sum of the column extensions) it becomes a
// running minimum,  E_c = min(E_{c-1}, CB_{c-1} + opn_c - S_c),  which is a per-lane chain plus one 5-step warp
// min-scan per row.  Columns are right-aligned as in the other paths; boundary columns between blocks travel
// through the same per-warp scratch, fetched a 32-row window ahead.
template <int C>
__device__ __forceinline__ void cost_pair_rows(const DevCM *cm, const int *s_cost16, const CostJob &J,
                                               const int4 *__restrict__ rowp, const int4 *__restrict__ colp,
                                               const int *__restrict__ g0v, int4 *bnd0, int4 *bnd1, int GO, int lane,
                                               int *__restrict__ cost_out) {
    constexpr int W = 32 * C;
    const int lasti = J.lasti, lastj = J.lastj;
    const int4 *rp = rowp + J.off_i;
    const int4 *cp = colp + J.off_j;
    const int *g0 = g0v + J.off_j;
    if (lastj + 1 <= TINY_L) {
        if (lane == 0) cost_out[J.out] = cost_affine_tiny(cm, s_cost16, rp, cp, g0, lasti, lastj);
        return;
    }
    if (lasti == 0) {  // no rows: minimum over row 0 at the last column (src/algn.c:2105-2109)
        if (lane == 0) cost_out[J.out] = imin(GO + g0[lastj], POY_INF);
        return;
    }
    const int nb = (lastj + W - 1) / W;
    const int pad = nb * W - lastj;
    int s_carry = 0;                         // S of the last column of the previous block
    for (int b = 0; b < nb; ++b) {
        const int jb = b * W + lane * C - pad;  // slot c <-> column jb + c + 1 (<= 0: replica of column 0)
        const int4 *bin = (b & 1) ? bnd0 : bnd1;
        int4 *bout = (b & 1) ? bnd1 : bnd0;
        const bool last_block = (b == nb - 1);
        // per-column constants: K = opn - S_j, Q = S_{j-1} + (column symbol has the gap bit ? GO : 0),
        // byte offset into a table row, flags (bit 4 gap bit, bit 5 previous symbol has it, bit 6 go_j == 0)
        int c_K[C], c_Q[C], c_tab[C], c_fl[C];
        int CBu[C], EVu[C], Eu[C], EBu[C];
        int s_prev_lane;                     // S of the column left of slot 0
        {
            int ext[C], opn[C];
            int run = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jb + c + 1;
                if (j >= 1) {
                    const int4 v = cp[j];
                    ext[c] = v.x; opn[c] = v.y;
                    c_fl[c] = (v.w & (PF_HASGAP | PF_PREVGAP)) | (v.z == 0 ? 64 : 0);
                    c_tab[c] = (v.w & 15) * 4;
                } else {
                    ext[c] = 0; opn[c] = 0; c_fl[c] = 0; c_tab[c] = 16 * 4;
                }
                run += ext[c];
            }
            int incl = run;                  // inclusive warp scan of the per-lane sums
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            s_prev_lane = s_carry + incl - run;
            int sj = s_prev_lane;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jb + c + 1;
                c_Q[c] = sj + ((c_fl[c] & PF_HASGAP) ? GO : 0);
                sj += ext[c];
                c_K[c] = opn[c] - sj;
                if (j >= 1) { CBu[c] = POY_INF; EVu[c] = POY_INF; Eu[c] = GO + g0[j] - sj; }
                else { CBu[c] = 0; EVu[c] = GO; Eu[c] = GO; }          // CB[0][0], EV[0][0], EH[0][0]
                EBu[c] = POY_INF;
            }
            s_carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        const int s_last = s_carry;          // S of the block's last column (lane 31 un-shifts EH with it)
        // left neighbour of lane 0 at row 0: column jb of row 0
        int4 bprev;                          // (CB, E, EV, EB) of the column left of the block, previous row
        if (b == 0) bprev = make_int4(0, GO, GO, POY_INF);
        else bprev = make_int4(POY_INF, GO + g0[b * W - pad] - s_prev_lane, POY_INF, POY_INF);
        int ev_col0 = GO;                    // EV[i][0] running sum (block 0), src/algn.c:2066-2070
        auto load_bnd = [&](int w) { const int r = 32 * w + lane + 1; return b > 0 ? bin[r <= lasti ? r : lasti] : make_int4(0, 0, 0, 0); };
        int4 bw = make_int4(0, 0, 0, 0), bw_next = load_bnd(0);
        int4 rnext = rp[1];
        bool prevH = false;                  // row 0 leaves EB = INF everywhere
        for (int i = 1; i <= lasti; ++i) {
            if (((i - 1) & 31) == 0) { bw = bw_next; bw_next = load_bnd(((i - 1) >> 5) + 1); }
            const int4 r = rnext;
            rnext = rp[i < lasti ? i + 1 : lasti];
            // values of the column to the left: previous row (diagonal) ...
            int xCB0 = __shfl_up_sync(0xffffffffu, CBu[C - 1], 1);
            int xEV0 = __shfl_up_sync(0xffffffffu, EVu[C - 1], 1);
            int xE0 = __shfl_up_sync(0xffffffffu, Eu[C - 1], 1);
            int xEB0 = __shfl_up_sync(0xffffffffu, EBu[C - 1], 1);
            int4 bcur = make_int4(POY_INF, POY_INF, 0, POY_INF);   // ... and this row (CB, E) for the EH chain
            if (b > 0) {
                bcur.x = __shfl_sync(0xffffffffu, bw.x, (i - 1) & 31);
                bcur.y = __shfl_sync(0xffffffffu, bw.y, (i - 1) & 31);
                bcur.z = __shfl_sync(0xffffffffu, bw.z, (i - 1) & 31);
                bcur.w = __shfl_sync(0xffffffffu, bw.w, (i - 1) & 31);
            }
            if (lane == 0) { xCB0 = bprev.x; xE0 = bprev.y; xEV0 = bprev.z; xEB0 = bprev.w; }
            const int vext = r.x, opnV = r.y, go_i = r.z;
            const bool rH = (r.w & PF_HASGAP) != 0, rP = (r.w & PF_PREVGAP) != 0;
            const char *rowbase = (const char *)(s_cost16 + (r.w & 15) * GF_TAB_COLS);
            if (!rH && !prevH) {
                // fast row: no cell of this row or the one above has both gap bits, so EB >= INF throughout,
                // go_i = GO, nothing is charged on CB <- EV, and CB <- EH is charged GO where the column has the bit
                sfor_down<C>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    int xCB, xEV, xE;
                    if constexpr (c == 0) { xCB = xCB0; xEV = xEV0; xE = xE0; }
                    else { xCB = CBu[c - 1]; xEV = EVu[c - 1]; xE = Eu[c - 1]; }
                    const int m = __vimin3_s32(xCB, xEV, xE + c_Q[c]);
                    EVu[c] = __viaddmin_s32(EVu[c], vext, CBu[c] + opnV);
                    CBu[c] = m + *(const int *)(rowbase + c_tab[c]);
                });
            } else {
                const int mask_i = rH ? -1 : 0;
                const int rowlim = rH ? -POY_INF : POY_INF;
                const int rowod = rP ? 2 * GO : 0;
                sfor_down<C>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    int xCB, xEV, xE, xEB;
                    if constexpr (c == 0) { xCB = xCB0; xEV = xEV0; xE = xE0; xEB = xEB0; }
                    else { xCB = CBu[c - 1]; xEV = EVu[c - 1]; xE = Eu[c - 1]; xEB = EBu[c - 1]; }
                    const int fl = c_fl[c];
                    const bool cH = (fl & PF_HASGAP) != 0;
                    const int go_j = (fl & 64) ? 0 : GO;
                    const int od = max(rowod, (fl & PF_PREVGAP) ? 2 * GO : 0);
                    const int lim = max(rowlim, cH ? -POY_INF : POY_INF);
                    const int eb = max(__viaddmin_s32(xCB, od, xEB), lim);
                    const int xEH = xE + c_Q[c] - (cH ? GO : 0);            // un-shift: Q = S_{j-1} + [cH] GO
                    int m = __viaddmin_s32(xEV, go_j & mask_i, xCB);
                    m = __viaddmin_s32(xEH, cH ? go_i : 0, m);
                    m = __viaddmin_s32(xEB, max(go_i, go_j), m);
                    EVu[c] = __viaddmin_s32(EVu[c], vext, CBu[c] + opnV);
                    CBu[c] = m + *(const int *)(rowbase + c_tab[c]);
                    EBu[c] = eb;
                });
            }
            prevH = rH;
            // column 0 has no opening alternative (EV[i][0] = EV[i-1][0] + vext, src/algn.c:2066-2070); its replicas
            // could only find one in row 1, from CB[0][0] = 0 when the first row symbol opens for free
            if (i == 1 && b == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c_tab[c] == 16 * 4) EVu[c] = GO + vext;
            }
            // EH chain: E_c = min(E_{c-1}, CB_{c-1} + K_c); per-lane running minimum, then a warp scan of the lane totals
            int cbleft = __shfl_up_sync(0xffffffffu, CBu[C - 1], 1);
            if (lane == 0) cbleft = bcur.x;
            int P[C];
            P[0] = cbleft + c_K[0];
#pragma unroll
            for (int c = 1; c < C; ++c) P[c] = __viaddmin_s32(CBu[c - 1], c_K[c], P[c - 1]);
            int incl = P[C - 1];
            if (lane == 0) incl = min(incl, bcur.y);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl = min(incl, t);
            }
            int X = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) X = bcur.y;
#pragma unroll
            for (int c = 0; c < C; ++c) Eu[c] = min(P[c], X);
            // F5: EV at the last column of an even row comes from clobbered predecessors
            if (last_block && !(i & 1) && lane == 31) EVu[C - 1] = POY_INF + imin(r.x, r.y);
            if (lane == 31 && !last_block) bout[i] = make_int4(CBu[C - 1], Eu[C - 1], EVu[C - 1], EBu[C - 1]);
            // the column left of the block, as the next row's diagonal neighbour
            if (b == 0) { ev_col0 += vext; bprev = make_int4(POY_INF, POY_INF, ev_col0, POY_INF); }
            else bprev = bcur;
        }
        if (last_block && lane == 31)
            cost_out[J.out] = imin(imin(Eu[C - 1] + s_last, EVu[C - 1]), imin(CBu[C - 1], EBu[C - 1]));
        __syncwarp();
    }
}

// ---- gap-free pairs: 3 states in a shifted domain (see the header of this file and DESIGN.md section 4) ----
// Every DP value of cell (i,j) is carried as  X'(i,j) = X(i,j) - S_j - R_i,  S_j = sum of the column gap
// extensions ge_1..ge_j (pool->d_g0), R_i = sum of the row gap extensions (low bits of pool->d_rowpk).  The
// shift is the same for all states of a cell, so every minimum is unchanged, and the additive constants move
// into the cost table:  tab'[a][b] = cost[a][b] - prepend[b] - cost[a][gap].  What is left per cell is
//     CB' = M'(i-1,j-1) + tab'[a_i][b_j]
//     EH' = min(EH'(i,j-1), CB'(i,j-1) + GO)           <- the only in-row dependency: ONE instruction per column
//     EV' = min(EV'(i-1,j), CB'(i-1,j) + GO)
//     M'  = min3(CB', EH', EV')
// and the answer is min3(..)(lasti,lastj) + S_lastj + R_lasti.  Values that stand for "infinity" only ever lose
// minima against finite values (the domain check keeps every finite value below INF and all costs are >= 0), so
// any stand-in >= INF gives the same result: INF itself is used for CB/EV of row 0, EH of column 0 and for the
// reference's clobbered EV of the last column on even rows (F5).
// Right-aligned columns: column lastj is always slot C-1 of lane 31 of the last block.  The padding on the left
// of block 0 consists of replicas of column 0 (table entry INF): they reproduce EH = INF, EV'[i][0] = GO and
// M'[i][0] = GO exactly and their CB stays >= INF.  Row codes are loaded 32 rows at a time, lane 0 picks its row
// by shuffle and every row then travels down the lanes with the DP values.
template <int C>
__device__ __forceinline__ void cost_pair_gf(const int *s_tab_i, const CostJob &J, const unsigned *__restrict__ rowpk,
                                             const int4 *__restrict__ colp, const int *__restrict__ g0v, int4 *bnd0,
                                             int4 *bnd1, int GO, int lane, int *__restrict__ cost_out) {
    constexpr int W = 32 * C;
    const unsigned tab_base = (unsigned)__cvta_generic_to_shared(s_tab_i);
    const unsigned GF_ROW_BYTES_U = GF_TAB_COLS * 128;
    const int lasti = J.lasti, lastj = J.lastj;
    const unsigned *rp = rowpk + J.off_i;
    const int4 *cp = colp + J.off_j;
    const int *g0 = g0v + J.off_j;
    if (lasti == 0) {  // no rows: minimum over row 0 at the last column (src/algn.c:2105-2109)
        if (lane == 0) cost_out[J.out] = lastj >= 1 ? min(GO + g0[lastj], POY_INF) : 0;
        return;
    }
    const int M0 = min(0, GO);
    const int nb = (lastj + W - 1) / W;
    const int pad = nb * W - lastj;
    for (int b = 0; b < nb; ++b) {
        const int jb = b * W + lane * C - pad;  // slot c <-> column jb + c + 1 (<= 0: replica of column 0)
        const int4 *bin = (b & 1) ? bnd0 : bnd1;
        int4 *bout = (b & 1) ? bnd1 : bnd0;
        const bool last_block = (b == nb - 1);
        int c_off[C], CBu[C], EVu[C], Mu[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jb + c + 1;
            if (j >= 1) {
                c_off[c] = (int)tab_base + ((cp[j].w & 15) << 7) + (lane << 2);
                CBu[c] = POY_INF; EVu[c] = POY_INF;
                Mu[c] = GO;                               // min3(INF, INF, EH[0][j] = GO + S_j) - S_j
            } else {
                c_off[c] = (int)tab_base + (16 << 7) + (lane << 2);
                CBu[c] = 0; EVu[c] = GO;                  // CB[0][0], EV[0][0]
                Mu[c] = M0;
            }
        }
        int dM = jb >= 1 ? GO : M0;                       // M'(0, jb): diagonal predecessor of slot 0 in row 1
        int oCB = POY_INF, oEH = POY_INF, oM = POY_INF;
        // Row codes and the boundary column of the previous block are consumed by lane 0 one row per step; both are
        // fetched 32 rows at a time with one coalesced load per lane, a whole window (32 steps) ahead of their
        // use, and handed to lane 0 by shuffle -- no load sits on the critical path of a step.
        unsigned rk = 0;
        auto load_win = [&](int w) { const int r = 32 * w + lane + 1; return rp[r <= lasti ? r : lasti] >> 28; };
        auto load_bnd = [&](int w) { const int r = 32 * w + lane + 1; return b > 0 ? bin[r <= lasti ? r : lasti] : make_int4(0, 0, 0, 0); };
        unsigned win = 0, win_next = load_win(0);
        int4 bw = make_int4(0, 0, 0, 0), bw_next = load_bnd(0);

        const int nsteps = lasti + 31;
        for (int s = 0; s < nsteps; ++s) {
            if ((s & 31) == 0) {
                win = win_next; bw = bw_next;
                win_next = load_win((s >> 5) + 1);
                bw_next = load_bnd((s >> 5) + 1);
            }
            const int i = s - lane + 1;
            int lCB = __shfl_up_sync(0xffffffffu, oCB, 1);
            int lEH = __shfl_up_sync(0xffffffffu, oEH, 1);
            int lM = __shfl_up_sync(0xffffffffu, oM, 1);
            unsigned rprev = __shfl_up_sync(0xffffffffu, rk, 1);
            const unsigned rfirst = __shfl_sync(0xffffffffu, win, s & 31);
            rk = lane == 0 ? rfirst : rprev;              // row i's table row travels down the lanes
            if (b > 0) {                                  // (uniform) boundary column of row s + 1 for lane 0
                const int bx = __shfl_sync(0xffffffffu, bw.x, s & 31);
                const int by = __shfl_sync(0xffffffffu, bw.y, s & 31);
                const int bz = __shfl_sync(0xffffffffu, bw.z, s & 31);
                if (lane == 0) { lCB = bx; lEH = by; lM = bz; }
            } else if (lane == 0) {                       // column 0: CB = EH = INF, EV'[i][0] = M'[i][0] = GO
                lCB = POY_INF; lEH = POY_INF; lM = GO;
            }
            if (i >= 1) {
                const unsigned irow = rk;
                int mD = dM;
                // Three passes, each of which overwrites its state array in place (no register copies at the loop
                // back edge): EV from the old CB, then CB from the old M, then the EH chain and the new M.
#pragma unroll
                for (int c = 0; c < C; ++c) EVu[c] = __viaddmin_s32(CBu[c], GO, EVu[c]);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    // address = row * (17 columns x 128 B) + column offset as ONE multiply-add: it runs on the
                    // FMA pipe and leaves the ALU pipe to the three min instructions
                    unsigned addr;
                    int diag;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(irow), "r"(GF_ROW_BYTES_U), "r"(c_off[c]));
                    asm("ld.shared.s32 %0, [%1];" : "=r"(diag) : "r"(addr));
                    CBu[c] = (c == 0 ? mD : Mu[c > 0 ? c - 1 : 0]) + diag;
                }
                int cbL = lCB, ehL = lEH;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    ehL = __viaddmin_s32(cbL, GO, ehL);
                    cbL = CBu[c];
                    Mu[c] = __vimin3_s32(cbL, ehL, EVu[c]);
                }
                // F5: EV at the last column of an even row comes from clobbered predecessors (>= INF)
                if (last_block && lane == 31 && !(i & 1)) EVu[C - 1] = POY_INF;
                dM = lM;
                oCB = cbL; oEH = ehL; oM = Mu[C - 1];
                if (lane == 31 && !last_block && i <= lasti) bout[i] = make_int4(oCB, oEH, oM, 0);
            }
        }
        if (last_block && lane == 31)                    // lane 31 finished row lasti in the last step
            cost_out[J.out] = __vimin3_s32(oCB, oEH, EVu[C - 1]) + g0[lastj] + (int)(rp[lasti] & 0x0FFFFFFFu);
        __syncwarp();
    }
}

// ---- one persistent kernel for both kinds of pair ------------------------------------------------------------
// Jobs sit in ONE list (general pairs first: they are the slower ones), warps pull them with an atomic counter and
// dispatch on J.gapfree, so the tail of either kind is filled by the other.
template <int CG, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_cost_affine(const DevCM *__restrict__ cm, const int4 *__restrict__ rowp, const unsigned *__restrict__ rowpk,
              const int4 *__restrict__ colp, const int *__restrict__ g0v, const CostJob *__restrict__ jobs, int njobs,
              int *counter, int4 *bound, size_t bound_stride, int *__restrict__ cost_out) {
    __shared__ int s_tab[16 * GF_TAB_COLS * 32];   // gap-free path: shifted 16 x 17 table, every entry replicated once per bank
    __shared__ int s_cost16[16 * GF_TAB_COLS];      // 4-state path and tiny pairs: plain table, rows padded to 17 entries
    for (int x = threadIdx.x; x < 16 * GF_TAB_COLS * 32; x += blockDim.x) {
        const int e = x >> 5, a = e / GF_TAB_COLS, b = e % GF_TAB_COLS;
        s_tab[x] = b < 16 ? cm->cost16[a * 16 + b] - cm->prepend[b] - cm->gapext[a] : POY_INF;
    }
    for (int x = threadIdx.x; x < 16 * GF_TAB_COLS; x += blockDim.x) {
        const int a = x / GF_TAB_COLS, b = x % GF_TAB_COLS;
        s_cost16[x] = b < 16 ? cm->cost16[a * 16 + b] : POY_INF;
    }
    __syncthreads();
    const int GO = cm->gap_open;
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
        if (J.gapfree) cost_pair_gf<CG>(s_tab, J, rowpk, colp, g0v, bnd0, bnd1, GO, lane, cost_out);
        else cost_pair_rows<(CG >= 32 ? 16 : 8)>(cm, s_cost16, J, rowp, colp, g0v, bnd0, bnd1, GO, lane, cost_out);
    }
}

cudaError_t launch_cost_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const CostJob *d_jobs, int njobs,
                               int *d_counter, int4 *d_bound, size_t bound_stride, int blocks, int *d_cost, int wide) {
    // columns per lane of the gap-free path: 32 (column blocks of 1024, half the boundary traffic, fewer per-step
    // overheads) pays off once the sequences span more than one such block; 16 wastes fewer lanes on short pairs.
    // POY_COST_C=16|32 overrides (tuning knob).
    int cg = wide ? 32 : 16;
    { const char *e = getenv("POY_COST_C"); if (e && (atoi(e) == 16 || atoi(e) == 24 || atoi(e) == 32)) cg = atoi(e); }
    if (cg == 24)
        k_cost_affine<24, 3><<<blocks, 128, 0, ctx->stream>>>(cm->d, pool->d_rowp, pool->d_rowpk, pool->d_colp, pool->d_g0, d_jobs, njobs,
                                                               d_counter, d_bound, bound_stride, d_cost);
    else if (cg == 32)
        k_cost_affine<32, 2><<<blocks, 128, 0, ctx->stream>>>(cm->d, pool->d_rowp, pool->d_rowpk, pool->d_colp, pool->d_g0, d_jobs, njobs,
                                                               d_counter, d_bound, bound_stride, d_cost);
    else
        k_cost_affine<16, 3><<<blocks, 128, 0, ctx->stream>>>(cm->d, pool->d_rowp, pool->d_rowpk, pool->d_colp, pool->d_g0, d_jobs, njobs,
                                                               d_counter, d_bound, bound_stride, d_cost);
    ctx->launches++;
    return cudaGetLastError();
}
