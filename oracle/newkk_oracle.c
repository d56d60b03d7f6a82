/* TEST INFRASTRUCTURE ONLY (same rules as do_oracle.c: only tests/, smoke() and bench.py's CPU legs may
 * load the library this file is built into).
 *
 * CPU restatement of the diagonal-storage Ukkonen alignment of src/newkkonen.c (Sequence.NewkkAlign,
 * src/sequence.ml:1831-2062), AFFINE model (newkkonen_CAML_algn_affine + newkkonen_CAML_backtrace_affine).
 *
 * Parity status: PINNED against the unmodified reference compiled into oracle/_ref/libpoyref*.so
 * (ref_newkk_driver.c; tests/test_newkk.py compares cost, both aligned rows and the final k on seeded
 * random / decorated / edge pairs in the build container, and commits reference-generated goldens under
 * tests/golden/newkk_golden.npz for the GPU box).
 *
 * What is restated, and how it differs from a transcription:
 *  - the reference keeps ONE matrix in diagonal-major storage across threshold doublings and re-evaluates
 *    only the cells whose neighbours changed (two queues per row, update_a_row :964-1073).  A cell's value
 *    is a pure function of its three neighbours and of its border status, and every cell whose inputs or
 *    border status changed is re-evaluated (first-time cells report "changed", old border cells have a
 *    first-time neighbour), so the matrix after a doubling equals a FRESH fill of the new band.  This
 *    file (and the CUDA kernel) therefore refill the band from scratch per doubling, in band-only
 *    row-major storage;
 *  - row 0 up to the base band is initialised by newkk_algn (:1390-1431) with its own direction word
 *    (DO_INSERT|END_INSERT|END_DELETE|END_DIAG); when the first band has k = 0 the reference recomputes
 *    those cells through update_internal_cell, which yields the same costs and gap counts and a direction
 *    word without END_INSERT for j >= 2 -- not observable by the traceback (an insertion run on row 0
 *    continues to (0,0) either way), so the initial word is kept;
 *  - the NON-affine entry point (newkkonen_CAML_algn) is not restated: update_internal_cell never sets
 *    costDiag on that path (:796, :855-856), so every interior cell gets cost 0 / DO_DIAG and the
 *    traceback walks out of the band and raises Failure (tests/test_newkk.py documents it against the
 *    compiled reference).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef unsigned char u8;
typedef unsigned short u16;

#define NK_INF 0x3fffffff /* INT_MAX/2, src/newkkonen.c:65-71 */
#define NK_MUCH_LONGER 100
#define NK_ALIGN 1
#define NK_DO_DELETE 16
#define NK_DO_INSERT 32
#define NK_END_INSERT 64
#define NK_END_DELETE 128
#define NK_END_DIAG 256
#define NK_DO_DIAG 512
#define NK_INS_EQ_DEL 1024

typedef struct { int cost, P, Q, ED, CD; u16 dir, g1, g2; } nkcell;

/* same layout as do_cm in do_oracle.c */
typedef struct {
    int cost[1024]; int worst[1024]; u8 median[1024]; int prepend[32]; int tail[32];
    int gap_open; int cost_model_type; int min_non0;
} nk_cm;

static int nk_add(int a, int b) { return (a >= NK_INF || b >= NK_INF) ? NK_INF : a + b; }   /* my_add */
static int nk_min(int a, int b) { return a < b ? a : b; }
static int nk_max(int a, int b) { return a > b ? a : b; }
/* cm_calc_cost on the calloc'd 32x32 table: row / column 0 are never set (src/cm.c:627) */
static int nk_cost(const nk_cm *c, int a, int b) { return (a == 0 || b == 0) ? 0 : c->cost[(a << 5) + b]; }
/* get_go / get_extgo, src/newkkonen.c:179-198 */
static int nk_go(int base, int prev, int idx, int go) {
    if (idx == 1 && (base & 16)) return 0;
    if (idx > 1 && !(prev & 16) && (base & 16)) return 0;
    return go;
}
static int nk_extgo(int base, int prev, int go) { return ((prev & 16) && !(base & 16)) ? go : 0; }

/* update_internal_cell + assign_best_cost_and_direction, src/newkkonen.c:600-870 (affine branch).
 * L / U / M = cells (i,j-1) / (i-1,j) / (i-1,j-1), NULL where the reference does not read them. */
static void nk_update(const nk_cm *c, const u8 *s1, const u8 *s2, int i, int j, int realgo, const nkcell *L,
                      const nkcell *U, const nkcell *M, nkcell *out) {
    const int b1 = s1[i], b2 = s2[j], p1 = i > 0 ? s1[i - 1] : 0, p2 = j > 0 ? s2[j - 1] : 0;
    const int go1 = nk_go(b1, p1, i, realgo), go2 = nk_go(b2, p2, j, realgo);
    const int xg1 = nk_extgo(b1, p1, realgo), xg2 = nk_extgo(b2, p2, realgo);
    int thisP = NK_INF, thisQ = NK_INF, thisED = NK_INF, thisCD = NK_INF;
    int costL, extL, openL, costR, extR, openR, costM, costD, extD, openD;
    int g1L = 0, g2L = 0, g1R = 0, g2R = 0, g1M = 0, g2M = 0;
    if (!L) costL = extL = openL = NK_INF;
    else {
        const int add = nk_cost(c, b2, 16);
        extL = nk_add(L->Q, add + xg2); openL = nk_add(L->CD, add + go2);
        costL = nk_min(openL, extL); thisQ = costL; g1L = L->g1; g2L = L->g2;
    }
    if (!U) costR = extR = openR = NK_INF;
    else {
        const int add = nk_cost(c, b1, 16);
        extR = nk_add(U->P, add + xg1); openR = nk_add(U->CD, add + go1);
        costR = nk_min(openR, extR); thisP = costR; g1R = U->g1; g2R = U->g2;
    }
    if (!M) costM = costD = extD = openD = NK_INF;
    else {
        const int add = nk_cost(c, b1 & 15, b2 & 15);
        const int b1n = b1 & 15, b2n = b2 & 15;
        const int fromR = M->P + nk_cost(c, b1n, b2) + (b1n == b1 ? 0 : realgo);
        const int fromL = M->Q + nk_cost(c, b1, b2n) + (b2n == b2 ? 0 : realgo);
        thisCD = nk_add(M->CD, add);
        thisCD = nk_min(thisCD, fromR); thisCD = nk_min(thisCD, fromL);
        thisCD = nk_min(thisCD, nk_add(M->ED, go1 + go2 + add));
        openD = nk_add(M->CD, go1 + go2);
        extD = nk_add(M->ED, ((b1 & 16) && (b2 & 16)) ? 0 : NK_INF);
        thisED = nk_min(extD, openD);
        costM = thisCD; costD = thisED; g1M = M->g1; g2M = M->g2;
    }
    {
        int best = costL, dir = NK_DO_INSERT, r1 = g1L + 1, r2 = g2L;
        if (costR <= best) {
            if (costR < best) { best = costR; dir = NK_DO_DELETE; r1 = g1R; r2 = g2R + 1; }
            else { dir |= NK_DO_DELETE; r1 = nk_max(r1, g1R); r2 = nk_max(r2, g2R + 1); }
        }
        if (costM <= best) {
            if (costM < best) { best = thisCD; dir = NK_ALIGN; r1 = g1M; r2 = g2M; }
            else { dir |= NK_ALIGN; r1 = nk_max(r1, g1M); r2 = nk_max(r2, g2M); }
        }
        if (costD <= best) {
            if (costD < best) { best = costD; dir = NK_DO_DIAG; r1 = g1M; r2 = g2M; }
            else { dir |= NK_DO_DIAG; r1 = nk_max(r1, g1M); r2 = nk_max(r2, g2M); }
        }
        if (extR >= openR) dir |= NK_END_DELETE;
        if (extL >= openL) dir |= NK_END_INSERT;
        if (extD >= openD) dir |= NK_END_DIAG;
        if (extR == extL && extR == best) dir |= NK_INS_EQ_DEL;
        out->cost = best; out->dir = (u16)dir; out->g1 = (u16)r1; out->g2 = (u16)r2;
        out->P = thisP; out->Q = thisQ; out->ED = thisED; out->CD = thisCD;
    }
}

typedef struct { int iterations, final_T, final_k; long long cells; } nk_stats;

/* newkk_algn (affine) + backtrace_affine.  s1 must be the shorter sequence (failwith otherwise -> INT_MIN).
 * r1 / r2 receive the aligned rows in reading order (capacity len1 + len2), lens[2] their lengths (NULL r1: cost only). */
int do_newkk_align_affine(const nk_cm *c, const u8 *s1, int len1, const u8 *s2, int len2, int swaped, u8 *r1, u8 *r2,
                          int *lens, nk_stats *st) {
    const int realgo = c->gap_open > 0 ? c->gap_open : 0;
    const int delta = len2 - len1, bb = delta + 1;
    int T, k = 0, i, j, cost = 0, iters = 0;
    long long cells = 0;
    nkcell *m = NULL;
    int W = 0;
    if (len1 > len2) return (-2147483647 - 1);
    if (st) memset(st, 0, sizeof *st);
    if ((long long)len1 * NK_MUCH_LONGER < len2) {   /* trivial_algn / trivial_backtrace, :1351-1356, 1498-1521 */
        int n = 0;
        for (i = 0; i < len1; i++) cost += nk_cost(c, s1[i] & 15, 16);
        for (i = 0; i < len2; i++) cost += nk_cost(c, s2[i] & 15, 16);
        if (r1) {   /* prepended in reading order of the inputs: both rows come out reversed */
            for (i = len2 - 1; i >= 0; i--) { r1[n] = 16; r2[n] = s2[i]; n++; }
            for (i = len1 - 1; i >= 0; i--) { r1[n] = s1[i]; r2[n] = 16; n++; }
            lens[0] = lens[1] = n;
        }
        return cost;
    }
    T = bb * c->min_non0;
    for (;;) {
        const int p = (T - delta) / 2, newp = (2 * T - delta) / 2;
        int gn;
        k = p >= len1 ? len1 - 1 : p;
        W = delta + 2 * k + 1;
        free(m);
        m = (nkcell *)malloc(sizeof(nkcell) * (size_t)len1 * (size_t)W);
        ++iters;
#define CELL(i, j) (m + (size_t)(i) * W + ((j) - (i) + k))
        for (i = 0; i < len1; i++) {
            const int startj = i - k > 0 ? i - k : 0;
            const int endj = i + delta + k < len2 - 1 ? i + delta + k : len2 - 1;
            for (j = startj; j <= endj; j++) {
                nkcell *o = CELL(i, j);
                ++cells;
                if (i == 0 && j == 0) {          /* :1392 */
                    o->cost = 0; o->dir = 0; o->g1 = o->g2 = 0; o->P = realgo; o->Q = realgo; o->ED = NK_INF; o->CD = 0;
                } else if (i == 0 && j < bb) {   /* first row of the base band, :1394-1431 */
                    const nkcell *l = CELL(0, j - 1);
                    const int b = s2[j], pb = s2[j - 1], f = pb & 16, f2 = b & 16;
                    const int add = nk_cost(c, b, 16);
                    const int goc = (j == 1) ? (f2 ? 0 : realgo) : ((!f && f2) ? 0 : realgo);
                    const int ext = (f && !f2) ? l->Q + add + realgo : l->Q + add;
                    const int opn = l->cost + add + goc;
                    const int cl = nk_min(opn, ext);
                    o->cost = cl; o->dir = NK_DO_INSERT | NK_END_INSERT | NK_END_DELETE | NK_END_DIAG;
                    o->g1 = (u16)(l->g1 + 1); o->g2 = l->g2; o->P = NK_INF; o->Q = cl; o->ED = NK_INF; o->CD = NK_INF;
                } else {
                    const int lb = (i - j == k), rb = (j - i == delta + k);
                    nk_update(c, s1, s2, i, j, realgo, (lb || j == 0) ? NULL : CELL(i, j - 1), (rb || i == 0) ? NULL : CELL(i - 1, j),
                              (i == 0 || j == 0) ? NULL : CELL(i - 1, j - 1), o);
                }
            }
        }
        cost = CELL(len1 - 1, len2 - 1)->cost;
        gn = nk_max(CELL(len1 - 1, len2 - 1)->g1, CELL(len1 - 1, len2 - 1)->g2);
        if (p > gn || newp - len2 + 1 >= 0) break;      /* increaseT, :1155-1171 */
        T *= 2;
    }
    if (st) { st->iterations = iters; st->final_T = T; st->final_k = k; st->cells = cells; }
    if (r1) {   /* backtrace_affine, :1666-1763 */
        const int cap = len1 + len2;
        int n1 = cap, n2 = cap, mode = 0;   /* 0 todo, 1 delete, 2 insert, 3 diagonal, 4 align */
        u8 *t1 = (u8 *)malloc((size_t)cap + 2), *t2 = (u8 *)malloc((size_t)cap + 2);
        i = len1 - 1; j = len2 - 1;
        while (i >= 0 && j >= 0) {
            const int dir = CELL(i, j)->dir;
            if (dir == 0) { t1[--n1] = 16; t2[--n2] = 16; i--; j--; continue; }
            if (mode == 0) {
                const int hi = dir & NK_DO_INSERT, hd = dir & NK_DO_DELETE, ha = dir & NK_ALIGN, hg = dir & NK_DO_DIAG;
                if (!swaped) mode = hd ? 1 : hi ? 2 : hg ? 3 : ha ? 4 : -1;
                else mode = hi ? 2 : hd ? 1 : hg ? 3 : ha ? 4 : -1;
                if (mode < 0) { free(t1); free(t2); free(m); return (-2147483647 - 1); }
            } else if (mode == 1) {
                t1[--n1] = s1[i]; t2[--n2] = 16; i--;
                if ((dir & NK_END_DELETE) || (dir & NK_INS_EQ_DEL)) mode = 0;
            } else if (mode == 2) {
                t1[--n1] = 16; t2[--n2] = s2[j]; j--;
                if ((dir & NK_END_INSERT) || (dir & NK_INS_EQ_DEL)) mode = 0;
            } else if (mode == 3) {
                if (dir & NK_END_DIAG) mode = 0;
                t1[--n1] = s1[i]; t2[--n2] = s2[j]; i--; j--;
            } else {
                t1[--n1] = s1[i] & 15; t2[--n2] = s2[j] & 15; i--; j--; mode = 0;
            }
        }
        memcpy(r1, t1 + n1, (size_t)(cap - n1)); memcpy(r2, t2 + n2, (size_t)(cap - n2));
        lens[0] = cap - n1; lens[1] = cap - n2;
        free(t1); free(t2);
    }
#undef CELL
    free(m);
    return cost;
}
