/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the POY5 dynamic-homology
 * pairwise alignment hot path.  Nothing in poy5_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED against the unmodified reference C compiled from
 * /root/reference/src into oracle/_ref/libpoyref.so (tests/test_oracle_vs_ref.py,
 * run in the build container) and against the golden vectors generated from that
 * library (tests/golden/, generator tests/golden/make_golden.py).  The reference
 * ships no tests or fixtures of its own for this path.
 *
 * Conventions: sequences are uint8 DNA bitsets
 * A=1 C=2 G=4 T=8 gap=16, element 0 is always the gap code and takes part in
 * the DP as row/column 0.  `len` counts that element.  All arithmetic is plain
 * wrapping int32; DO_INF (the reference's HIGH_NUM, src/algn.c:37) is added to,
 * never saturated.
 *
 * Every function cites the reference lines it restates.  The restatement keeps
 * the reference's *memory semantics* where they are observable (the row-buffer
 * aliasing of the cost-only entry point, the never-cleared unsigned-short
 * gap-count rows of the banded entry point) by laying its scratch out the same
 * way, but it is written from the recurrences, not transcribed.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define DO_INF 1000000
#define GAPBIT 16
#define NOGAP 15

typedef unsigned char u8;
typedef unsigned short u16;

/* Flat cost model for the 5-letter bitset alphabet (struct cm, src/cm.h:33-76,
 * restricted to combinations=1, level=0, lcm=5, gap=16). Tables indexed (a<<5)+b. */
typedef struct {
    int cost[1024];
    int worst[1024];
    u8 median[1024];
    int prepend[32];
    int tail[32];
    int gap_open;
    int cost_model_type; /* 0 linear, 1 affine, 2 no alignment */
    int min_non0;        /* cm_get_min_non0_cost, src/cm.c:1063-1089 */
} do_cm;

int do_cm_sizeof(void) { return (int)sizeof(do_cm); }

void do_cm_init(do_cm *c, const int *cost, const int *worst, const u8 *median, const int *prepend,
                const int *tail, int gap_open, int cost_model_type) {
    int i, m = 0x3fffffff; /* INT_MAX/2 */
    memcpy(c->cost, cost, sizeof c->cost);
    memcpy(c->worst, worst, sizeof c->worst);
    memcpy(c->median, median, sizeof c->median);
    memcpy(c->prepend, prepend, sizeof c->prepend);
    memcpy(c->tail, tail, sizeof c->tail);
    c->gap_open = gap_open;
    c->cost_model_type = cost_model_type;
    for (i = 0; i < 1024; i++)
        if (c->cost[i] > 0 && c->cost[i] < m) m = c->cost[i];
    c->min_non0 = m;
}

/* ---- per-position gap parameters ---------------------------------------- */
/* HAS_GAP_OPENING, src/algn.c:1240-1253 (gap_startNO = 0 branch) */
static int gap_opening_at(int idx, int prev, int cur, int go) {
    if (idx == 1 && (cur & GAPBIT)) return 0;
    if (idx > 1 && !(prev & GAPBIT) && (cur & GAPBIT)) return 0;
    return go;
}

/* Column parameters shared by both affine entry points
 * (src/algn.c:2026-2035 and :2225-2236): go_j, and the horizontal extension
 * cost *as computed inside the loop* (i.e. before hext[1] is overwritten). */
static void column_params(const do_cm *c, const u8 *sj, int lastj, int *gop, int *hext) {
    int j;
    hext[0] = 0;
    for (j = 1; j <= lastj; j++) {
        int ge = c->prepend[sj[j]]; /* gap_row = prec row 0 = prepend, A4 */
        gop[j] = gap_opening_at(j, sj[j - 1], sj[j], c->gap_open);
        hext[j] = ((sj[j - 1] & GAPBIT) && !(sj[j] & GAPBIT)) ? gop[j] + ge : ge;
    }
}

typedef struct { int ge, go, vext, nogap, hasgap, prevgap; const int *row; } rowp;

/* Row parameters (src/algn.c:2058-2065, 2273-2283) */
static rowp row_params(const do_cm *c, const u8 *si, int i) {
    rowp r;
    int ic = si[i], ip = si[i - 1];
    r.ge = c->cost[(ic << 5) + GAPBIT];               /* HAS_GAP_EXTENSION :1220 */
    r.go = gap_opening_at(i, ip, ic, c->gap_open);
    r.nogap = ic & NOGAP;
    r.hasgap = (ic & GAPBIT) != 0;
    r.prevgap = (ip & GAPBIT) != 0;
    r.vext = (i > 1 && (ip & GAPBIT) && !(ic & GAPBIT)) ? r.go + r.ge : r.ge;
    r.row = c->cost + (r.nogap << 5);
    return r;
}

/* ======================================================================== */
/* 1. cost-only affine: algn_CAML_cost_affine_3 -> algn_fill_plane_3_aff_nobt */
/*    (src/algn.c:2457-2515, 1822-1863, 1987-2110)                           */
/* ======================================================================== */
int do_cost_affine(const do_cm *c, const u8 *s1, int len1, const u8 *s2, int len2) {
    const u8 *si, *sj;
    int leni, lenj, L, lasti, lastj, stride, i, j, go = c->gap_open, res;
    int *M, *gop, *hext;
    int *cb[2], *eb[2], *ev[2], *eh[2];
    if (len1 <= len2) { si = s1; leni = len1; sj = s2; lenj = len2; }
    else { si = s2; leni = len2; sj = s1; lenj = len1; }
    L = lenj;              /* `largest`, :2479-2481 */
    lasti = leni - 1;
    lastj = lenj - 1;
    stride = lastj + 2;    /* `offset`, :2005 -- one more than the 2-row spacing L */
    /* Same flat layout as the reference scratch (:2489-2494); the second row of
     * each pair starts at +stride, so X_row1[lastj] IS Y_row0[0] for the
     * matrix Y laid out after X. */
    M = (int *)calloc((size_t)13 * L + 16, sizeof(int));
    cb[0] = M;          cb[1] = cb[0] + stride;
    eb[0] = M + 2 * L;  eb[1] = eb[0] + stride;
    ev[0] = M + 4 * L;  ev[1] = ev[0] + stride;
    eh[0] = M + 6 * L;  eh[1] = eh[0] + stride;
    gop = M + 10 * L;
    hext = M + 11 * L;
    /* row 0 (:1831-1852) */
    cb[0][0] = 0; eb[0][0] = DO_INF; eh[0][0] = go; ev[0][0] = go;
    for (j = 1; j <= lastj; j++) {
        eh[0][j] = eh[0][j - 1] + c->prepend[sj[j]];
        cb[0][j] = DO_INF; eb[0][j] = DO_INF; ev[0][j] = DO_INF;
    }
    column_params(c, sj, lastj, gop, hext);
    if (lastj >= 1) hext[1] = c->prepend[sj[1]]; /* :2035, A2 */
    for (i = 1; i <= lasti; i++) {
        int cur = i & 1, prv = cur ^ 1; /* row 0 lives in buffer 0 */
        int *CB = cb[cur], *EB = eb[cur], *EV = ev[cur], *EH = eh[cur];
        const int *pCB = cb[prv], *pEB = eb[prv], *pEV = ev[prv], *pEH = eh[prv];
        rowp r = row_params(c, si, i);
        int r0;
        /* column 0 (:2057-2072), in the reference's store order */
        EH[0] = DO_INF;
        r0 = pEV[0] + r.vext;
        EH[0] = DO_INF; CB[0] = r0; EB[0] = DO_INF; EV[0] = r0; CB[0] = DO_INF;
        for (j = 1; j <= lastj; j++) {
            int jc = sj[j], jp = sj[j - 1];
            int ge_j = c->prepend[jc], go_j = gop[j];
            int ext, opn, both, clean, dg, od, diag, a, v, h, d, xgo;
            /* extend horizontal (:1260-1273): ties take the opening */
            ext = EH[j - 1] + hext[j];
            opn = CB[j - 1] + go_j + ge_j;
            EH[j] = ext < opn ? ext : opn;
            /* extend vertical (:1306-1318) */
            ext = pEV[j] + r.vext;
            opn = pCB[j] + r.go + r.ge;
            EV[j] = ext < opn ? ext : opn;
            /* extend block diagonal (:1351-1369) */
            both = r.hasgap && (jc & GAPBIT);
            clean = !r.prevgap && !(jp & GAPBIT);
            dg = both ? 0 : DO_INF;
            od = both ? (clean ? 0 : 2 * go) : DO_INF;
            ext = pEB[j - 1] + dg;
            opn = pCB[j - 1] + od;
            EB[j] = ext < opn ? ext : opn;
            /* close block diagonal (:1401-1433) */
            diag = r.row[jc & NOGAP];
            xgo = go_j < r.go ? r.go : go_j;
            a = pCB[j - 1] + diag;
            v = r.hasgap ? pEV[j - 1] + diag + go_j : pEV[j - 1] + diag;
            h = (jc & GAPBIT) ? pEH[j - 1] + diag + r.go : pEH[j - 1] + diag;
            d = pEB[j - 1] + diag + xgo;
            if (a > v) a = v;
            if (a > h) a = h;
            if (a > d) a = d;
            CB[j] = a;
        }
    }
    {
        int cur = lasti >= 1 ? (lasti & 1) : 0;
        res = eh[cur][lastj];
        if (res > ev[cur][lastj]) res = ev[cur][lastj];
        if (res > eb[cur][lastj]) res = eb[cur][lastj];
        if (res > cb[cur][lastj]) res = cb[cur][lastj];
    }
    free(M);
    return res;
}

/* ======================================================================== */
/* 2. banded affine with traceback: algn_CAML_align_affine_3                 */
/*    (src/algn.c:2359-2447, 1866-1899, 2113-2354, 1715-1819, 126-176)       */
/* ======================================================================== */
/* 15-bit direction mask, src/algn.c:1185-1199 */
enum {
    A2A = 1, A2V = 2, A2H = 4, A2D = 8, BEG_B = 16, END_B = 32, BEG_V = 64, END_V = 128,
    BEG_H = 256, END_H = 512, DO_A = 1024, DO_V = 2048, DO_H = 4096, H_EQ_V = 8192, DO_D = 16384
};

static int imax(int a, int b) { return a > b ? a : b; }

/* algn_fill_gapnum (src/algn.c:126-176): max-plus over unsigned short rows;
 * evaluated in int, truncated on store (A7). p* = previous row, g* = current. */
static void gapnum(int j, int al, int ins, int del, const u16 *p1, u16 *g1, const u16 *p2, u16 *g2) {
    if (al) {
        if (del && ins) {
            g1[j] = (u16)imax(imax(p1[j - 1], g1[j - 1] + 1), imax(g1[j - 1] + 1, p1[j]));
            g2[j] = (u16)imax(imax(p2[j - 1], g2[j - 1]), imax(g2[j - 1], p2[j] + 1));
        } else if (del) {
            g1[j] = (u16)imax(p1[j], p1[j - 1]);
            g2[j] = (u16)imax(p2[j - 1], p2[j] + 1);
        } else if (ins) {
            g1[j] = (u16)imax(g1[j - 1] + 1, p1[j - 1]);
            g2[j] = (u16)imax(p2[j - 1], g2[j - 1]);
        } else {
            g1[j] = p1[j - 1];
            g2[j] = p2[j - 1];
        }
    } else if (ins) {
        if (del) {
            g1[j] = (u16)imax(g1[j - 1] + 1, p1[j]);
            g2[j] = (u16)imax(p2[j] + 1, g2[j - 1]);
        } else {
            g1[j] = (u16)(g1[j - 1] + 1);
            g2[j] = g2[j - 1];
        }
    } else {
        g1[j] = p1[j];
        g2[j] = (u16)(p2[j] + 1);
    }
}

typedef struct {
    int iterations;        /* number of band fills (T, 2T, 4T, ...) */
    int final_T, final_k;  /* threshold / half band of the accepted fill */
    long long cells;       /* band cells computed over all fills */
} do_align_stats;

/* One band fill with half-width parameter p: algn_newkk_test_aff (:2186-2306)
 * + algn_newkk_fill_a_row_aff (:2113-2180) + ASSIGN_MINIMUM (:1936-1981).
 * Scratch rows W (2 rows x 4 states), gm (4 u16 rows) and the direction matrix
 * persist across fills exactly like the reference's global scratch.         */
static int band_fill(const do_cm *c, const u8 *si, int lasti, const u8 *sj, int lastj, int p,
                     int *W, int *gop, int *hext, u16 **gmbuf, u16 *dir, int *res_cost, long long *cells) {
    int go = c->gap_open, stride = lastj + 2, i, j;
    int k = p >= lasti ? lasti - 1 : p; /* :2195-2196 */
    int delta = lastj - lasti;
    int *cb[2], *eb[2], *ev[2], *eh[2];
    u16 *g1 = gmbuf[0], *n1 = gmbuf[1], *g2 = gmbuf[2], *n2 = gmbuf[3]; /* gap_num1..4, :2415-2418 */
    cb[0] = W;              cb[1] = cb[0] + stride;
    eb[0] = W + 2 * stride; eb[1] = eb[0] + stride;
    ev[0] = W + 4 * stride; ev[1] = ev[0] + stride;
    eh[0] = W + 6 * stride; eh[1] = eh[0] + stride;
    /* row 0 as re-initialised by the band fill (:2222-2248, A1): finite CB */
    ev[0][0] = DO_INF; cb[0][0] = 0; g1[0] = 0; g2[0] = 0;
    column_params(c, sj, lastj, gop, hext);
    for (j = 1; j <= lastj; j++) {
        g1[j] = (u16)j; g2[j] = 0; n1[j] = (u16)-1; n2[j] = (u16)-1;
        dir[j] = DO_H;
        ev[0][j] = DO_INF;
        cb[0][j] = cb[0][j - 1] + hext[j];
        eh[0][j] = cb[0][j];
    }
    if (lastj >= 1) hext[1] = c->prepend[sj[1]]; /* :2248, A2 */
    for (i = 1; i <= lasti; i++) {
        int cur = i & 1, prv = cur ^ 1;
        int *CB = cb[cur], *EB = eb[cur], *EV = ev[cur], *EH = eh[cur];
        const int *pCB = cb[prv], *pEB = eb[prv], *pEV = ev[prv], *pEH = eh[prv];
        u16 *drow = dir + (size_t)i * stride;
        rowp r = row_params(c, si, i);
        int startj = i - k > 0 ? i - k : 0;
        int endj = (i + delta + k <= lastj - 1) ? i + delta + k : lastj; /* :2256-2257, A8 */
        for (j = startj; j <= endj; j++) {
            /* the reference primes jc with sj[startj] before the loop, so at the
             * left-border cell "previous column symbol" is sj[j] itself (:2126-2133) */
            int jc = sj[j], jp = j == startj ? sj[j] : sj[j - 1];
            int right = (j - i - (delta + 1) + 1 == k), left = (j == startj);
            int ge_j = c->prepend[jc], go_j = gop[j];
            int mask = 0, ext, opn, fin, m;
            if (!left) {
                ext = EH[j - 1] + hext[j];
                opn = CB[j - 1] + go_j + ge_j;
                if (ext < opn) { mask |= BEG_H; EH[j] = ext; } else { mask |= END_H; EH[j] = opn; }
            } else EH[j] = DO_INF;
            if (!right) {
                ext = pEV[j] + r.vext;
                opn = pCB[j] + r.go + r.ge;
                if (ext < opn) { mask |= BEG_V; EV[j] = ext; } else { mask |= END_V; EV[j] = opn; }
            } else EV[j] = DO_INF;
            if (j > 0) {
                int both = r.hasgap && (jc & GAPBIT), clean = !r.prevgap && !(jp & GAPBIT);
                int dg = both ? 0 : DO_INF, od = both ? (clean ? 0 : 2 * go) : DO_INF;
                int diag, xgo, a, v, h, d, cm;
                ext = pEB[j - 1] + dg;
                opn = pCB[j - 1] + od;
                if (ext < opn) { mask |= BEG_B; EB[j] = ext; } else { mask |= END_B; EB[j] = opn; }
                /* close block diagonal with tie mask, order A,V,H,D (:1436-1492, A6) */
                diag = r.row[jc & NOGAP];
                xgo = go_j < r.go ? r.go : go_j;
                a = pCB[j - 1] + diag;
                v = r.hasgap ? pEV[j - 1] + diag + go_j : pEV[j - 1] + diag;
                h = (jc & GAPBIT) ? pEH[j - 1] + diag + r.go : pEH[j - 1] + diag;
                d = pEB[j - 1] + diag + xgo;
                cm = A2A;
                if (a >= v) { if (a > v) { a = v; cm = A2V; } else cm |= A2V; }
                if (a >= h) { if (a > h) { a = h; cm = A2H; } else cm |= A2H; }
                if (a >= d) { if (a > d) { a = d; cm = A2D; } else cm |= A2D; }
                CB[j] = a;
                mask |= cm;
            } else { CB[j] = DO_INF; EB[j] = DO_INF; }
            /* final minimum with tie mask, order H,V,D,A (:1936-1977) */
            m = DO_H; fin = EH[j];
            if (fin >= EV[j]) { if (fin > EV[j]) { fin = EV[j]; m = DO_V; } else m |= DO_V; }
            if (fin >= EB[j]) { if (fin > EB[j]) { fin = EB[j]; m = DO_D; } else m |= DO_D; }
            if (fin >= CB[j]) { if (fin > CB[j]) { fin = CB[j]; m = DO_A; } else m |= DO_A; }
            if (fin == EH[j] && EH[j] == EV[j]) m |= H_EQ_V;
            mask |= m;
            gapnum(j, (mask & (DO_A | DO_D)) != 0, (mask & DO_H) != 0, (mask & DO_V) != 0, g1, n1, g2, n2);
            drow[j] = (u16)mask;
            if (j == lastj) *res_cost = fin; /* final_cost_matrix is one row (:2168, :2302) */
            (*cells)++;
        }
        if (i <= lasti - 1) { u16 *t = g1; g1 = n1; n1 = t; t = g2; g2 = n2; n2 = t; } /* :2293-2300 */
    }
    return imax(n1[lastj], n2[lastj]); /* :2305 */
}

/* traceback state machine backtrace_aff (src/algn.c:1715-1819, 1531-1619).
 * Outputs are produced by prepending; here they are written right-to-left into
 * caller buffers of capacity cap and then shifted to the front.             */
typedef struct { u8 *buf; int cap, pos; } rseq;
static void rprep(rseq *s, int v) { s->buf[--s->pos] = (u8)v; }
static int rfinish(rseq *s) {
    int n = s->cap - s->pos;
    memmove(s->buf, s->buf + s->pos, (size_t)n);
    return n;
}

static void indel_emit(rseq *med, rseq *medwg, int sym) {
    if (!(sym & GAPBIT)) { rprep(med, sym | GAPBIT); rprep(medwg, sym | GAPBIT); }
    else rprep(medwg, GAPBIT);
}

static int pick_mode(int dg, int al, int v, int h, int swaped) {
    /* choose_dir (:1594-1619): 1 vertical, 2 horizontal, 3 diagonal, 4 align */
    (void)al;
    if (!swaped) {
        if (v) return 1;
        if (h) return 2;
    } else {
        if (h) return 2;
        if (v) return 1;
    }
    return dg ? 3 : 4;
}

static void traceback(const do_cm *c, const u16 *dir, const u8 *si, int leni, const u8 *sj, int lenj,
                      int swaped, rseq *med, rseq *medwg, rseq *ri, rseq *rj) {
    int stride = lenj + 1, i = leni - 1, j = lenj - 1, mode = 0;
    const u16 *dp = dir + ((size_t)leni * stride - 2); /* :1735 == cell (leni-1, lenj-1) */
    int ic = si[i], jc = sj[j];
    while (i != 0 && j != 0) {
        int m = *dp;
        if (mode == 0) {
            mode = pick_mode(m & DO_D, m & DO_A, m & DO_V, m & DO_H, swaped);
        } else if (mode == 1) { /* vertical run (:1553-1572) */
            if (m & (END_V | H_EQ_V)) mode = 0;
            indel_emit(med, medwg, ic);
            rprep(ri, ic); rprep(rj, GAPBIT);
            i--; dp -= stride; ic = si[i];
        } else if (mode == 2) { /* horizontal run (:1531-1550) */
            if (m & (END_H | H_EQ_V)) mode = 0;
            indel_emit(med, medwg, jc);
            rprep(ri, GAPBIT); rprep(rj, jc);
            j--; dp -= 1; jc = sj[j];
        } else if (mode == 3) { /* block diagonal (:1752-1762) */
            if (m & END_B) mode = 0;
            rprep(ri, ic); rprep(rj, jc); rprep(medwg, GAPBIT);
            i--; j--; dp -= stride + 1; jc = sj[j]; ic = si[i];
        } else { /* align (:1763-1781) */
            int prep = c->median[((ic & NOGAP) << 5) + (jc & NOGAP)];
            mode = pick_mode(m & A2D, m & A2A, m & A2V, m & A2H, swaped);
            rprep(med, prep); rprep(medwg, prep);
            rprep(ri, ic); rprep(rj, jc);
            i--; j--; dp -= stride + 1; jc = sj[j]; ic = si[i];
        }
    }
    while (i != 0) { /* :1784-1797 */
        indel_emit(med, medwg, ic);
        rprep(ri, ic); rprep(rj, GAPBIT);
        i--; ic = si[i];
    }
    while (j != 0) { /* :1798-1811 */
        indel_emit(med, medwg, jc);
        rprep(ri, GAPBIT); rprep(rj, jc);
        j--; jc = sj[j];
    }
    rprep(ri, GAPBIT); rprep(rj, GAPBIT); rprep(medwg, GAPBIT);
    if (med->pos == med->cap || med->buf[med->pos] != GAPBIT) rprep(med, GAPBIT); /* :1815 */
}

/* Persistent scratch standing in for the reference's process-global
 * Matrix.default (grow-only, never cleared; src/matrix.ml:34).              */
typedef struct { int *W; u16 *gm[4]; u16 *dir; size_t wcap, gcap, dcap; } do_scratch;

do_scratch *do_scratch_new(void) { return (do_scratch *)calloc(1, sizeof(do_scratch)); }
void do_scratch_free(do_scratch *s) {
    if (s) { int k; free(s->W); for (k = 0; k < 4; k++) free(s->gm[k]); free(s->dir); free(s); }
}

static void scratch_fit(do_scratch *s, int leni, int lenj) {
    size_t w = (size_t)12 * (lenj + 2) + 16, g = (size_t)lenj + 2, d = (size_t)(leni + 1) * (lenj + 2);
    int k;
    if (s->wcap < w) { s->W = (int *)realloc(s->W, w * sizeof(int)); memset(s->W + s->wcap, 0, (w - s->wcap) * sizeof(int)); s->wcap = w; }
    if (s->gcap < g) {
        for (k = 0; k < 4; k++) { s->gm[k] = (u16 *)realloc(s->gm[k], g * sizeof(u16)); memset(s->gm[k] + s->gcap, 0, (g - s->gcap) * sizeof(u16)); }
        s->gcap = g;
    }
    if (s->dcap < d) { s->dir = (u16 *)realloc(s->dir, d * sizeof(u16)); memset(s->dir + s->dcap, 0, (d - s->dcap) * sizeof(u16)); s->dcap = d; }
}

/* requires leni <= lenj (the OCaml caller passes the shorter first and tells us
 * through `swaped` whether it exchanged them, src/sequence.ml:633-649).
 * Output buffers need capacity leni+lenj+2.  lens = {median, medianwg, resi, resj}.
 * Returns the cost, or INT_MIN if leni > lenj ("pass the shorter one as first"). */
int do_align_affine(const do_cm *c, do_scratch *sc, const u8 *si, int leni, const u8 *sj, int lenj,
                    int swaped, u8 *median, u8 *medianwg, u8 *resi, u8 *resj, int *lens,
                    do_align_stats *st) {
    int lasti = leni - 1, lastj = lenj - 1, delta = lastj - lasti, cap = leni + lenj + 2;
    int T, res = 0, j, stride = lastj + 2;
    int *W, *gop, *hext;
    do_align_stats local;
    rseq med = { median, cap, cap }, medwg = { medianwg, cap, cap }, ri = { resi, cap, cap }, rj = { resj, cap, cap };
    if (!st) st = &local;
    if (lenj < leni) return (-2147483647 - 1);
    scratch_fit(sc, leni, lenj);
    W = sc->W; gop = W + 8 * stride + 4; hext = gop + stride + 4;
    st->iterations = 0; st->cells = 0;
    /* initialize_matrices_affine (:1866-1899): its row-0 costs are overwritten by
     * the band fill except when there are no rows at all (leni == 1); the
     * final-cost row it writes is what an empty first sequence returns. */
    res = 0;
    {
        /* Only EB row 0 and EH[0][0] survive the band fill's own row-0 setup; they
         * are NOT refreshed between fills, so from the second fill on row 1 sees
         * whatever the previous fill left in row buffer 0 (kept in sc->W). */
        int *eb0 = W + 2 * stride, *eh0p = W + 6 * stride, eh0 = c->gap_open;
        eh0p[0] = c->gap_open;
        for (j = 0; j <= lastj; j++) eb0[j] = DO_INF;
        for (j = 1; j <= lastj; j++) { eh0 += c->prepend[sj[j]]; }
        if (lastj >= 1) res = eh0;
    }
    sc->dir[0] = 0xFFFF;
    T = (delta + 1) * c->min_non0; /* algn_fill_plane_3_aff :2348-2349 */
    for (;;) { /* algn_newkk_increaseT_aff :2311-2336 */
        int p = (T - delta) / 2, gap_num, newp;
        gap_num = band_fill(c, si, lasti, sj, lastj, p, W, gop, hext, sc->gm, sc->dir, &res, &st->cells);
        st->iterations++;
        st->final_T = T;
        st->final_k = p >= lasti ? lasti - 1 : p;
        newp = (2 * T - delta) / 2;
        if (gap_num < p || newp - lastj + 1 >= 0) break;
        T *= 2;
    }
    traceback(c, sc->dir, si, leni, sj, lenj, swaped, &med, &medwg, &ri, &rj);
    lens[0] = rfinish(&med); lens[1] = rfinish(&medwg); lens[2] = rfinish(&ri); lens[3] = rfinish(&rj);
    return res;
}

/* ---- multi-threaded batch runner (CPU baseline when oracle/_ref is unavailable) --------- */
#include <pthread.h>
#include <time.h>
typedef struct {
    const do_cm *cm; int mode, n; const u8 *seqs; const long long *off_i, *off_j; const int *len_i, *len_j;
    const u8 *swaped; int *cost; volatile int *next;
} do_batch_job;

static void *do_batch_worker(void *arg) {
    do_batch_job *j = (do_batch_job *)arg;
    do_scratch *sc = do_scratch_new();
    int maxl = 0, p;
    u8 *b0, *b1, *b2, *b3;
    for (p = 0; p < j->n; p++) { if (j->len_i[p] > maxl) maxl = j->len_i[p]; if (j->len_j[p] > maxl) maxl = j->len_j[p]; }
    b0 = malloc(2 * maxl + 4); b1 = malloc(2 * maxl + 4); b2 = malloc(2 * maxl + 4); b3 = malloc(2 * maxl + 4);
    for (;;) {
        int lens[4];
        p = __sync_fetch_and_add(j->next, 1);
        if (p >= j->n) break;
        if (j->mode == 0)
            j->cost[p] = do_cost_affine(j->cm, j->seqs + j->off_i[p], j->len_i[p], j->seqs + j->off_j[p], j->len_j[p]);
        else
            j->cost[p] = do_align_affine(j->cm, sc, j->seqs + j->off_i[p], j->len_i[p], j->seqs + j->off_j[p], j->len_j[p],
                                         j->swaped ? j->swaped[p] : 0, b0, b1, b2, b3, lens, NULL);
    }
    free(b0); free(b1); free(b2); free(b3);
    do_scratch_free(sc);
    return NULL;
}

/* mode 0 = cost only, 1 = banded align + traceback.  Returns wall seconds. */
double do_batch_affine(const do_cm *cm, int mode, int n, const u8 *seqs, const long long *off_i, const int *len_i,
                       const long long *off_j, const int *len_j, const u8 *swaped, int *cost, int nthreads) {
    pthread_t th[256];
    do_batch_job jobs[256];
    volatile int next = 0;
    struct timespec t0, t1;
    int t;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < nthreads; t++) {
        do_batch_job jb = { cm, mode, n, seqs, off_i, off_j, len_i, len_j, swaped, cost, &next };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, do_batch_worker, &jobs[t]);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ======================================================================== */
/* 3. linear-gap alignment: algn_CAML_simple_2 -> algn_nw -> algn_fill_plane_2 */
/*    (src/algn.c:3134, 2954-2975, 1134-1177), full plane algn_fill_plane    */
/*    (:927-973, 675-690, 458-533, 654-667), Ukkonen band algn_newkk_*       */
/*    (:1008-1130, 747-923, 543-647, 693-730, 978-1005), backtrace_2d        */
/*    (:3277-3327, 79-124)                                                    */
/* ======================================================================== */
enum { L_ALIGN = 1, L_INSERT = 2, L_DELETE = 4 }; /* src/matrices.h:22-27 */

typedef struct { int *mm; u16 *dm; size_t mcap, dcap; int lenX, lenY; } do_lin_scratch;
do_lin_scratch *do_lin_scratch_new(void) { return (do_lin_scratch *)calloc(1, sizeof(do_lin_scratch)); }
void do_lin_scratch_free(do_lin_scratch *s) { if (s) { free(s->mm); free(s->dm); free(s); } }
static void lin_fit(do_lin_scratch *s, int lenX, int lenY) {
    size_t n = (size_t)(lenX + 1) * (lenY + 1);
    if (s->mcap < n) { s->mm = (int *)realloc(s->mm, n * sizeof(int)); memset(s->mm + s->mcap, 0, (n - s->mcap) * sizeof(int)); s->mcap = n; }
    if (s->dcap < n) { s->dm = (u16 *)realloc(s->dm, n * sizeof(u16)); memset(s->dm + s->dcap, 0, (n - s->dcap) * sizeof(u16)); s->dcap = n; }
    s->lenX = lenX; s->lenY = lenY;
}

/* interior cell (algn_fill_row :458-533): every minimal candidate is recorded */
static void lin_cell(int *mm, const int *pm, u16 *dm, int j, int c_del, int gapc, int algc) {
    int t1 = pm[j] + c_del, t2 = mm[j - 1] + gapc, t3 = pm[j - 1] + algc;
    int m = t1 < t2 ? t1 : t2;
    if (t3 < m) m = t3;
    mm[j] = m;
    dm[j] = (u16)((t1 == m ? L_DELETE : 0) | (t2 == m ? L_INSERT : 0) | (t3 == m ? L_ALIGN : 0));
}
/* right-border cell: no DELETE candidate (algn_fill_ukk_right_cell :543-581) */
static void lin_right(int *mm, const int *pm, u16 *dm, int j, int gapc, int algc) {
    int t2 = mm[j - 1] + gapc, t3 = pm[j - 1] + algc;
    int m = t2 < t3 ? t2 : t3;
    mm[j] = m;
    dm[j] = (u16)((t2 == m ? L_INSERT : 0) | (t3 == m ? L_ALIGN : 0));
}
/* left-border cell: no INSERT candidate (algn_fill_ukk_left_cell :614-647) */
static void lin_left(int *mm, const int *pm, u16 *dm, int j, int c_del, int algc) {
    int t1 = pm[j] + c_del, t3 = pm[j - 1] + algc;
    int m = t1 < t3 ? t1 : t3;
    mm[j] = m;
    dm[j] = (u16)((t1 == m ? L_DELETE : 0) | (t3 == m ? L_ALIGN : 0));
}
/* last column additionally offers tail_cost (algn_fill_last_column :654-667) */
static void lin_last(int *mm, const int *pm, u16 *dm, int l, int tlc) {
    if (l > 0) {
        int cst = tlc + pm[l];
        if (cst < mm[l]) { mm[l] = cst; dm[l] = L_DELETE; }
        else if (cst == mm[l]) dm[l] |= L_DELETE;
    }
}

static int lin_full_plane(const do_cm *c, do_lin_scratch *sc, const u8 *s1, int lenX, const u8 *s2, int lenY) {
    int *mm = sc->mm, *nm = sc->mm, *tmp, i, j;
    u16 *dm = sc->dm;
    mm[0] = 0; dm[0] = L_ALIGN;
    for (j = 1; j < lenY; j++) { mm[j] = mm[j - 1] + c->prepend[s2[j]]; dm[j] = L_INSERT; }
    mm += lenY;
    for (i = 1, dm += lenY; i < lenX; i++, dm += lenY) {
        int a = s1[i], c_del = c->cost[(a << 5) + GAPBIT], tlc = c->tail[a];
        mm[0] = c_del + nm[0]; dm[0] = L_DELETE;                 /* algn_fill_full_row :679-680 */
        for (j = 1; j <= lenY - 1; j++) lin_cell(mm, nm, dm, j, c_del, c->cost[(GAPBIT << 5) + s2[j]], c->cost[(a << 5) + s2[j]]);
        lin_last(mm, nm, dm, lenY - 1, tlc);
        tmp = mm; mm = nm; nm = tmp;
    }
    return nm[lenY - 1];
}

/* gap count of the ALIGN > INSERT > DELETE traceback (backtrace_2d_gaps :978-1005) */
static int lin_trace_gaps(const do_lin_scratch *sc, int lenX, int lenY) {
    int nd = 0, ni = 0;
    long pos = (long)lenY * (lenX - 1) + lenY - 1;
    while (pos >= 0) {
        int d = sc->dm[pos];
        if (d & L_ALIGN) pos -= lenY + 1;
        else if (d & L_INSERT) { ni++; pos -= 1; }
        else { nd++; pos -= lenY; }
    }
    return nd > ni ? nd : ni;
}

static int lin_band_fill(const do_cm *c, do_lin_scratch *sc, const u8 *s1, int lenX, const u8 *s2, int lenY, int p,
                         int *cost, long long *cells) {
    int k = p >= lenX ? lenX - 1 : p, delta = lenY - lenX, i, j;
    int first = delta + 1 + p;
    int *a = sc->mm, *b = sc->mm + lenY;
    u16 *dm = sc->dm;
    if (first > lenY) first = lenY;
    a[0] = 0; dm[0] = L_ALIGN;                                   /* algn_fill_first_row :693-717 */
    for (j = 1; j < first; j++) { a[j] = a[j - 1] + c->prepend[s2[j]]; dm[j] = L_INSERT; }
    for (i = 1; i < lenX; i++) {
        int sym = s1[i], c_del = c->cost[(sym << 5) + GAPBIT], tlc = c->tail[sym];
        int left = (i - k) > 0, startj = left ? i - k : 0;
        int right = (i + delta + k) <= (lenY - 1), endj = right ? i + delta + k : lenY - 1;
        int len = endj - startj + 1;
        u16 *d = dm + (size_t)i * lenY;
#define GAPC(j) (c->cost[(GAPBIT << 5) + s2[j]])
#define ALGC(j) ((j) == 0 ? c->tail[sym] : c->cost[(sym << 5) + s2[j]])
        if (left && right) {                                      /* algn_fill_extending_left_right :786-831 */
            if (len == 1) { b[startj] = a[startj - 1] + ALGC(startj); d[startj] = L_ALIGN; }
            else {
                lin_left(b, a, d, startj, c_del, ALGC(startj));
                for (j = startj + 1; j <= startj + len - 2; j++) lin_cell(b, a, d, j, c_del, GAPC(j), ALGC(j));
                lin_right(b, a, d, startj + len - 1, GAPC(startj + len - 1), ALGC(startj + len - 1));
            }
        } else if (right) {                                       /* algn_fill_extending_right :746-784 */
            b[0] = a[0] + ALGC(0); d[0] = L_DELETE;               /* first cell uses alg_row[0] = tail (A9) */
            for (j = 1; j <= len - 2; j++) lin_cell(b, a, d, j, c_del, GAPC(j), ALGC(j));
            lin_right(b, a, d, len - 1, GAPC(len - 1), ALGC(len - 1));
        } else if (left) {                                        /* algn_fill_extending_left :833-886 */
            lin_left(b, a, d, startj, c_del, ALGC(startj));
            for (j = startj + 1; j <= startj + len - 1; j++) lin_cell(b, a, d, j, c_del, GAPC(j), ALGC(j));
            lin_last(b, a, d, startj + len - 1, tlc);
        } else {                                                  /* algn_fill_no_extending :888-923 */
            b[0] = a[0] + ALGC(0); d[0] = L_DELETE;
            for (j = 1; j <= lenY - 1; j++) lin_cell(b, a, d, j, c_del, GAPC(j), ALGC(j));
            lin_last(b, a, d, lenY - 1, tlc);
        }
#undef GAPC
#undef ALGC
        *cells += len;
        a = b;
        if (i < lenX - 1) b = a + lenY;
    }
    *cost = a[lenY - 1];
    return lin_trace_gaps(sc, lenX, lenY);
}

/* s1 must be the shorter sequence (the OCaml caller guarantees it, src/sequence.ml:917-925).
 * Leaves the direction matrix in `sc` for do_backtrace_linear, like the reference leaves it in
 * Matrix.default between c_cost_2 and extract_edited_2.  mode_out: 0 full plane, 1 Ukkonen. */
int do_cost_linear(const do_cm *c, do_lin_scratch *sc, const u8 *s1, int lenX, const u8 *s2, int lenY, int deltawh,
                   do_align_stats *st) {
    int height, T, cost = 0;
    do_align_stats local;
    if (!st) st = &local;
    st->iterations = 0; st->cells = 0; st->final_T = 0; st->final_k = -1;
    if (lenX > lenY) return (-2147483647 - 1);
    lin_fit(sc, lenX, lenY);
    height = (lenX - lenY) + 50 + deltawh;                         /* algn_nw_limit :2963, algn_fill_plane_2 :1141-1144 */
    if (height > lenX) height = lenX;
    T = (lenY - lenX + 1) * c->min_non0;
    if ((float)lenX >= 1.5f * (float)lenY) return lin_full_plane(c, sc, s1, lenX, s2, lenY);
    if (!((2 * height) < lenX) && 8 >= (lenX - height)) return lin_full_plane(c, sc, s1, lenX, s2, lenY);
    for (;;) {                                                      /* algn_newkk_increaseT :1117-1130 */
        int p = (T - (lenY - lenX)) / 2, newp, gap_num;
        gap_num = lin_band_fill(c, sc, s1, lenX, s2, lenY, p, &cost, &st->cells);
        st->iterations++; st->final_T = T; st->final_k = p >= lenX ? lenX - 1 : p;
        newp = (2 * T - (lenY - lenX)) / 2;
        if ((gap_num + 1) < p || newp - lenY + 1 >= 0) return cost;
        T *= 2;
    }
}

/* backtrace_2d (:3277-3327): ALIGN first, then DELETE before INSERT, or INSERT before DELETE when
 * the caller swapped the operands.  Outputs need capacity lenX+lenY; lens = {r1, r2}. */
void do_backtrace_linear(const do_lin_scratch *sc, const u8 *s1, const u8 *s2, int swaped, u8 *r1, u8 *r2, int *lens) {
    int lenX = sc->lenX, lenY = sc->lenY, cap = lenX + lenY, a1 = lenX, a2 = lenY;
    long pos = (long)lenY * (lenX - 1) + lenY - 1;
    rseq o1 = { r1, cap, cap }, o2 = { r2, cap, cap };
    while (pos >= 0) {
        int d = sc->dm[pos];
        if (d & L_ALIGN) { rprep(&o1, s1[--a1]); rprep(&o2, s2[--a2]); pos -= lenY + 1; }
        else {
            int ins = swaped ? ((d & L_INSERT) != 0) : !(d & L_DELETE);
            if (ins) { rprep(&o1, GAPBIT); rprep(&o2, s2[--a2]); pos -= 1; }
            else { rprep(&o1, s1[--a1]); rprep(&o2, GAPBIT); pos -= lenY; }
        }
    }
    lens[0] = rfinish(&o1); lens[1] = rfinish(&o2);
}

/* ======================================================================== */
/* 4. O(L) column-wise helpers                                               */
/* ======================================================================== */
/* seq_get_median_2d_with_gaps / _no_gaps (src/seq.c:241-272) */
int do_median_2(const do_cm *c, const u8 *a, const u8 *b, int len, int with_gaps, u8 *out) {
    int i, n = 0;
    u8 *tmp = (u8 *)malloc((size_t)len + 2);
    if (!with_gaps) tmp[n++] = GAPBIT;
    for (i = 0; i < len; i++) {
        int m = c->median[(a[i] << 5) + b[i]];
        if (with_gaps || m != GAPBIT) tmp[n++] = (u8)m;
    }
    memcpy(out, tmp, (size_t)n);
    free(tmp);
    return n;
}
/* algn_union (src/algn.c:3657-3666) */
void do_union(const u8 *a, const u8 *b, int len, u8 *out) { int i; for (i = 0; i < len; i++) out[i] = a[i] | b[i]; }

/* algn_calculate_from_2_aligned (src/algn.c:3003-3090), bitset alphabet branch: cost of an aligned
 * pair under `table` (cost -> algn_verify_2, worst -> algn_worst_2) with the gap-opening automaton */
static int from_2_aligned(const do_cm *c, const int *table, const u8 *s1, const u8 *s2, int len) {
    int i, res = 0, go = c->gap_open, gap_row = 0;
    i = ((s1[0] & GAPBIT) && (s2[0] & GAPBIT)) ? 1 : 0;
    for (; i < len; i++) {
        int a = s1[i], b = s2[i];
        if (gap_row == 0) {
            if ((a & GAPBIT) && !(b & GAPBIT)) { res += go; gap_row = 1; }
            else if ((b & GAPBIT) && !(a & GAPBIT)) { res += go; gap_row = 2; }
        } else if (gap_row == 1) {
            if (!(a & GAPBIT)) {
                if ((b & GAPBIT) && !(a & GAPBIT)) { res += go; gap_row = 2; }
                else gap_row = 0;
            }
        } else {
            if (!(b & GAPBIT)) {
                if (a & GAPBIT) { res += go; gap_row = 1; }
                else gap_row = 0;
            }
        }
        res += table[(a << 5) + b];
    }
    return res;
}
int do_worst_2(const do_cm *c, const u8 *a, const u8 *b, int len) { return from_2_aligned(c, c->worst, a, b, len); }
int do_verify_2(const do_cm *c, const u8 *a, const u8 *b, int len) { return from_2_aligned(c, c->cost, a, b, len); }

/* algn_ancestor_2 (src/algn.c:3603-3626) with algn_correct_blocks_affine (:3561-3601) and
 * algn_remove_gaps (:3539-3559); combinations = 1 */
int do_ancestor_2(const do_cm *c, const u8 *s1, const u8 *s2, int len, u8 *out) {
    int i, n = 0, gap = GAPBIT;
    u8 *sm = (u8 *)malloc((size_t)len + 2);
    for (i = 0; i < len; i++) {
        int m = c->median[(s1[i] << 5) + s2[i]];
        if (m == 0) { free(sm); return (-2147483647 - 1); }       /* failwith "median should not be 0" */
        sm[i] = (u8)m;
    }
    if (c->cost_model_type != 1) {
        /* non-affine: pure-gap medians are dropped while prepending, then the leading gap is restored */
        out[n++] = (u8)gap;
        for (i = 0; i < len; i++) if (sm[i] != gap) out[n++] = sm[i];
    } else {
        int extending = 0, inside = 0, prev_block = 0;
        for (i = 0; i < len; i++) {
            int ab = s1[i], bb = s2[i], sb = sm[i];
            if (!inside && (!(ab & gap) || !(bb & gap))) inside = 0;
            else if (inside && (!(ab & gap) || !(bb & gap))) inside = 0;
            else if (((ab & gap) || (bb & gap)) && ((ab != gap) || (bb != gap))) inside = 1;
            else inside = 0;
            if (((gap & ab) || (gap & bb)) && !(sb & gap) && !extending) { prev_block = inside; extending = 1; }
            else if ((gap & ab) && (gap & bb) && (sb & gap) && (sb != gap) && extending && inside && !prev_block) { sb = (~gap) & sb; prev_block = 0; }
            else if ((gap & ab) && (gap & bb) && extending == 1) { prev_block = inside; extending = 0; }
            sm[i] = (u8)sb;
        }
        out[n++] = (u8)gap;                                         /* algn_remove_gaps restores the leading gap */
        for (i = 0; i < len; i++) if (sm[i] != gap) out[n++] = sm[i];
    }
    free(sm);
    return n;
}
