/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from
 * the product path (poy5_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library this
 * file is built into (oracle/_ref/libpoyref*.so).
 *
 * Thin plain-C driver around the UNMODIFIED reference C of amnh/poy5, compiled
 * in place from /root/reference/src (algn.c textually includes matrices.c,
 * cm.c, seq.c, array_pool.c, union.c) against the OCaml-runtime shim in
 * oracle/shim/caml.  No reference source is copied into this repository; this
 * file only calls the reference's own `*_CAML_*` entry points, exactly as the
 * OCaml `external` declarations in src/sequence.ml:613-631, src/matrix.ml:23-29
 * and src/cost_matrix.ml:44-90 would.
 *
 * Handles are opaque `void*` (the shim's calloc'd custom blocks).  Sequences
 * cross this boundary as plain uint8 arrays in reading order, element 0 being
 * the leading gap code (SURVEY.md section 8 notation).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <pthread.h>
#include <time.h>
#include <assert.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include "seq.h"
#include "matrices.h"
#include "cm.h"

/* ---- reference entry points (defined in the reference objects) ---------- */
value seq_CAML_create(value cap);
value seq_CAML_prepend(value s, value v);
value mat_CAML_create_general(value a);
value mat_CAML_flush_memory(value vm);
value cm_CAML_create(value a_sz, value combine, value aff, value go, value all, value level,
                     value combine_number, value gap_start, value tie_breaker);
value cm_CAML_set_cost(value a, value b, value c, value v);
value cm_CAML_set_worst(value a, value b, value c, value v);
value cm_CAML_set_median(value a, value b, value c, value v);
value cm_CAML_set_prepend(value a, value b, value v);
value cm_CAML_set_tail(value a, value b, value v);
value cm_CAML_set_affine(value c, value do_aff, value go);
value cm_CAML_create_3d(value a_sz, value combine, value aff, value go, value d, value all,
                        value level, value map_sz, value gap_start);
value cm_CAML_set_cost_3d(value a, value b, value c, value cc, value v);
value cm_CAML_set_median_3d(value a, value b, value c, value cp, value v);
value algn_CAML_cost_affine_3(value si, value sj, value cm, value am);
value algn_CAML_align_affine_3(value si, value sj, value cm, value am, value resi, value resj,
                               value median, value medianwg, value swaped);
value algn_CAML_simple_2(value s1, value s2, value c, value a, value deltawh);
value algn_CAML_backtrace_2d(value s1, value s2, value s1p, value s2p, value a, value c, value swap);
value algn_CAML_align_2d(value s1, value s2, value c, value a, value s1p, value s2p, value deltawh,
                         value swaped);
value algn_CAML_ancestor_2(value sa, value sb, value cm, value sab);
value algn_CAML_union(value s1, value s2, value su);
value algn_CAML_worst_2(value s1, value s2, value c);
value algn_CAML_verify_2(value s1, value s2, value c);
value algn_CAML_myers(value sa, value sb);
value algn_CAML_align_3d(value s1, value s2, value s3, value c, value a, value s1p, value s2p,
                         value s3p, value uk);
value seq_CAML_median_2_no_gaps(value s1, value s2, value m, value sm);
value seq_CAML_median_2_with_gaps(value s1, value s2, value m, value sm);
int cm_get_min_non0_cost(cmt c);

/* ---- failwith -> error code -------------------------------------------- */
/* shared with ref_newkk_driver.c (same shared object) */
__thread jmp_buf *ref_jmp_shared = NULL;
#define ref_jmp ref_jmp_shared
static __thread char ref_errmsg[512];

void caml_failwith(const char *msg) {
    strncpy(ref_errmsg, msg ? msg : "", sizeof(ref_errmsg) - 1);
    ref_errmsg[sizeof(ref_errmsg) - 1] = 0;
    if (ref_jmp) longjmp(*ref_jmp, 1);
    fprintf(stderr, "poyref: uncaught Failure: %s\n", ref_errmsg);
    abort();
}
void caml_invalid_argument(const char *msg) { caml_failwith(msg); }
void caml_raise_out_of_memory(void) { caml_failwith("Out of memory"); }
const char *ref_last_error(void) { return ref_errmsg; }

#define REF_FAIL (-2147483647 - 1)
#define GUARD_BEGIN jmp_buf jb; jmp_buf *saved = ref_jmp; ref_jmp = &jb; if (setjmp(jb) == 0) {
#define GUARD_END(failval) ref_jmp = saved; } else { ref_jmp = saved; return (failval); }

int ref_long_sequences(void) {
#ifdef USE_LONG_SEQUENCES
    return 1;
#else
    return 0;
#endif
}

/* ---- helpers ------------------------------------------------------------ */
static value make_seq(const unsigned char *s, int len, int cap) {
    value v = seq_CAML_create(Val_int(cap));
    int i;
    for (i = len - 1; i >= 0; i--) seq_CAML_prepend(v, Val_int(s[i]));
    return v;
}
static int read_seq(value v, unsigned char *out) {
    seqt s;
    Seq_custom_val(s, v);
    if (out) memcpy(out, s->begin, (size_t)s->len);
    return s->len;
}
static void free_val(value v) { free((void *)v); }

/* OCaml-value sequences for tests that drive `*_CAML_*`-shaped entry points directly (tests/test_caml_stubs.py) */
void *ref_seq_new(const unsigned char *s, int len, int cap) { return (void *)make_seq(s, len, cap); }
int ref_seq_read(void *v, unsigned char *out) { return read_seq((value)v, out); }
void ref_val_free(void *v) { free(v); }
long ref_val_int(int x) { return (long)Val_int(x); }
int ref_int_val(long v) { return Int_val((value)v); }

/* ---- scratch matrices (Matrix.default, src/matrix.ml:34) ---------------- */
void *ref_mat_new(void) { return (void *)mat_CAML_create_general(Val_int(0)); }
void ref_mat_free(void *m) { mat_CAML_flush_memory((value)m); free(m); }
/* Pre-grow the scratch like a long-running POY process would have: the cost
 * entry point needs 12*largest ints but sizes for (largest+1)^2 (overflow for
 * largest < 10 on a fresh scratch, src/algn.c:2485-2494). */
void ref_mat_reserve(void *m, int w, int d) {
    mat_setup_size(Matrices_struct((value)m), w, d, 0, 0, 5, 0);
}

/* ---- 2-D cost matrix: 5-letter bitset alphabet with all combinations ---- */
/* Mirrors Cost_matrix.Two_D.create 5 true model go 31 0 31?.. : the OCaml side
 * calls `create a_sz use_comb model go all_elements level num_comb gap_start tb`
 * (src/cost_matrix.ml:1162-1165).  For level=0 num_comb is irrelevant to the
 * table layout (bit-indexed 32x32x2). */
void *ref_cm_new(int a_sz, int cost_model_type, int gap_open, int all_elements) {
    int ncomb = (1 << a_sz) - 1;
    GUARD_BEGIN
    value v = cm_CAML_create(Val_int(a_sz), Val_int(1), Val_int(cost_model_type), Val_int(gap_open),
                             Val_int(all_elements), Val_int(0), Val_int(ncomb), Val_int(0), Val_int(1));
    ref_jmp = saved;
    return (void *)v;
    GUARD_END(NULL)
}
void ref_cm_free(void *cm) {
    cmt c = Cost_matrix_struct((value)cm);
    free(c->combmap); free(c->comb2list); free(c->cost); free(c->worst);
    free(c->prepend_cost); free(c->tail_cost); free(c->median);
    free(cm);
}
void ref_cm_set_affine(void *cm, int model, int go) { cm_CAML_set_affine((value)cm, Val_int(model), Val_int(go)); }
void ref_cm_set_cost(void *cm, int a, int b, int v) { cm_CAML_set_cost(Val_int(a), Val_int(b), (value)cm, Val_int(v)); }
void ref_cm_set_worst(void *cm, int a, int b, int v) { cm_CAML_set_worst(Val_int(a), Val_int(b), (value)cm, Val_int(v)); }
void ref_cm_set_median(void *cm, int a, int b, int v) { cm_CAML_set_median(Val_int(a), Val_int(b), (value)cm, Val_int(v)); }
void ref_cm_set_prepend(void *cm, int a, int v) { cm_CAML_set_prepend(Val_int(a), Val_int(v), (value)cm); }
void ref_cm_set_tail(void *cm, int a, int v) { cm_CAML_set_tail(Val_int(a), Val_int(v), (value)cm); }
int ref_cm_min_non0(void *cm) { return cm_get_min_non0_cost(Cost_matrix_struct((value)cm)); }
/* bulk load of n x n tables indexed [a*n+b], a,b in 1..n-1 (n = 1<<a_sz) */
void ref_cm_load(void *cm, int n, const int *cost, const int *worst, const unsigned char *median,
                 const int *prepend, const int *tail) {
    int a, b;
    for (a = 1; a < n; a++) {
        for (b = 1; b < n; b++) {
            ref_cm_set_cost(cm, a, b, cost[a * n + b]);
            ref_cm_set_worst(cm, a, b, worst[a * n + b]);
            ref_cm_set_median(cm, a, b, median[a * n + b]);
        }
        ref_cm_set_prepend(cm, a, prepend[a]);
        ref_cm_set_tail(cm, a, tail[a]);
    }
}

/* ---- affine entry points (src/algn.c:2457, 2359) ------------------------ */
int ref_cost_affine(void *cm, void *mat, const unsigned char *s1, int len1, const unsigned char *s2,
                    int len2) {
    value a = make_seq(s1, len1, len1), b = make_seq(s2, len2, len2);
    int res;
    GUARD_BEGIN
    res = Int_val(algn_CAML_cost_affine_3(a, b, (value)cm, (value)mat));
    GUARD_END((free_val(a), free_val(b), REF_FAIL))
    free_val(a); free_val(b);
    return res;
}

/* out buffers must hold leni+lenj+2 bytes each; lens[4] = median, medianwg, resi, resj */
int ref_align_affine(void *cm, void *mat, const unsigned char *si, int leni, const unsigned char *sj,
                     int lenj, int swaped, unsigned char *median, unsigned char *medianwg,
                     unsigned char *resi, unsigned char *resj, int *lens) {
    int cap = leni + lenj + 2, res;
    value a = make_seq(si, leni, leni), b = make_seq(sj, lenj, lenj);
    value vm = seq_CAML_create(Val_int(cap)), vw = seq_CAML_create(Val_int(cap));
    value vi = seq_CAML_create(Val_int(cap)), vj = seq_CAML_create(Val_int(cap));
    GUARD_BEGIN
    res = Int_val(algn_CAML_align_affine_3(a, b, (value)cm, (value)mat, vi, vj, vm, vw, Val_int(swaped)));
    GUARD_END((free_val(a), free_val(b), free_val(vm), free_val(vw), free_val(vi), free_val(vj), REF_FAIL))
    lens[0] = read_seq(vm, median);
    lens[1] = read_seq(vw, medianwg);
    lens[2] = read_seq(vi, resi);
    lens[3] = read_seq(vj, resj);
    free_val(a); free_val(b); free_val(vm); free_val(vw); free_val(vi); free_val(vj);
    return res;
}

/* ---- linear-gap entry points (src/algn.c:3134, 3424, 3500) --------------- */
int ref_cost_linear(void *cm, void *mat, const unsigned char *s1, int len1, const unsigned char *s2,
                    int len2, int deltawh) {
    value a = make_seq(s1, len1, len1), b = make_seq(s2, len2, len2);
    int res;
    GUARD_BEGIN
    res = Int_val(algn_CAML_simple_2(a, b, (value)cm, (value)mat, Val_int(deltawh)));
    GUARD_END((free_val(a), free_val(b), REF_FAIL))
    free_val(a); free_val(b);
    return res;
}
/* s1 is the LONGER sequence here (src/sequence.ml:1019-1033). lens[2] = r1, r2 */
int ref_align_linear(void *cm, void *mat, const unsigned char *s1, int len1, const unsigned char *s2,
                     int len2, int deltawh, int swaped, unsigned char *r1, unsigned char *r2, int *lens) {
    int cap = len1 + len2 + 2, res;
    value a = make_seq(s1, len1, len1), b = make_seq(s2, len2, len2);
    value v1 = seq_CAML_create(Val_int(cap)), v2 = seq_CAML_create(Val_int(cap));
    GUARD_BEGIN
    res = Int_val(algn_CAML_align_2d(a, b, (value)cm, (value)mat, v1, v2, Val_int(deltawh), Val_int(swaped)));
    GUARD_END((free_val(a), free_val(b), free_val(v1), free_val(v2), REF_FAIL))
    lens[0] = read_seq(v1, r1);
    lens[1] = read_seq(v2, r2);
    free_val(a); free_val(b); free_val(v1); free_val(v2);
    return res;
}

/* ---- O(L) column-wise helpers ------------------------------------------- */
int ref_ancestor_2(void *cm, const unsigned char *a, const unsigned char *b, int len, unsigned char *out) {
    value va = make_seq(a, len, len), vb = make_seq(b, len, len), vo = seq_CAML_create(Val_int(len + 2));
    int n;
    GUARD_BEGIN
    algn_CAML_ancestor_2(va, vb, (value)cm, vo);
    GUARD_END((free_val(va), free_val(vb), free_val(vo), REF_FAIL))
    n = read_seq(vo, out);
    free_val(va); free_val(vb); free_val(vo);
    return n;
}
int ref_median_2(void *cm, const unsigned char *a, const unsigned char *b, int len, int with_gaps,
                 unsigned char *out) {
    value va = make_seq(a, len, len), vb = make_seq(b, len, len), vo = seq_CAML_create(Val_int(len + 2));
    int n;
    if (with_gaps) seq_CAML_median_2_with_gaps(va, vb, (value)cm, vo);
    else seq_CAML_median_2_no_gaps(va, vb, (value)cm, vo);
    n = read_seq(vo, out);
    free_val(va); free_val(vb); free_val(vo);
    return n;
}
int ref_union(const unsigned char *a, const unsigned char *b, int len, unsigned char *out) {
    value va = make_seq(a, len, len), vb = make_seq(b, len, len), vo = seq_CAML_create(Val_int(len + 1));
    int n;
    algn_CAML_union(va, vb, vo);
    n = read_seq(vo, out);
    free_val(va); free_val(vb); free_val(vo);
    return n;
}
int ref_worst_2(void *cm, const unsigned char *a, const unsigned char *b, int len) {
    value va = make_seq(a, len, len), vb = make_seq(b, len, len);
    int r = Int_val(algn_CAML_worst_2(va, vb, (value)cm));
    free_val(va); free_val(vb);
    return r;
}
int ref_verify_2(void *cm, const unsigned char *a, const unsigned char *b, int len) {
    value va = make_seq(a, len, len), vb = make_seq(b, len, len);
    int r = Int_val(algn_CAML_verify_2(va, vb, (value)cm));
    free_val(va); free_val(vb);
    return r;
}

/* algn_CAML_myers (src/algn.c:3697-3743); keeps static state between calls, see scripts/ref_myers_probe.py */
int ref_myers(const unsigned char *a, int la, const unsigned char *b, int lb) {
    value va = make_seq(a, la, la), vb = make_seq(b, lb, lb);
    int r = Int_val(algn_CAML_myers(va, vb));
    free_val(va); free_val(vb);
    return r;
}

/* ---- 3-D cube (src/algn.c:2977, 3345, 3500) ------------------------------ */
void *ref_cm3d_new(int a_sz, int all_elements) {
    int ncomb = (1 << a_sz) - 1;
    GUARD_BEGIN
    value v = cm_CAML_create_3d(Val_int(a_sz), Val_int(1), Val_int(0), Val_int(0), Val_int(3),
                                Val_int(all_elements), Val_int(0), Val_int(ncomb), Val_int(0));
    ref_jmp = saved;
    return (void *)v;
    GUARD_END(NULL)
}
void ref_cm3d_set(void *cm, int a, int b, int c, int cost, int median) {
    cm_CAML_set_cost_3d(Val_int(a), Val_int(b), Val_int(c), (value)cm, Val_int(cost));
    cm_CAML_set_median_3d(Val_int(a), Val_int(b), Val_int(c), (value)cm, Val_int(median));
}
int ref_align_3d(void *cm3, void *mat, const unsigned char *s1, int len1, const unsigned char *s2, int len2,
                 const unsigned char *s3, int len3, unsigned char *r1, unsigned char *r2,
                 unsigned char *r3, int *lens) {
    int cap = len1 + len2 + len3 + 3, res;
    value a = make_seq(s1, len1, len1), b = make_seq(s2, len2, len2), c = make_seq(s3, len3, len3);
    value v1 = seq_CAML_create(Val_int(cap)), v2 = seq_CAML_create(Val_int(cap)), v3 = seq_CAML_create(Val_int(cap));
    GUARD_BEGIN
    res = Int_val(algn_CAML_align_3d(a, b, c, (value)cm3, (value)mat, v1, v2, v3, Val_int(0)));
    GUARD_END((free_val(a), free_val(b), free_val(c), free_val(v1), free_val(v2), free_val(v3), REF_FAIL))
    lens[0] = read_seq(v1, r1);
    lens[1] = read_seq(v2, r2);
    lens[2] = read_seq(v3, r3);
    free_val(a); free_val(b); free_val(c); free_val(v1); free_val(v2); free_val(v3);
    return res;
}

/* ---- multi-threaded batch runners for the CPU baseline -------------------
 * One scratch per worker thread (the reference's Matrix.default is process
 * global and not re-entrant; Parmap forks -- here each thread owns a private
 * `matrices` object, which is equivalent).  Pairs are laid out in one packed
 * byte buffer: pair p uses seqs[off_i[p] .. +len_i[p]) and seqs[off_j[p] ..).
 */
typedef struct {
    void *cm; int mode; int n; const unsigned char *seqs;
    const long long *off_i, *off_j; const int *len_i, *len_j; const unsigned char *swaped;
    int *cost; int tid, nthreads; volatile int *next;
    long long medians_bytes;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    void *mat = ref_mat_new();
    int maxl = 0, p;
    unsigned char *b0 = NULL, *b1 = NULL, *b2 = NULL, *b3 = NULL;
    for (p = 0; p < j->n; p++) { if (j->len_i[p] > maxl) maxl = j->len_i[p]; if (j->len_j[p] > maxl) maxl = j->len_j[p]; }
    if (maxl < 16) ref_mat_reserve(mat, 32, 32);
    if (j->mode == 1) {
        b0 = malloc(2 * maxl + 4); b1 = malloc(2 * maxl + 4); b2 = malloc(2 * maxl + 4); b3 = malloc(2 * maxl + 4);
    }
    for (;;) {
        int lens[4];
        p = __sync_fetch_and_add(j->next, 1);
        if (p >= j->n) break;
        if (j->mode == 0)
            j->cost[p] = ref_cost_affine(j->cm, mat, j->seqs + j->off_i[p], j->len_i[p], j->seqs + j->off_j[p], j->len_j[p]);
        else
            j->cost[p] = ref_align_affine(j->cm, mat, j->seqs + j->off_i[p], j->len_i[p], j->seqs + j->off_j[p],
                                          j->len_j[p], j->swaped ? j->swaped[p] : 0, b0, b1, b2, b3, lens);
    }
    free(b0); free(b1); free(b2); free(b3);
    ref_mat_free(mat);
    return NULL;
}

/* mode 0 = algn_CAML_cost_affine_3, 1 = algn_CAML_align_affine_3. Returns wall seconds. */
double ref_batch_affine(void *cm, int mode, int n, const unsigned char *seqs, const long long *off_i,
                        const int *len_i, const long long *off_j, const int *len_j,
                        const unsigned char *swaped, int *cost, int nthreads) {
    pthread_t th[256];
    batch_job jobs[256];
    volatile int next = 0;
    struct timespec t0, t1;
    int t;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < nthreads; t++) {
        batch_job jb = { cm, mode, n, seqs, off_i, off_j, len_i, len_j, swaped, cost, t, nthreads, &next, 0 };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
