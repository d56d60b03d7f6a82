/* Reference-shaped single-pair entry points: the SAME symbols and OCaml-value signatures as the hot stubs of
 * amnh/poy5's libpoycside, implemented as a batch of ONE through libpoy5b200.so -- so that the OCaml side
 * (`external cost_2_affine : s -> s -> Cost_matrix.Two_D.m -> Matrix.m -> int = "algn_CAML_cost_affine_3"`,
 * `external align_affine_3 : ... = "algn_CAML_align_affine_3_bc" "algn_CAML_align_affine_3"`, src/sequence.ml:613-631)
 * links unchanged against this object instead of the alignment half of src/algn.c:
 *
 *   algn_CAML_cost_affine_3       src/algn.c:2457-2515
 *   algn_CAML_align_affine_3[_bc] src/algn.c:2359-2455
 *
 * Built inside the POY source tree (it needs the OCaml runtime headers and the reference's seq.h / cm.h for the layout
 * of the custom blocks; nothing else of the reference is used):
 *   gcc -O2 -fPIC -shared -I`ocamlc -where` -I$POY/src -I$REPO/include poy_caml_stubs.c -L$REPO/poy5_b200 -lpoy5b200
 * The batch entry points (include/poy5_b200.h) are what a tree search should call at the Parmap seam; these wrappers
 * exist so that every other caller of the two stubs keeps working, one alignment per call (upload + launch + read
 * back: latency bound, no throughput claim).
 *
 * Like the reference's `Matrix.default` scratch, the context is one process-wide object created on first use; the
 * `am` (Matrix.m) argument is accepted and ignored.  Failures raise OCaml `Failure` with the library's message
 * (the reference's own text for the conditions it checks).  */
#include <string.h>
#include <stdlib.h>
#include <assert.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include "seq.h"
#include "cm.h"
#include "poy5_b200.h"

static poy_ctx *g_ctx = NULL;
static poy_cm *g_cm = NULL;
static poy_cm_host g_cm_img;

static void stub_fail(const char *msg) { caml_failwith((char *)msg); }

static poy_ctx *the_ctx(void) {
    if (!g_ctx) {
        const char *dev = getenv("POY_CUDA_DEVICE");
        poy_status s = poy_ctx_create(dev ? atoi(dev) : 0, NULL, &g_ctx);
        if (s != POY_OK) stub_fail(poy_status_string(s));
    }
    return g_ctx;
}

/* struct cm (src/cm.h:33-76) -> device cost model, re-uploaded only when the tables changed (the OCaml side mutates
 * a matrix through the cm_CAML_set_* setters, then uses it for thousands of alignments) */
static poy_cm *the_cm(const struct cm *c) {
    poy_cm_host h;
    if (c->lcm != 5 || c->combinations == 0 || c->level > 1)
        stub_fail("poy5_b200: only the 5-letter bitset alphabet with combinations (DNA) is served by the CUDA path");
    memset(&h, 0, sizeof h);
    memcpy(h.cost, c->cost, sizeof h.cost);
    if (c->worst) memcpy(h.worst, c->worst, sizeof h.worst);
    memcpy(h.median, c->median, sizeof h.median);
    if (c->prepend_cost) memcpy(h.prepend, c->prepend_cost, sizeof h.prepend);
    if (c->tail_cost) memcpy(h.tail, c->tail_cost, sizeof h.tail);
    h.gap_open = c->gap_open; h.cost_model_type = c->cost_model_type;
    h.is_identity = c->is_identity; h.is_metric = c->is_metric;
    if (!g_cm || memcmp(&h, &g_cm_img, sizeof h) != 0) {
        poy_ctx *ctx = the_ctx();
        if (g_cm) { poy_cm_free(ctx, g_cm); g_cm = NULL; }
        if (poy_cm_upload(ctx, &h, &g_cm) != POY_OK) stub_fail(poy_last_error(ctx));
        g_cm_img = h;
    }
    return g_cm;
}

static poy_pool *pool_of_two(const struct seq *a, const struct seq *b) {
    poy_ctx *ctx = the_ctx();
    poy_pool *pool = NULL;
    int64_t off[3];
    uint8_t *buf = (uint8_t *)malloc((size_t)a->len + (size_t)b->len + 1);
    if (!buf) stub_fail("Out of memory");
    memcpy(buf, a->begin, (size_t)a->len);
    memcpy(buf + a->len, b->begin, (size_t)b->len);
    off[0] = 0; off[1] = a->len; off[2] = (int64_t)a->len + b->len;
    if (poy_pool_upload(ctx, buf, off, 2, &pool) != POY_OK) { free(buf); stub_fail(poy_last_error(ctx)); }
    free(buf);
    return pool;
}

/* seq_prepend (src/seq.c) on the custom block: the caller allocated the capacity */
static void prepend_all(struct seq *s, const uint8_t *src, int n) {
    int x;
    if (s->len + n > s->cap) stub_fail("poy5_b200: result sequence too short");
    for (x = n - 1; x >= 0; x--) { s->begin = s->begin - 1; *(s->begin) = src[x]; s->len = s->len + 1; }
}

value algn_CAML_cost_affine_3(value si, value sj, value cm, value am) {
    CAMLparam4(si, sj, cm, am);
    struct seq *a, *b;
    const struct cm *c = Cost_matrix_struct(cm);
    poy_ctx *ctx = the_ctx();
    poy_cm *dcm = the_cm(c);
    poy_pool *pool;
    int32_t ia = 0, ib = 1, cost = 0;
    poy_status s;
    Seq_custom_val(a, si);
    Seq_custom_val(b, sj);
    pool = pool_of_two(a, b);
    s = poy_batch_cost_affine(ctx, dcm, pool, 1, &ia, &ib, &cost);     /* either order, like the reference (:2496-2513) */
    poy_pool_free(ctx, pool);
    if (s != POY_OK) stub_fail(poy_last_error(ctx));
    CAMLreturn(Val_int(cost));
}

value algn_CAML_align_affine_3(value si, value sj, value cm, value am, value resi, value resj, value median,
                               value medianwg, value swaped) {
    CAMLparam5(si, sj, cm, am, resi);
    CAMLxparam4(resj, median, medianwg, swaped);
    struct seq *a, *b, *ri, *rj, *md, *mw;
    const struct cm *c = Cost_matrix_struct(cm);
    poy_ctx *ctx = the_ctx();
    poy_cm *dcm = the_cm(c);
    poy_pool *pool;
    int32_t i0 = 0, i1 = 1, cost = 0, lens[4];
    int64_t out_off = 0;
    uint8_t sw = (uint8_t)(Bool_val(swaped) ? 1 : 0);
    uint8_t *buf;
    size_t cap;
    poy_status s;
    Seq_custom_val(a, si); Seq_custom_val(b, sj);
    Seq_custom_val(ri, resi); Seq_custom_val(rj, resj); Seq_custom_val(md, median); Seq_custom_val(mw, medianwg);
    if (a->len > b->len) stub_fail("pass the shorter one as first");     /* src/algn.c:2396 */
    cap = (size_t)a->len + (size_t)b->len + 2;
    buf = (uint8_t *)malloc(4 * cap);
    if (!buf) stub_fail("Out of memory");
    pool = pool_of_two(a, b);
    s = poy_batch_align_affine(ctx, dcm, pool, 1, &i0, &i1, &sw, &out_off, &cost, buf, buf + cap, buf + 2 * cap, buf + 3 * cap, lens, NULL);
    poy_pool_free(ctx, pool);
    if (s != POY_OK) { free(buf); stub_fail(poy_last_error(ctx)); }
    /* each output is right-justified in its capacity-(len_i + len_j + 2) slot, i.e. already "prepended" */
    prepend_all(md, buf + cap - lens[0], lens[0]);
    prepend_all(mw, buf + 2 * cap - lens[1], lens[1]);
    prepend_all(ri, buf + 3 * cap - lens[2], lens[2]);
    prepend_all(rj, buf + 4 * cap - lens[3], lens[3]);
    free(buf);
    CAMLreturn(Val_int(cost));
}

value algn_CAML_align_affine_3_bc(value *argv, int argn) {
    (void)argn;
    return algn_CAML_align_affine_3(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7], argv[8]);
}
