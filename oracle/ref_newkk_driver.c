/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path
 * (poy5_b200/).  Second translation unit of oracle/_ref/libpoyref*.so (see ref_driver.c for the
 * rules): drives the UNMODIFIED reference src/newkkonen.c (+ queue_with_linkedlist.c) and
 * src/ukkCommon.c / src/ukk.checkp.c through their own `*_CAML_*` entry points, exactly as the
 * OCaml `external` declarations of src/sequence.ml:1834-1847 (Sequence.NewkkAlign) and
 * src/sequence.ml:1309-1311 (powell_3D_align) would.  No reference source is copied here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <assert.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include "seq.h"
#include "cm.h"

value seq_CAML_create(value cap);
value seq_CAML_prepend(value s, value v);
value newkkonen_CAML_create_general(value a);
value newkkonen_CAML_algn(value s1, value s2, value c, value a, value swaped);
value newkkonen_CAML_algn_affine(value s1, value s2, value c, value a, value swaped);
value newkkonen_CAML_backtrace(value s1, value s2, value s1p, value s2p, value c, value a, value swaped);
value newkkonen_CAML_backtrace_affine(value s1, value s2, value s1p, value s2p, value c, value a, value swaped);
value newkkonen_CAML_get_k(value m);
void newkkmat_CAML_free(value m);
value powell_3D_align(value sa, value sb, value sc, value ra, value rb, value rc, value mm, value go, value ge);

/* the jump buffer of ref_driver.c's caml_failwith (same shared object) */
extern __thread jmp_buf *ref_jmp_shared;
#define REF_FAIL (-2147483647 - 1)
#define GUARD_BEGIN jmp_buf jb; jmp_buf *saved = ref_jmp_shared; ref_jmp_shared = &jb; if (setjmp(jb) == 0) {
#define GUARD_END(failval) ref_jmp_shared = saved; } else { ref_jmp_shared = saved; return (failval); }

static value mk_seq(const unsigned char *s, int len, int cap) {
    value v = seq_CAML_create(Val_int(cap));
    int i;
    for (i = len - 1; i >= 0; i--) seq_CAML_prepend(v, Val_int(s[i]));
    return v;
}
static int rd_seq(value v, unsigned char *out) {
    seqt s;
    Seq_custom_val(s, v);
    if (out) memcpy(out, s->begin, (size_t)s->len);
    return s->len;
}

/* ---- Sequence.NewkkAlign (src/sequence.ml:1831-1925) ---------------------- */
void *ref_newkk_new(void) { return (void *)newkkonen_CAML_create_general(Val_int(0)); }
void ref_newkk_free(void *m) { newkkmat_CAML_free((value)m); free(m); }
int ref_newkk_get_k(void *m) { return Int_val(newkkonen_CAML_get_k((value)m)); }

/* cost only: newkk_cost2 / newkk_cost2_affine; s1 must be the shorter sequence */
int ref_newkk_cost(void *cm, void *m, const unsigned char *s1, int len1, const unsigned char *s2, int len2,
                   int affine, int swaped) {
    value a = mk_seq(s1, len1, len1), b = mk_seq(s2, len2, len2);
    int res;
    GUARD_BEGIN
    res = affine ? Int_val(newkkonen_CAML_algn_affine(a, b, (value)cm, (value)m, Val_int(swaped)))
                 : Int_val(newkkonen_CAML_algn(a, b, (value)cm, (value)m, Val_int(swaped)));
    GUARD_END((free((void *)a), free((void *)b), REF_FAIL))
    free((void *)a); free((void *)b);
    return res;
}

/* cost + traceback: NewkkAlign.align_2 without the operand exchange (the caller passes the shorter one first
 * and `swaped`); r1 / r2 must hold len1 + len2 bytes (get_alignment allocates sz1 + sz2); lens[2] = their lengths */
int ref_newkk_align(void *cm, void *m, const unsigned char *s1, int len1, const unsigned char *s2, int len2,
                    int affine, int swaped, unsigned char *r1, unsigned char *r2, int *lens) {
    int cap = len1 + len2 + 2, res;
    value a = mk_seq(s1, len1, len1), b = mk_seq(s2, len2, len2);
    value v1 = seq_CAML_create(Val_int(cap)), v2 = seq_CAML_create(Val_int(cap));
    GUARD_BEGIN
    if (affine) {
        res = Int_val(newkkonen_CAML_algn_affine(a, b, (value)cm, (value)m, Val_int(swaped)));
        newkkonen_CAML_backtrace_affine(a, b, v1, v2, (value)cm, (value)m, Val_int(swaped));
    } else {
        res = Int_val(newkkonen_CAML_algn(a, b, (value)cm, (value)m, Val_int(swaped)));
        newkkonen_CAML_backtrace(a, b, v1, v2, (value)cm, (value)m, Val_int(swaped));
    }
    GUARD_END((free((void *)a), free((void *)b), free((void *)v1), free((void *)v2), REF_FAIL))
    lens[0] = rd_seq(v1, r1);
    lens[1] = rd_seq(v2, r2);
    free((void *)a); free((void *)b); free((void *)v1); free((void *)v2);
    return res;
}

/* ---- powell_3D_align (src/ukkCommon.c:109-145; Sequence.Align.align_3_powell, src/sequence.ml:1309-1340) ----
 * The three inputs are complete sequences: copySequence (src/ukkCommon.c:86-107) skips element 0, the leading gap;
 * r1..r3 must hold len1+len2+len3 bytes each; lens[3] = lengths of the three aligned rows. */
int ref_powell_3d(const unsigned char *s1, int len1, const unsigned char *s2, int len2, const unsigned char *s3, int len3,
                  int mm, int go, int ge, unsigned char *r1, unsigned char *r2, unsigned char *r3, int *lens) {
    int cap = len1 + len2 + len3 + 3, res;
    value a = mk_seq(s1, len1, len1 + 1), b = mk_seq(s2, len2, len2 + 1), c = mk_seq(s3, len3, len3 + 1);
    value v1 = seq_CAML_create(Val_int(cap)), v2 = seq_CAML_create(Val_int(cap)), v3 = seq_CAML_create(Val_int(cap));
    GUARD_BEGIN
    res = Int_val(powell_3D_align(a, b, c, v1, v2, v3, Val_int(mm), Val_int(go), Val_int(ge)));
    GUARD_END((free((void *)a), free((void *)b), free((void *)c), free((void *)v1), free((void *)v2), free((void *)v3), REF_FAIL))
    lens[0] = rd_seq(v1, r1);
    lens[1] = rd_seq(v2, r2);
    lens[2] = rd_seq(v3, r3);
    free((void *)a); free((void *)b); free((void *)c); free((void *)v1); free((void *)v2); free((void *)v3);
    return res;
}
