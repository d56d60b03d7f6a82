/* TEST INFRASTRUCTURE ONLY. Minimal stand-in for the OCaml runtime header
 * <caml/mlvalues.h>, sufficient to compile the unmodified reference C
 * (src/algn.c and friends) WITHOUT an OCaml toolchain.
 * OCaml ints are tagged: Val_int(x) = 2x+1. */
#ifndef POY_SHIM_MLVALUES_H
#define POY_SHIM_MLVALUES_H
#include <stdint.h>
#include <stddef.h>
typedef intptr_t value;
typedef uintptr_t uvalue;
typedef uintptr_t mlsize_t;
typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef int32_t int32;
typedef uint32_t uint32;
#define Val_long(x) ((value)(((uintptr_t)(intptr_t)(x) << 1) + 1))
#define Long_val(x) ((intptr_t)(x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int)Long_val(x))
#define Unsigned_long_val(x) ((uintptr_t)(x) >> 1)
#define Val_unit Val_int(0)
#define Val_bool(x) Val_int((x) != 0)
#define Bool_val(x) Int_val(x)
#define Val_true Val_int(1)
#define Val_false Val_int(0)
#define Is_long(x) (((x) & 1) != 0)
#define Is_block(x) (((x) & 1) == 0)
#define Field(x, i) (((value *)(x))[i])
#define Store_field(b, i, v) (Field(b, i) = (v))
#define Wosize_val(v) ((mlsize_t)(((value *)(v))[-1]))
#define Double_val(v) (*(double *)(v))
#define String_val(v) ((char *)(v))
#define Bp_val(v) ((char *)(v))
#define Double_field(v, i) (((double *)(v))[i])
#define Store_double_field(v, i, d) (((double *)(v))[i] = (d))
#endif
