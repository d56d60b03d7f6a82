/* TEST INFRASTRUCTURE ONLY: minimal bigarray descriptor. */
#ifndef POY_SHIM_BIGARRAY_H
#define POY_SHIM_BIGARRAY_H
#include "mlvalues.h"
#include "custom.h"
#include <stdarg.h>
#include <string.h>
struct caml_bigarray { void *data; intnat num_dims; intnat flags; void *proxy; intnat dim[4]; };
#define caml_ba_array caml_bigarray
#define Bigarray_val(v) ((struct caml_bigarray *)Data_custom_val(v))
#define Caml_ba_array_val(v) Bigarray_val(v)
#define Data_bigarray_val(v) (Bigarray_val(v)->data)
#define Caml_ba_data_val(v) Data_bigarray_val(v)
enum { BIGARRAY_FLOAT32 = 0, BIGARRAY_FLOAT64, BIGARRAY_SINT8, BIGARRAY_UINT8, BIGARRAY_SINT16,
       BIGARRAY_UINT16, BIGARRAY_INT32, BIGARRAY_INT64, BIGARRAY_CAML_INT, BIGARRAY_NATIVE_INT,
       BIGARRAY_COMPLEX32, BIGARRAY_COMPLEX64 };
#define BIGARRAY_C_LAYOUT 0
#define BIGARRAY_FORTRAN_LAYOUT 0x100
#define CAML_BA_FLOAT32 BIGARRAY_FLOAT32
#define CAML_BA_FLOAT64 BIGARRAY_FLOAT64
#define CAML_BA_INT32 BIGARRAY_INT32
#define CAML_BA_C_LAYOUT BIGARRAY_C_LAYOUT
static inline value poy_shim_alloc_bigarray(int flags, int num_dims, void *data, intnat *dim) {
    value v = poy_shim_alloc_custom(NULL, sizeof(struct caml_bigarray));
    struct caml_bigarray *b = Bigarray_val(v);
    int i; b->data = data; b->num_dims = num_dims; b->flags = flags;
    for (i = 0; i < num_dims && i < 4; i++) b->dim[i] = dim[i];
    return v;
}
static inline value poy_shim_alloc_bigarray_dims(int flags, int num_dims, void *data, ...) {
    intnat dim[4] = {0,0,0,0}; va_list ap; int i;
    va_start(ap, data);
    for (i = 0; i < num_dims && i < 4; i++) dim[i] = va_arg(ap, intnat);
    va_end(ap);
    return poy_shim_alloc_bigarray(flags, num_dims, data, dim);
}
#define alloc_bigarray poy_shim_alloc_bigarray
#define caml_ba_alloc poy_shim_alloc_bigarray
#define alloc_bigarray_dims poy_shim_alloc_bigarray_dims
#define caml_ba_alloc_dims poy_shim_alloc_bigarray_dims
#endif
