/* TEST INFRASTRUCTURE ONLY: block allocation as plain calloc with a size header. */
#ifndef POY_SHIM_ALLOC_H
#define POY_SHIM_ALLOC_H
#include "mlvalues.h"
#include <stdlib.h>
#include <string.h>
static inline value poy_shim_alloc(mlsize_t n, int tag) {
    value *b = (value *)calloc(n + 1, sizeof(value)); (void)tag;
    if (!b) abort();
    b[0] = (value)n; return (value)(b + 1);
}
#define caml_alloc(n, tag) poy_shim_alloc((n), (tag))
#define caml_alloc_tuple(n) poy_shim_alloc((n), 0)
#define caml_alloc_small(n, tag) poy_shim_alloc((n), (tag))
#define alloc_tuple caml_alloc_tuple
#define alloc caml_alloc
static inline value caml_copy_double(double d) { value v = poy_shim_alloc(1, 253); memcpy((void*)v, &d, sizeof d); return v; }
static inline value caml_copy_string(const char *s) { size_t n = strlen(s); value v = poy_shim_alloc(n / sizeof(value) + 1, 252); memcpy((void*)v, s, n + 1); return v; }
#define copy_double caml_copy_double
#define copy_string caml_copy_string
#endif
