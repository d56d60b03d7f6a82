/* TEST INFRASTRUCTURE ONLY: custom blocks are plain calloc'd memory with a
 * one-word header holding the operations pointer. */
#ifndef POY_SHIM_CUSTOM_H
#define POY_SHIM_CUSTOM_H
#include "mlvalues.h"
#include <stdlib.h>
struct custom_operations {
    char *identifier;
    void (*finalize)(value v);
    int (*compare)(value v1, value v2);
    long (*hash)(value v);
    void (*serialize)(value v, unsigned long *wsize_32, unsigned long *wsize_64);
    unsigned long (*deserialize)(void *dst);
};
#define custom_finalize_default NULL
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define Data_custom_val(v) ((void *)(((value *)(v)) + 1))
#define Custom_ops_val(v) (*((struct custom_operations **)(v)))
static inline value poy_shim_alloc_custom(struct custom_operations *ops, unsigned long size) {
    value *b = (value *)calloc(1, sizeof(value) + size);
    if (!b) abort();
    b[0] = (value)ops;
    return (value)b;
}
#define caml_alloc_custom(ops, size, mem, max) poy_shim_alloc_custom((ops), (size))
#define alloc_custom(ops, size, mem, max) poy_shim_alloc_custom((ops), (size))
#define caml_register_custom_operations(ops) ((void)(ops))
#define register_custom_operations(ops) ((void)(ops))
#endif
