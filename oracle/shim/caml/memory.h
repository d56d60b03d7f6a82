/* TEST INFRASTRUCTURE ONLY: GC-root macros become no-ops (no OCaml GC here). */
#ifndef POY_SHIM_MEMORY_H
#define POY_SHIM_MEMORY_H
#include "mlvalues.h"
#include <stdlib.h>
#define CAMLparam0()
#define CAMLparam1(a)
#define CAMLparam2(a,b)
#define CAMLparam3(a,b,c)
#define CAMLparam4(a,b,c,d)
#define CAMLparam5(a,b,c,d,e)
#define CAMLxparam1(a)
#define CAMLxparam2(a,b)
#define CAMLxparam3(a,b,c)
#define CAMLxparam4(a,b,c,d)
#define CAMLxparam5(a,b,c,d,e)
#define CAMLlocal1(a) value a = 0
#define CAMLlocal2(a,b) value a = 0, b = 0
#define CAMLlocal3(a,b,c) value a = 0, b = 0, c = 0
#define CAMLlocal4(a,b,c,d) value a = 0, b = 0, c = 0, d = 0
#define CAMLlocal5(a,b,c,d,e) value a = 0, b = 0, c = 0, d = 0, e = 0
#define CAMLreturn(x) return (x)
#define CAMLreturn0 return
#define CAMLreturnT(t, x) return (x)
#define caml_stat_alloc malloc
#define caml_stat_free free
#define stat_alloc malloc
#define stat_free free
#define caml_modify(p, v) (*(p) = (v))
#define caml_initialize(p, v) (*(p) = (v))
#define register_global_root(x)
#define caml_register_global_root(x)
#define caml_remove_global_root(x)
#endif
