/* TEST INFRASTRUCTURE ONLY: failwith longjmps back to the driver if a
 * handler is armed, else aborts. */
#ifndef POY_SHIM_FAIL_H
#define POY_SHIM_FAIL_H
#include "mlvalues.h"
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
void caml_raise_out_of_memory(void) __attribute__((noreturn));
#define failwith caml_failwith
#define invalid_argument caml_invalid_argument
#define raise_out_of_memory caml_raise_out_of_memory
#endif
