/* TEST INFRASTRUCTURE ONLY: (de)serialisation hooks are never exercised. */
#ifndef POY_SHIM_INTEXT_H
#define POY_SHIM_INTEXT_H
#include "mlvalues.h"
static inline void caml_serialize_int_1(int i) { (void)i; }
static inline void caml_serialize_int_2(int i) { (void)i; }
static inline void caml_serialize_int_4(int32_t i) { (void)i; }
static inline void caml_serialize_int_8(int64_t i) { (void)i; }
static inline void caml_serialize_block_1(void *d, long n) { (void)d; (void)n; }
static inline void caml_serialize_block_2(void *d, long n) { (void)d; (void)n; }
static inline void caml_serialize_block_4(void *d, long n) { (void)d; (void)n; }
static inline void caml_serialize_block_8(void *d, long n) { (void)d; (void)n; }
static inline int caml_deserialize_uint_1(void) { return 0; }
static inline int caml_deserialize_sint_1(void) { return 0; }
static inline int caml_deserialize_uint_2(void) { return 0; }
static inline int caml_deserialize_sint_2(void) { return 0; }
static inline uint32_t caml_deserialize_uint_4(void) { return 0; }
static inline int32_t caml_deserialize_sint_4(void) { return 0; }
static inline int64_t caml_deserialize_sint_8(void) { return 0; }
static inline void caml_deserialize_block_1(void *d, long n) { (void)d; (void)n; }
static inline void caml_deserialize_block_2(void *d, long n) { (void)d; (void)n; }
static inline void caml_deserialize_block_4(void *d, long n) { (void)d; (void)n; }
static inline void caml_deserialize_block_8(void *d, long n) { (void)d; (void)n; }
#define serialize_int_1 caml_serialize_int_1
#define serialize_int_2 caml_serialize_int_2
#define serialize_int_4 caml_serialize_int_4
#define serialize_int_8 caml_serialize_int_8
#define serialize_block_1 caml_serialize_block_1
#define serialize_block_2 caml_serialize_block_2
#define serialize_block_4 caml_serialize_block_4
#define serialize_block_8 caml_serialize_block_8
#define deserialize_uint_1 caml_deserialize_uint_1
#define deserialize_sint_1 caml_deserialize_sint_1
#define deserialize_uint_2 caml_deserialize_uint_2
#define deserialize_sint_2 caml_deserialize_sint_2
#define deserialize_uint_4 caml_deserialize_uint_4
#define deserialize_sint_4 caml_deserialize_sint_4
#define deserialize_sint_8 caml_deserialize_sint_8
#define deserialize_block_1 caml_deserialize_block_1
#define deserialize_block_2 caml_deserialize_block_2
#define deserialize_block_4 caml_deserialize_block_4
#define deserialize_block_8 caml_deserialize_block_8
#endif
