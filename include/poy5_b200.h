/* poy5_b200 -- B200-native batched direct-optimization (DO) alignment for POY5.
 *
 * C ABI of libpoy5b200.so.  This is the drop-in boundary for the hot path of
 * amnh/poy5's libpoycside C stubs (src/libpoycside.clib): every entry point is
 * the BATCH twin of one reference stub and cites the stub it replaces.  Plain
 * pointers and sizes only; no CUDA, torch or OCaml types appear here.
 *
 * Data conventions (identical to the reference):
 *   - a sequence is an array of uint8 DNA bitset codes A=1 C=2 G=4 T=8 gap=16
 *     (ambiguity = OR); element 0 is always the gap code and `len` counts it
 *     (struct seq, src/seq.h:52-61);
 *   - sequences live in a POOL: one packed byte buffer plus nseq+1 offsets;
 *     pairs name pool entries by index, so a node sequence used by thousands
 *     of candidate pairs is uploaded once (the Parmap candidate list of
 *     src/ptree.ml:1356-1408 is exactly such a batch);
 *   - costs are int32; there is no saturation (HIGH_NUM = 1000000 sentinel
 *     arithmetic of src/algn.c:37 is reproduced bit for bit).
 *
 * Errors: every call returns a poy_status (0 = OK).  The reference raises
 * OCaml `Failure` (failwith); here the same conditions come back as codes and
 * poy_last_error() holds the reference's message text.  There is NO CPU
 * fallback: without a usable CUDA device the context cannot be created.
 */
#ifndef POY5_B200_H
#define POY5_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define POY_API __attribute__((visibility("default")))
#else
#define POY_API
#endif

typedef int poy_status;
enum {
    POY_OK = 0,
    POY_ERR_CUDA = -1,          /* CUDA runtime failure (message in poy_last_error) */
    POY_ERR_NO_DEVICE = -2,     /* no CUDA device: the product path refuses to run */
    POY_ERR_ARG = -3,           /* bad argument (null pointer, index out of range, ...) */
    POY_ERR_ORDER = -4,         /* "pass the shorter one as first" (src/algn.c:2396) */
    POY_ERR_COST_RANGE = -5,    /* cost model outside the domain in which the reference is defined:
                                   negative entries, or path costs that can reach HIGH_NUM */
    POY_ERR_MODEL = -6,         /* entry point does not match cm.cost_model_type */
    POY_ERR_NOMEM = -7
};

typedef struct poy_ctx poy_ctx;     /* one per process per GPU: stream, scratch pools */
typedef struct poy_cm poy_cm;       /* device-resident 2-D cost model (struct cm, src/cm.h:33-76) */
typedef struct poy_cm3d poy_cm3d;   /* device-resident 3-D cost model (struct cm_3d, src/cm.h:253-280) */
typedef struct poy_pool poy_pool;   /* device-resident packed sequences + per-base gap parameters */

/* ---- context ------------------------------------------------------------- */
/* `stream` is a cudaStream_t passed as void* (0 = the library creates its own
 * non-blocking stream).  All device work of this context is ordered on it. */
POY_API poy_status poy_ctx_create(int device, void *stream, poy_ctx **out);
POY_API void poy_ctx_destroy(poy_ctx *ctx);
POY_API const char *poy_last_error(const poy_ctx *ctx);
POY_API const char *poy_status_string(poy_status s);
/* cap (bytes) on the direction-matrix arena used by the traceback entry points;
 * batches larger than the cap are processed in waves.  Default 8 GiB. */
POY_API poy_status poy_ctx_set_arena_limit(poy_ctx *ctx, uint64_t bytes);
POY_API poy_status poy_ctx_synchronize(poy_ctx *ctx);
/* returns the context's grow-only device scratch (direction arenas, job arrays, cached pool blocks) to the driver; it is
 * re-allocated on demand.  For callers that switch between workloads of very different shape. */
POY_API poy_status poy_ctx_trim(poy_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py: gpu_launches) */
POY_API uint64_t poy_ctx_launch_count(const poy_ctx *ctx);
/* cumulative counters of this context (SURVEY.md 8d "additionally report cells_computed"): out[0] kernel launches,
 * out[1] band cells computed by the banded entry points (summed over every fill of the threshold-doubling schedule),
 * out[2] probe fills (no direction bytes), out[3] fills with direction bytes, out[4] thresholds repeated after a probe,
 * out[5] schedule rounds, out[6] pairs aligned by the banded entry points, out[7] reserved */
POY_API poy_status poy_ctx_stats(const poy_ctx *ctx, int64_t out[8]);

/* ---- cost model: replaces the cm_CAML_* setters + cm.c tables --------------
 * Host-side image of `struct cm` for the 5-letter bitset alphabet
 * (combinations = 1, level = 0, lcm = 5, gap = 16, all_elements = 31).  Tables
 * are indexed (a << 5) + b like cm_calc_cost_position (src/cm.c:903-911). */
typedef struct {
    int32_t cost[1024];      /* c->cost   */
    int32_t worst[1024];     /* c->worst  */
    uint8_t median[1024];    /* c->median */
    int32_t prepend[32];     /* c->prepend_cost (src/cost_matrix.ml:994-1001) */
    int32_t tail[32];        /* c->tail_cost */
    int32_t gap_open;        /* c->gap_open */
    int32_t cost_model_type; /* 0 linear, 1 affine, 2 no alignment (c->cost_model_type) */
    int32_t is_identity;     /* c->is_identity: the input matrix has a zero diagonal (src/cost_matrix.ml:1097-1104, 1169-1170) */
    int32_t is_metric;       /* c->is_metric: positive, symmetric and zero diagonal (src/cost_matrix.ml:1117-1139, 1171-1177) */
} poy_cm_host;

/* Cost_matrix.Two_D construction restated in C++ (src/cost_matrix.ml:721-804,
 * 862-897, 994-1016, 1140-1189): fills `full` (c2_full) and `original`
 * (c2_original) from a 5x5 single-letter matrix in row-major order
 * (A,C,G,T,gap).  gap_open < 0 means "linear" (no set_cost_model call);
 * otherwise the affine re-fill of src/data.ml:5937-5964 is applied. */
POY_API poy_status poy_cm_fill(const int32_t single[25], int32_t gap_open, poy_cm_host *full, poy_cm_host *original);
/* cm_get_min_non0_cost (src/cm.c:1063-1089) */
POY_API int32_t poy_cm_min_non0(const poy_cm_host *cm);
/* Cost_matrix.Two_D.get_closest (src/cost_matrix.ml:1387-1428) */
POY_API int32_t poy_cm_get_closest(const poy_cm_host *cm, int32_t a, int32_t b);

POY_API poy_status poy_cm_upload(poy_ctx *ctx, const poy_cm_host *cm, poy_cm **out);
POY_API void poy_cm_free(poy_ctx *ctx, poy_cm *cm);

/* ---- 3-D cost model: struct cm_3d (src/cm.h:253-280) for the bit-indexed 5-letter alphabet -----------------------
 * Tables indexed (((a << 5) + b) << 5) + c like cm_calc_cost_position_3d (src/cm.c:945-953).
 *  poy_cm3d_fill    Cost_matrix.Three_D.of_two_dim_comb (src/cost_matrix.ml:1605-1652) restated in C++: cost = min
 *                   over the five single elements of the three 2-D costs (the gap only if the 2-D matrix is metric or
 *                   at least two inputs contain it), median = the lowest minimising bit (pick_bit)
 *  poy_cm3d_upload  cm_set_val_3d + the cm_CAML_set_*_3d setters (src/cm.c:704): the tables become device resident
 *  poy_batch_median_3  the column-wise three-way median of Sequence.Align.align_3_powell_inter
 *                   (src/sequence.ml:1342-1369): for three aligned rows of equal length, medianwg[x] =
 *                   Three_D.median a[x] b[x] c[x]; median = gap :: the non-gap ones; cost3 = sum of Three_D.cost.
 *                   Rows are packed HOST buffers (triple p = rows_x[off[p] .. off[p] + len[p])); the outputs of
 *                   triple p start at out_off[p] (capacity len[p] + 1), out_len[p] = length of median. */
typedef struct {
    int32_t cost[32768];
    uint8_t median[32768];
} poy_cm3d_host;
POY_API poy_status poy_cm3d_fill(const poy_cm_host *c2, poy_cm3d_host *out);
POY_API poy_status poy_cm3d_upload(poy_ctx *ctx, const poy_cm3d_host *cm3, poy_cm3d **out);
POY_API void poy_cm3d_free(poy_ctx *ctx, poy_cm3d *cm3);
POY_API poy_status poy_batch_median_3(poy_ctx *ctx, const poy_cm3d *cm3, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                      const uint8_t *rows_c, const int64_t *off, const int32_t *len, const int64_t *out_off,
                                      uint8_t *median, uint8_t *medianwg, int32_t *out_len, int32_t *cost3);

/* ---- sequence pool: replaces seq_CAML_create/prepend for batch inputs ------
 * `data` holds nseq sequences back to back; sequence s is
 * data[offsets[s] .. offsets[s+1]).  The *_dev variant takes DEVICE pointers
 * (already resident in HBM); the plain variant takes HOST pointers and copies. */
POY_API poy_status poy_pool_upload(poy_ctx *ctx, const uint8_t *data, const int64_t *offsets, int32_t nseq, poy_pool **out);
POY_API poy_status poy_pool_from_device(poy_ctx *ctx, const uint8_t *d_data, const int64_t *d_offsets,
                                const int64_t *h_offsets, int32_t nseq, poy_pool **out);
POY_API void poy_pool_free(poy_ctx *ctx, poy_pool *pool);

/* ---- batch twin of algn_CAML_cost_affine_3 (src/algn.c:2457-2515) ----------
 * cost[p] = Sequence.Align.cost_2 (affine) of pool[a[p]] vs pool[b[p]]:
 * full-matrix 4-state affine DO cost, either argument order, including the
 * reference's even-row/last-column EV behaviour.
 * a, b, cost are HOST arrays of n entries (the *_dev variant: DEVICE arrays). */
POY_API poy_status poy_batch_cost_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                 const int32_t *a, const int32_t *b, int32_t *cost);
POY_API poy_status poy_batch_cost_affine_dev(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                     const int32_t *d_a, const int32_t *d_b, int32_t *d_cost);

/* ---- batch twin of algn_CAML_align_affine_3 (src/algn.c:2359-2447) ---------
 * Ukkonen-banded affine DO alignment with the reference's threshold-doubling
 * schedule, on-device traceback (backtrace_aff, src/algn.c:1715-1819) and
 * median construction.  Requires len(pool[si[p]]) <= len(pool[sj[p]])
 * (POY_ERR_ORDER otherwise); swaped[p] is the flag the OCaml caller passes
 * (src/sequence.ml:636) and selects the traceback tie-break.
 *
 * Outputs, per pair p, exactly like the four result `struct seq`s the caller
 * allocates with capacity cap_p = len_i + len_j + 2 and the stub fills by
 * prepending: each of median / medianwg / resi / resj is a packed byte buffer;
 * pair p owns the slot [out_off[p], out_off[p] + cap_p) and its sequence is
 * RIGHT-justified in that slot (begin = slot_end - len), out_len[4*p + {0,1,2,3}]
 * = lengths of median, medianwg, resi, resj.  Any output pointer may be NULL.
 * stats (optional, n x 4 int32): iterations, final T, final k, band cells / 1024. */
POY_API poy_status poy_batch_align_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                  const int32_t *si, const int32_t *sj, const uint8_t *swaped,
                                  const int64_t *out_off, int32_t *cost, uint8_t *median,
                                  uint8_t *medianwg, uint8_t *resi, uint8_t *resj, int32_t *out_len,
                                  int32_t *stats);
/* same, all array arguments are DEVICE pointers except out_off_host (needed to size waves) */
POY_API poy_status poy_batch_align_affine_dev(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                      const int32_t *d_si, const int32_t *d_sj, const uint8_t *d_swaped,
                                      const int32_t *h_si, const int32_t *h_sj,
                                      const int64_t *d_out_off, int32_t *d_cost, uint8_t *d_median,
                                      uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj,
                                      int32_t *d_out_len, int32_t *d_stats);

/* ---- batch twins of algn_CAML_simple_2 (src/algn.c:3134) and algn_CAML_align_2d (:3500) -----
 * Linear-gap DO alignment (cost_model_type 0 or 2): algn_nw -> algn_fill_plane_2 chooses the full
 * plane or the Ukkonen band from the lengths and deltawh[p] exactly like the reference
 * (src/algn.c:1134-1177); deltawh is what the OCaml caller computes in Sequence.Align.cost_2
 * (src/sequence.ml:868-925).  s1[p] must be the shorter sequence (POY_ERR_ORDER otherwise).
 * align: r1 / r2 receive the two aligned rows of backtrace_2d (src/algn.c:3277-3327), pair p
 * right-justified in the slot [out_off[p], out_off[p] + len1 + len2) (create_edited_2 allocates
 * len1+len2, src/sequence.ml:1019-1033); out_len[2*p + {0,1}] = their lengths; swaped[p] selects
 * the insertion/deletion preference.  stats as for the affine entry point. */
POY_API poy_status poy_batch_cost_linear(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                         const int32_t *s1, const int32_t *s2, const int32_t *deltawh, int32_t *cost);
POY_API poy_status poy_batch_align_linear(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n,
                                          const int32_t *s1, const int32_t *s2, const int32_t *deltawh,
                                          const uint8_t *swaped, const int64_t *out_off, int32_t *cost, uint8_t *r1,
                                          uint8_t *r2, int32_t *out_len, int32_t *stats);

/* ---- batch twin of Sequence.NewkkAlign (src/sequence.ml:1831-2062): newkkonen_CAML_algn_affine +
 * newkkonen_CAML_backtrace_affine (src/newkkonen.c:1472-1495, 1787-1803) ------------------------------------------
 * The diagonal-storage Ukkonen alignment with the gap-count stop rule (newkk_algn / increaseT / ukktest /
 * update_internal_cell, src/newkkonen.c:1360-1444, 1155-1171, 1077-1153, 680-870), affine model only: the reference's
 * non-affine entry point never sets costDiag (:796, :855-856), returns cost 0 and its traceback raises -> POY_ERR_MODEL.
 * s1[p] must be the shorter sequence (POY_ERR_ORDER = "newkkonen.newkk_algn, s1 len > s2 len"); swaped[p] is what
 * NewkkAlign.align_2 passes (1 if it exchanged the operands).  r1 / r2 receive the two aligned rows of
 * backtrace_affine, pair p RIGHT-justified in the slot [out_off[p], out_off[p] + len1 + len2) (get_alignment
 * allocates sz1 + sz2, src/sequence.ml:1862-1877), out_len[2*p + {0,1}] = their lengths.  r1 == NULL: cost only
 * (NewkkAlign.cost_2).  stats (optional, n x 4): threshold doublings, 0, final k (newkkonen_CAML_get_k), 1 if the
 * pair took the trivial path (len1 * 100 < len2, trivial_algn :1351-1356).  All arrays are HOST pointers. */
POY_API poy_status poy_batch_newkk_align(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n, const int32_t *s1,
                                         const int32_t *s2, const uint8_t *swaped, const int64_t *out_off, int32_t *cost,
                                         uint8_t *r1, uint8_t *r2, int32_t *out_len, int32_t *stats);

/* ---- column-wise helpers over ALIGNED rows (O(L) per pair) ---------------------------------
 * rows_a / rows_b are packed byte buffers; pair p uses rows_x[off[p] .. off[p] + len[p]).
 *  poy_batch_median_2     seq_CAML_median_2_with_gaps / _no_gaps (src/seq.c:241-296): out slot of pair p
 *                         starts at out_off[p] (capacity len[p] + 1), left-justified, out_len[p] bytes
 *  poy_batch_union        algn_CAML_union (src/algn.c:3657-3678): out has the layout of the inputs
 *  poy_batch_aligned_cost algn_CAML_verify_2 (use_worst = 0) / algn_CAML_worst_2 (use_worst = 1)
 *                         (src/algn.c:3003-3130) = Sequence.Align.max_cost_2 / verify;
 *                         use_worst = 2 / 3: Sequence.Align.recost ~first_gap:true / false (src/sequence.ml:1244-1307,
 *                         OCaml in the reference), the cost of two aligned rows with one gap opening per gap block
 *  poy_batch_ancestor_2   algn_CAML_ancestor_2 (src/algn.c:3603-3626,3742): capacity len[p] + 1
 *  poy_batch_closest      the column map of Sequence.Align.closest (src/sequence.ml:1180-1237): column i becomes
 *                         Cost_matrix.Two_D.get_closest cm parent.(i) mine.(i) (src/cost_matrix.ml:1387-1428), then
 *                         remove_gaps2 (src/sequence.ml:209-222); capacity len[p] + 1 */
POY_API poy_status poy_batch_median_2(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                      const int64_t *off, const int32_t *len, int32_t with_gaps, const int64_t *out_off,
                                      uint8_t *out, int32_t *out_len);
POY_API poy_status poy_batch_union(poy_ctx *ctx, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b, const int64_t *off,
                                   const int32_t *len, uint8_t *out);
POY_API poy_status poy_batch_aligned_cost(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                          const int64_t *off, const int32_t *len, int32_t use_worst, int32_t *cost);
POY_API poy_status poy_batch_ancestor_2(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_a, const uint8_t *rows_b,
                                        const int64_t *off, const int32_t *len, const int64_t *out_off, uint8_t *out,
                                        int32_t *out_len);
POY_API poy_status poy_batch_closest(poy_ctx *ctx, const poy_cm *cm, int32_t n, const uint8_t *rows_parent, const uint8_t *rows_mine,
                                     const int64_t *off, const int32_t *len, const int64_t *out_off, uint8_t *out, int32_t *out_len);

/* ---- batch twins of SeqCS.DOS.distance / SeqCS.DOS.median (src/seqCS.ml:701-774, 985-1084) ----------------
 * The per-locus policy the OCaml side applies around the alignment stubs, restated once above the batch entry
 * points so that the candidate seam (src/ptree.ml:1356-1408) can hand over (a[p], b[p]) in ANY order:
 *  poy_dos_distance  cost[p] = DOS.distance: missing_distance if either sequence is empty (Sequence.is_empty,
 *                    src/seqCS.ml:705-709); else Sequence.Align.cost_2 under c2_ORIGINAL -- affine:
 *                    algn_CAML_cost_affine_3; linear: shorter first, deltaw = max(|len a - len b|, 8) folded into
 *                    deltawh = count_gaps + deltaw_calc (src/sequence.ml:868-925), algn_CAML_simple_2
 *  poy_dos_median    DOS.median: an empty child yields the other child with cost 0 (identity matrices; see poy_dos_median2)
 *                    (src/seqCS.ml:991-1039); else Sequence.Align.align_affine_3 under c2_FULL with the shorter
 *                    sequence first and swaped = len a > len b (src/sequence.ml:633-649).  Pair p owns the slot
 *                    [out_off[p], out_off[p] + len_a + len_b + 2) of `median`; its median sequence is
 *                    RIGHT-justified there, out_len[p] bytes long; cost2[p] is the alignment cost.
 *                    Linear / no-alignment models: Sequence.Align.align_2 + ancestor_2 (src/seqCS.ml:1058-1071).
 * All array arguments are HOST pointers. */
POY_API poy_status poy_dos_distance(poy_ctx *ctx, const poy_cm *c2_original, const poy_pool *pool, int32_t n,
                                    const int32_t *a, const int32_t *b, int32_t missing_distance, int32_t *cost);
POY_API poy_status poy_dos_median(poy_ctx *ctx, const poy_cm *c2_full, const poy_pool *pool, int32_t n, const int32_t *a,
                                  const int32_t *b, const int64_t *out_off, int32_t *cost2, uint8_t *median,
                                  int32_t *out_len);

/* same, with the identity rule of DOS.median made explicit: when c2_original (may be NULL: c2_full's flag is used) has a
 * non-zero diagonal, a median with one empty child costs Sequence.Align.recost x x c2_original instead of 0
 * (src/seqCS.ml:992-996, 1027-1031).  Both entry points accept affine AND linear / no-alignment models (linear:
 * Sequence.Align.align_2 + ancestor_2, src/seqCS.ml:1058-1071). */
POY_API poy_status poy_dos_median2(poy_ctx *ctx, const poy_cm *c2_full, const poy_cm *c2_original, const poy_pool *pool, int32_t n,
                                   const int32_t *a, const int32_t *b, const int64_t *out_off, int32_t *cost2, uint8_t *median,
                                   int32_t *out_len);

/* ---- device-resident node store -------------------------------------------------------------------
 * The heap of `Sequence.s` values of a tree pass (src/seq.h:52-61; SeqCS.DOS.median allocates a fresh one per call,
 * src/seqCS.ml:985-1084) kept in HBM: sequences are immutable and named by an int32 id in order of creation.
 *  poy_store_append    copies host sequences in (observed leaves); *first_id = id of the first one
 *  poy_store_median    DOS.median of (a[p], b[p]) for n pairs of ids; the medians are appended to the store WITHOUT
 *                      leaving the device -- only out_id / out_len / cost2 (12 bytes per pair) come back.  A pair with an
 *                      empty child returns the other child's id (no copy).
 *  poy_store_distance  DOS.distance over ids (cost-only, c2_original)
 *  poy_store_truncate  stack discipline: forget every id >= nseq (temporaries of a finished batch of candidates)
 *  poy_store_read      copies sequences back to the host (reports, tests)
 *  poy_store_pool      the store as a poy_pool, for every poy_batch_* entry point */
typedef struct poy_store poy_store;
POY_API poy_status poy_store_create(poy_ctx *ctx, int64_t cap_bytes, int32_t cap_seqs, poy_store **out);
POY_API void poy_store_free(poy_ctx *ctx, poy_store *st);
POY_API const poy_pool *poy_store_pool(const poy_store *st);
POY_API int32_t poy_store_count(const poy_store *st);
POY_API int64_t poy_store_bytes(const poy_store *st);
POY_API poy_status poy_store_append(poy_ctx *ctx, poy_store *st, const uint8_t *data, const int64_t *offsets, int32_t nseq,
                                    int32_t *first_id);
POY_API poy_status poy_store_truncate(poy_ctx *ctx, poy_store *st, int32_t nseq);
POY_API poy_status poy_store_lengths(const poy_store *st, int32_t n, const int32_t *ids, int32_t *len);
POY_API poy_status poy_store_read(poy_ctx *ctx, const poy_store *st, int32_t n, const int32_t *ids, const int64_t *out_off,
                                  uint8_t *out);
POY_API poy_status poy_store_median(poy_ctx *ctx, poy_store *st, const poy_cm *c2_full, const poy_cm *c2_original, int32_t n,
                                    const int32_t *a, const int32_t *b, int32_t *out_id, int32_t *out_len, int32_t *cost2);
POY_API poy_status poy_store_distance(poy_ctx *ctx, poy_store *st, const poy_cm *c2_original, int32_t n, const int32_t *a,
                                      const int32_t *b, int32_t missing_distance, int32_t *cost);

/* ---- INT32 / DPX issue-rate micro-benchmark (roofline denominator) ---------
 * Runs independent chains of one instruction class at full occupancy and
 * returns thread-level operations per second.  kind: 0 IADD3, 1 IMNMX (min),
 * 2 VIADDMNMX (__viaddmin_s32), 3 VIMNMX3 (__vimin3_s32), 4 mixed add+viaddmin
 * in the ratio of the cost-only cell. */
POY_API poy_status poy_microbench_int(poy_ctx *ctx, int32_t kind, double *ops_per_second, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* POY5_B200_H */
