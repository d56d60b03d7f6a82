// Internal declarations shared by the kernels and the host side of libpoy5b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "../../include/poy5_b200.h"

#define POY_INF 1000000  // HIGH_NUM, src/algn.c:37 -- added to, never saturated
#define POY_GAP 16       // TMPGAP, src/algn.c:1202
#define POY_NOGAP 15     // NTMPGAP

// ---- device-resident cost model ------------------------------------------------
struct DevCM {
    int cost16[256];     // cost[(a&15)][(b&15)], the only part the affine kernels index (src/algn.c:2073, A5)
    int cost32[1024];    // full table (linear-gap kernels, worst/verify)
    int worst32[1024];
    uint8_t median32[1024];
    uint8_t closest32[1024]; // Cost_matrix.Two_D.get_closest a b at [(a << 5) + b] (src/cost_matrix.ml:1387-1428)
    int prepend[32];
    int tail[32];
    int gapext[32];      // cost[a][gap]  (HAS_GAP_EXTENSION, src/algn.c:1220)
    int gap_open;
    int model;
    int min_non0;
    int max_entry;       // largest table entry; host-side domain check
};

// ---- per-base gap parameters, one int4 per base per role ------------------------
// x = extension cost when the gap continues through this base
//     (row role: si_vertical_extension, src/algn.c:2063-2065; column role:
//      sj_horizontal_extension after the [1] overwrite, :2028-2035)
// y = opening cost: go + ge  (si_gap_opening + si_gap_extension)
// z = go                     (HAS_GAP_OPENING, :1240-1253)
// w = flags: bits 0-3 code&15, bit 4 code has the gap bit, bit 5 previous base has the gap bit
#define PF_HASGAP 16
#define PF_PREVGAP 32
// precomputed window-entry fields of the band kernel in the same word (k_params writes them, band2.cu masks them out)
#define PF_ROW_CLASS_SHIFT 8      // row role: surcharge class, bits 8-10
#define PF_ROW_GF_SHIFT 11        // row role: code & 15 at bits 11-14 (gap-free table row offset)
#define PF_ROW_TAB_SHIFT 27       // row role: code & 15 at bits 27-30 (table row offset << 16)
#define PF_ROW_MASK 0x78000700
#define PF_ROW_GF_MASK 0x00007800
#define PF_COL_LB_SHIFT 12        // column role: class on the left border, bits 12-14 (class itself: bits 5-7)
#define PF_COL_TAB_SHIFT 23       // column role: code & 15 at bits 23-26 (table column offset << 16)
#define PF_COL_MASK 0x078070E0

// what k_params reads from a cost model: pools (and the node store) remember the signature their per-base parameters
// were computed for, so switching between c2_full and c2_original (identical under the default affine tables,
// src/seqCS.ml:51-55) does not recompute them
struct ParamSig {
    int prepend[32], gapext[32], gap_open, valid;
};

struct poy_cm {
    uint64_t uid;          // unique per upload
    DevCM *d;
    poy_cm_host h;
    int min_non0, max_entry;
    ParamSig sig;
};

struct poy_pool {
    uint8_t *d_data;       // packed codes
    int64_t *d_off;        // nseq+1 offsets
    int4 *d_rowp;          // per-base parameters, row role
    int4 *d_colp;          // per-base parameters, column role
    unsigned *d_rowpk;     // row role, packed for the gap-free cost kernel: prefix sum of cost[s][gap] | table row << 28
    int *d_h0;             // banded entry point: CB[0][j] = sum of in-loop hext (src/algn.c:2244)
    int *d_g0;             // cost-only entry point: EH[0][j] - GO = sum of prepend (src/algn.c:1847)
    uint8_t *d_gapfree;    // per sequence: 1 if no base at index >= 1 carries the gap bit
    int64_t *h_off;        // host copy of the offsets
    uint8_t *h_gapfree;    // host copy of d_gapfree (valid once the parameters have been computed)
    uint8_t *h_empty;      // per sequence: every symbol is the gap code (Sequence.is_empty); filled on first use by dos.cu
    int32_t *h_gapcnt;     // per sequence: symbols that carry the gap bit (Sequence.count_gaps); filled with h_empty
    int32_t nseq;
    int64_t nbytes;
    bool owns_data;
    ParamSig sig;              // signature of the cost model the parameters were computed for (valid = 0: none)
    int32_t params_upto;       // sequences [0, params_upto) have parameters for `sig`
    int32_t flags_upto;        // sequences [0, flags_upto) have h_empty / h_gapcnt
    uint8_t *d_flags;          // device staging of the per-sequence flags (empty | gap count), 8 bytes per sequence
    size_t caps[9];            // capacities of the cached device blocks backing the arrays above
    // node store (store.cu): capacities the arrays above were allocated for; 0 for plain pools
    int64_t cap_bytes;
    int32_t cap_seqs;
};

#define POY_N_AUX 8
struct poy_ctx {
    int device;
    cudaStream_t stream;
    bool owns_stream;
    int sm_count;
    uint64_t arena_limit;
    uint64_t launches;
    char err[512];
    // grow-only device scratch
    void *d_scratch[14];
    size_t scratch_cap[14];
    void *h_pinned[6];
    size_t pinned_cap[6];
    // auxiliary streams + events: independent launches of one wave run concurrently so that the tail of one
    // overlaps the body of the next (8: a latency-bound round has up to 6-8 band classes, none should queue behind another)
    cudaStream_t aux[POY_N_AUX];
    cudaEvent_t ev_fork, ev_join[POY_N_AUX];
    // traceback lane: the traceback of the pairs that stopped in round r runs on its own stream while round r+1 is
    // being filled (two direction arenas / job arrays in turn); ev_tb_done[p] guards the reuse of arena p
    cudaStream_t tb_stream;
    cudaEvent_t ev_fin, ev_tb_done[2];
    bool tb_pending[2];
    // small cache of device blocks released by freed pools: tree-search drivers create and destroy thousands of
    // short-lived pools, and cudaMalloc / cudaFree (a device-wide sync) would dominate their batches
    void *cache_ptr[64];
    size_t cache_cap[64];
    int cache_n;
    size_t cache_bytes;
    int64_t stat_band_cells, stat_probe, stat_full, stat_repeat, stat_rounds, stat_pairs;   // poy_ctx_stats
    int64_t hint_max_len;  // set by the host entry points for the next *_dev call: longest sequence among the submitted pairs
    // second lane for large banded batches: the batch is cut in two halves whose threshold-doubling rounds run as
    // independent pipelines (own stream set, arenas and host thread), so that the tail of one half's round is
    // filled by the other half's kernels
    poy_ctx *twin;
    bool is_twin;
    cudaEvent_t ev_twin_start, ev_twin_done;
};

// ---- work descriptors ----------------------------------------------------------------
struct CostJob {          // one cost-only alignment; rows = shorter sequence
    int64_t off_i, off_j; // pool offsets of the row / column sequence
    int lasti, lastj;     // last valid index (= len - 1)
    int out;              // index into the cost output array
    int gapfree;          // both sequences free of gap bits
};

struct BandJob {          // one band fill of one pair
    int64_t off_i, off_j;
    int lasti, lastj;
    int k;                // half band (already clamped, src/algn.c:2195-2196)
    int pair;             // index into the per-pair state arrays
    int swaped;           // bit 0 operands were exchanged by the caller, bit 1 full plane (linear), bit 2 gap-free pair,
                          // bit 3 probe fill (no direction bytes), bit 4 no traceback wanted (a probe's verdict is final),
                          // bit 5 repeat of a probed threshold: restore the stale state the probe started from,
                          // bit 6 the direction bytes are in the tagged format of k_band2 (cell_gf / cell_gen)
    int stride;           // bytes per anti-diagonal in the direction arena
    int64_t dir_off;      // byte offset of this pair's direction block in the arena
    int64_t eb_off;       // int offset of this pair's stale-EB row in the state arena
    // speculative threshold doublings (host.cu, align_impl; swaped bit 7 set): several fills of one pair run at once,
    // each behind its predecessor.  dep = progress slot of the previous doubling's fill (-1: none), need = the row that
    // fill must have completed before this one may read the stale EB row (SPEC_DONE: all of it), T / eh00 = threshold
    // and EH[0][0] of THIS fill (the per-slot PairState then only receives results)
    int dep, need, T, eh00;
};
#define SPEC_DONE 0x40000000      // progress value of a completed fill whose stop rule did not fire
#define SPEC_STOP 0x40000001      // ... of a completed fill whose stop rule fired: the doublings after it are not needed
#define SPEC_ABORT 0x40000002     // ... of a fill abandoned because an earlier doubling stopped
#define SPEC_FAILED (-1)          // ... of one that gave up waiting for its predecessor (the host re-runs it)

struct PairState {        // survives across band fills of the same pair
    int T;                // current threshold
    int eh00;             // EH[0][0] as left behind by the previous fill
    int eh00_snap;        // ... and as it was before the latest probe fill
    int cost;             // result of the latest fill
    int gapnum;           // max gap-count of the latest fill
    int iterations;
    int done;
    long long cells;
};

// ---- host-side helpers shared by host.cu / dos.cu / store.cu -------------------------------------------------------
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return poy_cuda_fail(ctx, e_, #call); } while (0)
enum { SL_JOBS = 0, SL_JOBS2 = 1, SL_BOUND = 2, SL_STATE = 3, SL_EBROW = 4, SL_DIR = 5, SL_MISC = 6, SL_WORK = 7,
       SL_STORE = 8, SL_STORE2 = 9, SL_DIR2 = 10, SL_JOBSB = 11, SL_COUNT = 12 };
poy_status poy_fail(poy_ctx *ctx, poy_status s, const char *msg);
poy_status poy_cuda_fail(poy_ctx *ctx, cudaError_t e, const char *where);
poy_status poy_scratch(poy_ctx *ctx, int slot, size_t bytes, void **out);     // grow-only device scratch slot
poy_status poy_pinned(poy_ctx *ctx, int slot, size_t bytes, void **out);      // grow-only pinned host slot
cudaError_t cached_alloc(poy_ctx *ctx, void **out, size_t bytes, size_t *cap_out);
void cached_free(poy_ctx *ctx, void *p, size_t cap);
// Every entry point binds the calling thread to the context's device first: host threads other than the one that
// created the context start out on device 0.
static inline void bind_device(const poy_ctx *ctx) { if (ctx) cudaSetDevice(ctx->device); }
// banded alignment of n pairs (affine; linear when h_deltawh != NULL); all d_* are device pointers, h_* host arrays
poy_status align_split(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int32_t n, const int32_t *h_si,
                       const int32_t *h_sj, const uint8_t *h_swaped, const int64_t *d_out_off, int32_t *d_cost,
                       uint8_t *d_median, uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj,
                       int32_t *d_out_len, int32_t *h_stats, const int32_t *h_deltawh = nullptr);

// launchers (defined in the .cu files)
cudaError_t launch_params(poy_ctx *ctx, const poy_cm *cm, poy_pool *pool, int s0, int s1);
cudaError_t launch_seq_flags(poy_ctx *ctx, const poy_pool *pool, int s0, int s1, int2 *d_flags);
poy_status ensure_params(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool);
poy_status ensure_flags(poy_ctx *ctx, const poy_pool *pool);
cudaError_t launch_build_cost_jobs(poy_ctx *ctx, const poy_pool *pool, int n, const int *d_a, const int *d_b,
                                   CostJob *d_jobs, int *d_counts);
cudaError_t launch_cost_affine(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const CostJob *d_jobs, int njobs,
                               int *d_counter, int4 *d_bound, size_t bound_stride, int blocks, int *d_cost, int wide);
cudaError_t launch_band2(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs, int cls,
                         bool gapfree, bool probe, int *d_counter, PairState *d_state, int *d_ebrow, uint8_t *d_dir, bool lowlat = false,
                         int *d_prog = nullptr);
int band2_class_for(long long B);
int band2_stride_for(int cls, long long B);
cudaError_t launch_band_lin(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs, int cls,
                            int *d_counter, PairState *d_state, uint8_t *d_dir);
cudaError_t launch_band_lin_generic(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                    PairState *d_state, uint8_t *d_dir, int *d_work, size_t work_stride, int blocks);
cudaError_t launch_lin_finish(poy_ctx *ctx, const poy_pool *pool, const BandJob *d_jobs, int njobs, PairState *d_state, uint8_t *d_done);
cudaError_t launch_traceback_lin(poy_ctx *ctx, const poy_pool *pool, const BandJob *d_jobs, int njobs, const uint8_t *d_done,
                                 const uint8_t *d_dir, const int64_t *d_out_off, uint8_t *d_r1, uint8_t *d_r2, int *d_out_len);
cudaError_t launch_band_generic(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                PairState *d_state, int *d_ebrow, uint8_t *d_dir, int *d_work, size_t work_stride,
                                int blocks);
cudaError_t launch_band_finish(poy_ctx *ctx, const BandJob *d_jobs, int njobs, PairState *d_state, uint8_t *d_done,
                               const int *d_g0, int gap_open, const int *d_prog = nullptr);
cudaError_t launch_traceback(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                             const uint8_t *d_done, const uint8_t *d_dir, const int64_t *d_out_off, uint8_t *d_median,
                             uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj, int *d_out_len);
cudaError_t launch_stale_snapshot(poy_ctx *ctx, const BandJob *d_jobs, int njobs, PairState *d_state, int *d_eb, int *d_eb_snap);
cudaError_t launch_gather_cost(poy_ctx *ctx, const PairState *d_state, int n, int *d_cost);
cudaError_t launch_fill_int(poy_ctx *ctx, int *d, int64_t n, int v);
cudaError_t launch_microbench(poy_ctx *ctx, int kind, int iters, unsigned long long *d_cycles, int *d_sink,
                              float *ms, double *ops);

cudaError_t launch_median_2(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                            const int *len, int with_gaps, const int64_t *out_off, uint8_t *out, int *out_len);
cudaError_t launch_union(poy_ctx *ctx, int64_t total, const uint8_t *a, const uint8_t *b, uint8_t *out);
cudaError_t launch_aligned_cost(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                                const int *len, int use_worst, int *cost);
// swap (optional, device): per pair, exchange the two rows; cap_adjust: unused spare bytes at the slot end (documentation only)
cudaError_t launch_ancestor_2(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                              const int *len, const int64_t *out_off, uint8_t *out, int *out_len, const uint8_t *swap = nullptr,
                              int cap_adjust = 0);
poy_status dos_self_recost(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, int n, const int32_t *ids, int32_t *cost);
poy_status dos_median_device(poy_ctx *ctx, const poy_cm *c2_full, const poy_pool *pool, int m, const int32_t *a, const int32_t *b,
                             const int64_t *h_slot, const int64_t *d_slot, const int64_t *d_slot_end, int64_t slot_total,
                             int32_t *d_cost, uint8_t *d_median, int32_t *d_mlen, bool *right_justified);
poy_status pool_alloc(poy_ctx *ctx, poy_pool *p, int64_t nb, int32_t ns);
poy_status pool_new(poy_ctx *ctx, const int64_t *h_off, int32_t nseq, int32_t cap_seqs, poy_pool **out);
