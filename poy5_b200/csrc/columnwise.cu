// O(L) column-wise helpers over pairs of ALIGNED rows (batch twins):
//   seq_CAML_median_2_with_gaps / _no_gaps  (src/seq.c:241-272)
//   algn_CAML_union                          (src/algn.c:3657-3678)
//   algn_CAML_worst_2 / algn_CAML_verify_2   (src/algn.c:3003-3130)  -- cost of an aligned pair under the
//                                             worst / cost table with the gap-opening automaton
//   algn_CAML_ancestor_2                     (src/algn.c:3603-3626, 3561-3601, 3539-3559)
// Rows are packed byte buffers; pair p uses a[off[p] .. off[p]+len[p]) and b[off[p] ..).
#include "common.cuh"

// warp per pair: medians of all columns; without gaps the pure-gap medians are squeezed out and the
// leading gap restored (ballot compaction keeps the order).  with_gaps == 2 maps the columns through the
// get_closest table instead (a = parent row, b = own row) and squeezes the gaps out the same way.
__global__ void __launch_bounds__(128)
k_median_2(const DevCM *__restrict__ cm, int n, const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
           const int64_t *__restrict__ off, const int *__restrict__ len, int with_gaps, const int64_t *__restrict__ out_off,
           uint8_t *out, int *out_len) {
    __shared__ uint8_t s_med[1024];
    for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_med[x] = with_gaps == 2 ? cm->closest32[x] : cm->median32[x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= n) return;
    const uint8_t *ra = a + off[p], *rb = b + off[p];
    uint8_t *o = out + out_off[p];
    const int L = len[p];
    int w = 0;
    if (with_gaps != 1) { if (lane == 0) o[0] = POY_GAP; w = 1; }
    for (int x0 = 0; x0 < L; x0 += 32) {
        const int x = x0 + lane;
        int m = 0; bool keep = false;
        if (x < L) { m = s_med[((ra[x] & 31) << 5) + (rb[x] & 31)]; keep = with_gaps == 1 || m != POY_GAP; }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) o[w + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)m;
        w += __popc(mask);
    }
    if (lane == 0) out_len[p] = w;
}

__global__ void k_union(int64_t total, const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, uint8_t *out) {
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x)
        out[x] = a[x] | b[x];
}

// thread per pair: algn_calculate_from_2_aligned (bitset alphabet branch)
__global__ void k_aligned_cost(const DevCM *__restrict__ cm, int n, const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                               const int64_t *__restrict__ off, const int *__restrict__ len, int use_worst, int *cost) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int *table = use_worst == 1 ? cm->worst32 : cm->cost32;
    const uint8_t *s1 = a + off[p], *s2 = b + off[p];
    const int L = len[p], go = cm->gap_open;
    if (use_worst >= 2) {
        const int gopen = cm->model == 1 ? cm->gap_open : 0;
        // Sequence.Align.recost ?first_gap a b cm (src/sequence.ml:1244-1307, combination branch): one gap opening per
        // block of columns in which either symbol carries the gap bit; use_worst == 2: first_gap = true (column 0 is
        // the shared leading gap and is skipped), 3: first_gap = false
        int res = 0;
        bool blk = false;
        for (int i = use_worst == 2 ? 1 : 0; i < L; ++i) {
            const int x = s1[i] & 31, y = s2[i] & 31;
            const bool g = ((x | y) & POY_GAP) != 0;
            res += cm->cost32[(x << 5) + y] + ((g && !blk) ? gopen : 0);
            blk = g;
        }
        cost[p] = res;
        return;
    }
    int res = 0, gap_row = 0;
    int i = (L > 0 && (s1[0] & POY_GAP) && (s2[0] & POY_GAP)) ? 1 : 0;
    for (; i < L; ++i) {
        const int x = s1[i] & 31, y = s2[i] & 31;
        if (gap_row == 0) {
            if ((x & POY_GAP) && !(y & POY_GAP)) { res += go; gap_row = 1; }
            else if ((y & POY_GAP) && !(x & POY_GAP)) { res += go; gap_row = 2; }
        } else if (gap_row == 1) {
            if (!(x & POY_GAP)) {
                if (y & POY_GAP) { res += go; gap_row = 2; }
                else gap_row = 0;
            }
        } else {
            if (!(y & POY_GAP)) {
                if (x & POY_GAP) { res += go; gap_row = 1; }
                else gap_row = 0;
            }
        }
        res += table[(x << 5) + y];
    }
    cost[p] = res;
}

// thread per pair: algn_ancestor_2.  out slot capacity len+1.  A zero median (failwith "median should not be 0")
// is reported as out_len = -1.
__global__ void k_ancestor_2(const DevCM *__restrict__ cm, int n, const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                             const int64_t *__restrict__ off, const int *__restrict__ len, const int64_t *__restrict__ out_off,
                             uint8_t *out, int *out_len, const uint8_t *__restrict__ swap) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const bool sw = swap && swap[p];
    const uint8_t *s1 = (sw ? b : a) + off[p], *s2 = (sw ? a : b) + off[p];
    uint8_t *o = out + out_off[p];
    const int L = len[p], gap = POY_GAP;
    const bool affine = cm->model == 1;
    int w = 0, extending = 0, inside = 0, prev_block = 0;
    bool bad = false;
    o[w++] = (uint8_t)gap;   // both branches end by (re)placing the leading gap in front of the gap-free medians
    for (int i = 0; i < L; ++i) {
        const int ab = s1[i] & 31, bb = s2[i] & 31;
        int sb = cm->median32[(ab << 5) + bb];
        if (sb == 0) bad = true;
        if (affine) {        // algn_correct_blocks_affine
            if (!inside && (!(ab & gap) || !(bb & gap))) inside = 0;
            else if (inside && (!(ab & gap) || !(bb & gap))) inside = 0;
            else if (((ab & gap) || (bb & gap)) && ((ab != gap) || (bb != gap))) inside = 1;
            else inside = 0;
            if (((gap & ab) || (gap & bb)) && !(sb & gap) && !extending) { prev_block = inside; extending = 1; }
            else if ((gap & ab) && (gap & bb) && (sb & gap) && (sb != gap) && extending && inside && !prev_block) { sb = (~gap) & sb; prev_block = 0; }
            else if ((gap & ab) && (gap & bb) && extending == 1) { prev_block = inside; extending = 0; }
        }
        if (sb != gap) o[w++] = (uint8_t)sb;
    }
    out_len[p] = bad ? -1 : w;
}

cudaError_t launch_median_2(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                            const int *len, int with_gaps, const int64_t *out_off, uint8_t *out, int *out_len) {
    if (n <= 0) return cudaSuccess;
    k_median_2<<<(n + 3) / 4, 128, 0, ctx->stream>>>(cm->d, n, a, b, off, len, with_gaps, out_off, out, out_len);
    ctx->launches++;
    return cudaGetLastError();
}
cudaError_t launch_union(poy_ctx *ctx, int64_t total, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    if (total <= 0) return cudaSuccess;
    int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_union<<<blocks, 256, 0, ctx->stream>>>(total, a, b, out);
    ctx->launches++;
    return cudaGetLastError();
}
cudaError_t launch_aligned_cost(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                                const int *len, int use_worst, int *cost) {
    if (n <= 0) return cudaSuccess;
    k_aligned_cost<<<(n + 127) / 128, 128, 0, ctx->stream>>>(cm->d, n, a, b, off, len, use_worst, cost);
    ctx->launches++;
    return cudaGetLastError();
}
cudaError_t launch_ancestor_2(poy_ctx *ctx, const poy_cm *cm, int n, const uint8_t *a, const uint8_t *b, const int64_t *off,
                              const int *len, const int64_t *out_off, uint8_t *out, int *out_len, const uint8_t *swap, int) {
    if (n <= 0) return cudaSuccess;
    k_ancestor_2<<<(n + 127) / 128, 128, 0, ctx->stream>>>(cm->d, n, a, b, off, len, out_off, out, out_len, swap);
    ctx->launches++;
    return cudaGetLastError();
}
