// Ukkonen-banded affine DO alignment with direction matrix, and the traceback that
// builds the median: batch twin of algn_CAML_align_affine_3 (src/algn.c:2359-2447) =
// algn_newkk_test_aff / algn_newkk_fill_a_row_aff / ASSIGN_MINIMUM / algn_fill_gapnum
// (src/algn.c:2186-2306, 2113-2180, 1936-1981, 126-176) + backtrace_aff (:1715-1819).
//
// Band geometry (src/algn.c:2256-2257, 2130-2131): row i covers columns
// [max(i-k,0), min(i+delta+k, lastj)], delta = lastj-lasti; the first cell of a row is
// a "left border" (EH = INF), the cell j = i+delta+k a "right border" (EV = INF).
// In diagonal coordinates d = j - i + k the band is d in [0, B), B = delta + 2k + 1.
//
// Parallelisation (register-resident kernels: band2.cu; this file keeps the reference-shaped
// cell, the generic any-width fallback and the traceback): anti-diagonal wavefront.  Thread t
// owns D consecutive diagonals [t*D, (t+1)*D) and keeps, per diagonal, only the state of the
// LATEST cell on it (CB, EV, EH, EB and the two gap counters) in registers.  Cells of
// one anti-diagonal a = i + j all have d = a + k (mod 2), so a step updates every
// other diagonal in place: the left neighbour (i, j-1) is the latest cell of diagonal
// d-1, the upper neighbour (i-1, j) the latest cell of d+1, the diagonal neighbour the
// previous cell of d itself.  Only the strip edges cross lanes: one __shfl_up (even
// sub-step) or one __shfl_down (odd sub-step) of four values.
//
// Direction storage: the reference keeps a (leni x (lenj+1)) unsigned-short matrix
// (0.2 GB at 10 kb); here only band cells are stored, one byte each, anti-diagonal
// major so that a warp's stores of one step are contiguous: cell (i,j) lives at
// dir[(i+j) * stride + ((j-i+k) >> 1)].  The 15-bit mask of src/algn.c:1185-1199 is
// reduced, using the per-pair `swaped` flag that fixes the traceback priorities
// (choose_dir, :1594-1619), to the 7 bits the traceback can observe:
//   bits 0-1  mode chosen from `todo`   (0 vertical, 1 horizontal, 2 block diagonal, 3 align)
//   bits 2-3  mode chosen after an align step (same coding, from the ALIGN_TO_* ties)
//   bit 4     vertical run stops here   (END_VERTICAL | HORIZONTAL_EQ_VERTICAL)
//   bit 5     horizontal run stops here (END_HORIZONTAL | HORIZONTAL_EQ_VERTICAL)
//   bit 6     block-diagonal run stops here (END_BLOCK)
// This plain format is what the fallback kernel below writes.  The register-resident kernels of band2.cu write a
// "tagged" variant (same fields; the two gap codes are stored in priority order instead of V/H order, the END_*
// bits inverted, bit 7 = HORIZONTAL_EQ_VERTICAL, no ALIGN_TO field for gap-free pairs); BandJob::swaped bit 6
// tells k_traceback which one it is reading.
//
// State that the reference leaves behind from one band fill to the next (its row
// buffers are not re-initialised between threshold doublings): row 0 of EB and
// EH[0][0].  Both are carried explicitly per pair (PairState.eh00, the eb row).
#include <type_traits>
#include "common.cuh"

struct CellIn {
    int lCB, lEH, lG1, lG2;   // (i, j-1)
    int uCB, uEV, uG1, uG2;   // (i-1, j)
    int dCB, dEV, dEH, dEB, dG1, dG2;  // (i-1, j-1)
};
struct CellOut {
    int CB, EV, EH, EB, G1, G2;
    int fin;
    unsigned dirbyte;
};

// compile-time loop: indices reach the body as constants, so the per-diagonal state arrays are
// scalar-replaced into registers (a plain `#pragma unroll` loop left them in local memory)
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

__device__ __forceinline__ int imax_(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int imin_(int a, int b) { return a < b ? a : b; }

// One band cell.  r = row parameters of i, c = column parameters of j.
__device__ __forceinline__ void band_cell(const CellIn &in, const int4 r, const int4 c, const int *s_cost16, int GO,
                                          bool lb, bool rb, bool jpos, int swaped, CellOut &o) {
    unsigned stopH = 0, stopV = 0, stopB = 0;
    int EH, EV, EB, CB;
    if (!lb) {
        const int ext = in.lEH + c.x, opn = in.lCB + c.y;
        if (ext < opn) EH = ext; else { EH = opn; stopH = 1; }
    } else EH = POY_INF;
    if (!rb) {
        const int ext = in.uEV + r.x, opn = in.uCB + r.y;
        if (ext < opn) EV = ext; else { EV = opn; stopV = 1; }
    } else EV = POY_INF;
    unsigned a2 = 0;  // bit0 A2A, bit1 A2V, bit2 A2H, bit3 A2D
    if (jpos) {
        // at the left border the reference's "previous column symbol" is the column symbol itself
        const bool pg_j = lb ? ((c.w & PF_HASGAP) != 0) : ((c.w & PF_PREVGAP) != 0);
        const bool both = (r.w & c.w & PF_HASGAP) != 0;
        const bool clean = !(r.w & PF_PREVGAP) && !pg_j;
        const int dg = both ? 0 : POY_INF;
        const int od = both ? (clean ? 0 : 2 * GO) : POY_INF;
        {
            const int ext = in.dEB + dg, opn = in.dCB + od;
            if (ext < opn) EB = ext; else { EB = opn; stopB = 1; }
        }
        const int diag = s_cost16[(r.w & 15) * 16 + (c.w & 15)];
        const int xgo = c.z < r.z ? r.z : c.z;
        int a = in.dCB + diag;
        const int v = (r.w & PF_HASGAP) ? in.dEV + diag + c.z : in.dEV + diag;
        const int h = (c.w & PF_HASGAP) ? in.dEH + diag + r.z : in.dEH + diag;
        const int d = in.dEB + diag + xgo;
        a2 = 1;
        if (a >= v) { if (a > v) { a = v; a2 = 2; } else a2 |= 2; }
        if (a >= h) { if (a > h) { a = h; a2 = 4; } else a2 |= 4; }
        if (a >= d) { if (a > d) { a = d; a2 = 8; } else a2 |= 8; }
        CB = a;
    } else { CB = POY_INF; EB = POY_INF; }
    // final minimum, order H, V, D, A (ASSIGN_MINIMUM)
    unsigned m = 4;  // bit0 DO_A, bit1 DO_V, bit2 DO_H, bit3 DO_D
    int fin = EH;
    if (fin >= EV) { if (fin > EV) { fin = EV; m = 2; } else m |= 2; }
    if (fin >= EB) { if (fin > EB) { fin = EB; m = 8; } else m |= 8; }
    if (fin >= CB) { if (fin > CB) { fin = CB; m = 1; } else m |= 1; }
    if (fin == EH && EH == EV) { stopH = 1; stopV = 1; }
    // gap counters (algn_fill_gapnum): max-plus over the chosen predecessors, unsigned short
    int c1 = -1, c2 = -1;
    if (m & 9) { c1 = in.dG1; c2 = in.dG2; }
    if (m & 4) { c1 = imax_(c1, in.lG1 + 1); c2 = imax_(c2, in.lG2); }
    if (m & 2) { c1 = imax_(c1, in.uG1); c2 = imax_(c2, in.uG2 + 1); }
    o.G1 = c1 & 0xFFFF; o.G2 = c2 & 0xFFFF;
    // compress to the traceback byte
    unsigned todo, nxt;
    if (!swaped) {
        todo = (m & 2) ? 0u : (m & 4) ? 1u : (m & 8) ? 2u : 3u;
        nxt = (a2 & 2) ? 0u : (a2 & 4) ? 1u : (a2 & 8) ? 2u : 3u;
    } else {
        todo = (m & 4) ? 1u : (m & 2) ? 0u : (m & 8) ? 2u : 3u;
        nxt = (a2 & 4) ? 1u : (a2 & 2) ? 0u : (a2 & 8) ? 2u : 3u;
    }
    o.dirbyte = todo | (nxt << 2) | (stopV << 4) | (stopH << 5) | (stopB << 6);
    o.CB = CB; o.EV = EV; o.EH = EH; o.EB = EB; o.fin = fin;
}

// ---- generic fallback: one CTA per pair, diagonal state in global memory -------------------
// Used for bands wider than the register-resident kernel covers (very dissimilar pairs).
// work layout per job: 6 planes of `wstride` ints (CB, EV, EH, EB, G1, G2).
__global__ void __launch_bounds__(256)
k_band_generic(const DevCM *__restrict__ cm, const int4 *__restrict__ rowp, const int4 *__restrict__ colp,
               const int *__restrict__ h0v, const BandJob *__restrict__ jobs, int njobs, PairState *state, int *ebrow,
               uint8_t *dir, int *work, size_t work_stride) {
    __shared__ int s_cost16[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_cost16[x] = cm->cost16[x];
    __syncthreads();
    const int GO = cm->gap_open;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const BandJob J = jobs[job];
        const int lasti = J.lasti, lastj = J.lastj, k = J.k, swaped = J.swaped & 1;
        if (lasti == 0) continue;
        const int delta = lastj - lasti, B = delta + 2 * k + 1;
        const int4 *rp = rowp + J.off_i;
        const int4 *cp = colp + J.off_j;
        const int *h0 = h0v + J.off_j;
        int *eb = ebrow + J.eb_off;
        PairState *st = state + J.pair;
        uint8_t *dbase = dir + J.dir_off;
        const int stride = J.stride;
        const size_t ws = work_stride / 6;
        int *wCB = work + (size_t)blockIdx.x * work_stride, *wEV = wCB + ws, *wEH = wEV + ws, *wEB = wEH + ws,
            *wG1 = wEB + ws, *wG2 = wG1 + ws;
        const int eh00 = st->eh00;
        for (int d = threadIdx.x; d < B; d += blockDim.x) {
            const int j0 = d - k;
            if (j0 >= 0 && j0 <= lastj) {
                wCB[d] = h0[j0]; wEH[d] = j0 == 0 ? eh00 : h0[j0]; wEV[d] = POY_INF; wEB[d] = eb[j0];
                wG1[d] = j0 & 0xFFFF; wG2[d] = 0;
            } else {
                wCB[d] = wEV[d] = wEH[d] = wEB[d] = POY_INF; wG1[d] = wG2[d] = 0;
            }
        }
        __syncthreads();
        const int a_end = lasti + lastj;
        for (int a = 1; a <= a_end; ++a) {
            const int par = (a + k) & 1;
            // valid diagonals: i = (a-d+k)/2 in [1,lasti], j = a-i in [0,lastj]
            for (int d = 2 * threadIdx.x + par; d < B; d += 2 * blockDim.x) {
                const int i = (a - d + k) >> 1, j = a - i;
                if (i >= 1 && i <= lasti && j >= 0 && j <= lastj) {
                    CellIn in;
                    if (d > 0) { in.lCB = wCB[d - 1]; in.lEH = wEH[d - 1]; in.lG1 = wG1[d - 1]; in.lG2 = wG2[d - 1]; }
                    else { in.lCB = in.lEH = POY_INF; in.lG1 = in.lG2 = 0; }
                    if (d + 1 < B) { in.uCB = wCB[d + 1]; in.uEV = wEV[d + 1]; in.uG1 = wG1[d + 1]; in.uG2 = wG2[d + 1]; }
                    else { in.uCB = in.uEV = POY_INF; in.uG1 = in.uG2 = 0; }
                    in.dCB = wCB[d]; in.dEV = wEV[d]; in.dEH = wEH[d]; in.dEB = wEB[d]; in.dG1 = wG1[d]; in.dG2 = wG2[d];
                    CellOut o;
                    band_cell(in, rp[i], cp[j], s_cost16, GO, d == 0 || j == 0, d == B - 1, j > 0, swaped, o);
                    wCB[d] = o.CB; wEV[d] = o.EV; wEH[d] = o.EH; wEB[d] = o.EB; wG1[d] = o.G1; wG2[d] = o.G2;
                    dbase[(size_t)a * stride + (d >> 1)] = (uint8_t)o.dirbyte;
                    if (!(i & 1) && (d <= 1 || i >= lasti - 1)) eb[j] = o.EB;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int d = delta + k;
            st->cost = imin_(imin_(wEH[d], wEV[d]), imin_(wEB[d], wCB[d]));
            st->gapnum = imax_(wG1[d], wG2[d]);
            if (imin_(k, lasti) >= 2) st->eh00 = POY_INF;
        }
        __syncthreads();
    }
}

// ---- traceback: backtrace_aff (src/algn.c:1715-1819) -------------------------------------------
// One WARP per pair.  The walk itself is a serial pointer chase (lane 0), so what matters is the
// latency of each step: the warp stages a 64 x 32-byte tile of direction bytes around the current
// cell in shared memory with one coalesced load (64 anti-diagonals x 64 diagonals), lane 0 walks
// until it leaves the tile recording one move code per step -- nothing but the direction byte and the
// mode is on that chain -- and the 32 lanes then emit what the steps produce (symbol loads, median
// look-ups, four output streams) in parallel; the tile is re-centred.  Results are written right to left into the
// caller's capacity-(leni+lenj+2) slots, i.e. exactly like seq_prepend fills a struct seq.
#define TB_ROWS 64
#define TB_COLS 32
__global__ void __launch_bounds__(128)
k_traceback(const DevCM *__restrict__ cm, const uint8_t *__restrict__ data, const BandJob *__restrict__ jobs, int njobs,
            const uint8_t *__restrict__ done, const uint8_t *__restrict__ dir, const int64_t *__restrict__ out_off,
            uint8_t *median, uint8_t *medianwg, uint8_t *resi, uint8_t *resj, int *out_len) {
    __shared__ __align__(16) uint8_t s_tile[4][TB_ROWS * TB_COLS];
    __shared__ uint8_t s_moves[4][TB_ROWS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 4 + w;
    if (t >= njobs) return;
    const BandJob J = jobs[t];
    if (done && done[J.pair] != 1) return;  // this pair's band is not final yet (or was only probed)
    const uint8_t *si = data + J.off_i, *sj = data + J.off_j;
    const int k = J.k;
    const bool tagged = (J.swaped & 64) != 0, swaped = (J.swaped & 1) != 0, gf_todo = tagged && (J.swaped & 4) != 0;
    const int B = (J.lastj - J.lasti) + 2 * k + 1;
    const uint8_t *db = dir + J.dir_off;
    const int stride = J.stride;
    const int cap = J.lasti + J.lastj + 4;  // len_i + len_j + 2
    const int64_t base = out_off ? out_off[J.pair] : 0;
    uint8_t *pm = median ? median + base + cap : nullptr;
    uint8_t *pw = medianwg ? medianwg + base + cap : nullptr;
    uint8_t *pi = resi ? resi + base + cap : nullptr;
    uint8_t *pj = resj ? resj + base + cap : nullptr;
    int nm = 0, nw = 0, ni = 0, nj = 0;
    int first_m = -1;  // value currently at the front of median
#define PUT(ptr, cnt, v) do { ++(cnt); if (ptr) *(--(ptr)) = (uint8_t)(v); } while (0)
#define PUT_M(v) do { first_m = (v); PUT(pm, nm, v); } while (0)
#define INDEL(sym) do { if (!((sym) & POY_GAP)) { PUT_M((sym) | POY_GAP); PUT(pw, nw, (sym) | POY_GAP); } else PUT(pw, nw, POY_GAP); } while (0)
    int i = J.lasti, j = J.lastj;
    int mode = 4;  // 0 vertical, 1 horizontal, 2 diagonal, 3 align, 4 todo
    uint8_t *tile = s_tile[w];
    uint8_t *mv = s_moves[w];
    const unsigned lt = (1u << lane) - 1u;
    for (;;) {
        i = __shfl_sync(0xffffffffu, i, 0);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (i == 0 || j == 0) break;
        // tile: anti-diagonals a_hi-63 .. a_hi, direction-byte columns c0 .. c0+31
        const int a_hi = i + j, i_start = i, j_start = j;
        int dcur = j - i + k;
        dcur = dcur < 0 ? 0 : (dcur >= B ? B - 1 : dcur);
        int c0 = ((dcur >> 1) - 12) & ~15;
        if (c0 > stride - TB_COLS) c0 = stride - TB_COLS;
        if (c0 < 0) c0 = 0;
        for (int q = lane; q < TB_ROWS * 2; q += 32) {
            const int r = q >> 1, half = q & 1;
            const int a = a_hi - r;
            if (a >= 0 && c0 + half * 16 + 16 <= stride)
                *(uint4 *)(tile + r * TB_COLS + half * 16) = *(const uint4 *)(db + (size_t)a * stride + c0 + half * 16);
        }
        __syncwarp();
        // phase 1, lane 0: the walk proper -- direction bytes and the mode only, one move code per step (a tile holds at
        // most 64 steps: every step leaves its anti-diagonal)
        int nsteps = 0;
        if (lane == 0) {
            while (i != 0 && j != 0) {
                const int r = a_hi - (i + j);
                int d = j - i + k;
                d = d < 0 ? 0 : (d >= B ? B - 1 : d);
                const int cb = (d >> 1) - c0;
                if (r >= TB_ROWS || cb < 0 || cb >= TB_COLS) break;  // left the tile
                unsigned b = tile[r * TB_COLS + cb];
                if (tagged) {
                    // tagged format (cell_gf / cell_gen in band2.cu) -> the plain one: codes 0/1 are the two gap
                    // directions in priority order (H first for swapped operands), the END_* bits are stored
                    // inverted and bit 7 is HORIZONTAL_EQ_VERTICAL
                    unsigned t = b & 3u, nx = (b >> 2) & 3u;
                    if (swaped && t < 2u) t ^= 1u;
                    if (swaped && nx < 2u) nx ^= 1u;
                    const unsigned heqv = (b & 128u) ? 48u : 0u;
                    b = t | (nx << 2) | ((b & 112u) ^ 112u) | heqv;
                }
                if (mode == 4) mode = b & 3;
                mv[nsteps++] = (uint8_t)mode;
                if (mode == 0) { if (b & 16) mode = 4; --i; }
                else if (mode == 1) { if (b & 32) mode = 4; --j; }
                else if (mode == 2) { if (b & 64) mode = 4; --i; --j; }
                else {
                    // gap-free pairs store no ALIGN_TO code: it equals the todo code of the cell the walk moves to
                    mode = gf_todo ? 4 : (int)((b >> 2) & 3);
                    --i; --j;
                }
            }
        }
        nsteps = __shfl_sync(0xffffffffu, nsteps, 0);
        __syncwarp();
        // phase 2, all lanes: what each step emits (backtrace_aff writes right to left; one element per step into the two
        // rows and the median with gaps, the median proper only where a symbol survives)
        int ci = 0, cj = 0;      // rows / columns consumed by the steps before this chunk
        for (int t0 = 0; t0 < nsteps; t0 += 32) {
            const int st = t0 + lane;
            const bool valid = st < nsteps;
            const int m = valid ? mv[st] : 4;
            const unsigned bi = __ballot_sync(0xffffffffu, valid && m != 1), bj = __ballot_sync(0xffffffffu, valid && m != 0);
            int vm = 0, vw = POY_GAP, vi = POY_GAP, vj = POY_GAP;
            bool em = false;
            if (valid) {
                const int it = i_start - ci - __popc(bi & lt), jt = j_start - cj - __popc(bj & lt);
                const int ic = si[it], jc = sj[jt];
                if (m == 0) { em = !(ic & POY_GAP); vm = ic | POY_GAP; vw = em ? vm : POY_GAP; vi = ic; }
                else if (m == 1) { em = !(jc & POY_GAP); vm = jc | POY_GAP; vw = em ? vm : POY_GAP; vj = jc; }
                else if (m == 2) { vi = ic; vj = jc; }
                else { em = true; vm = cm->median32[((ic & POY_NOGAP) << 5) + (jc & POY_NOGAP)]; vw = vm; vi = ic; vj = jc; }
            }
            const unsigned bm = __ballot_sync(0xffffffffu, em);
            if (valid) {
                const int at = ni + st + 1;      // ni == nj == nw: one element per step
                if (pi) pi[-at] = (uint8_t)vi;
                if (pj) pj[-at] = (uint8_t)vj;
                if (pw) pw[-at] = (uint8_t)vw;
                if (em && pm) pm[-(nm + __popc(bm & lt) + 1)] = (uint8_t)vm;
            }
            if (bm) first_m = __shfl_sync(0xffffffffu, vm, 31 - __clz(bm));
            nm += __popc(bm);
            ci += __popc(bi); cj += __popc(bj);
        }
        ni += nsteps; nj += nsteps; nw += nsteps;
        __syncwarp();
    }
    // (the pointers of the serial tail below: where the steps above left off)
    if (pm) pm -= nm;
    if (pw) pw -= nw;
    if (pi) pi -= ni;
    if (pj) pj -= nj;
    if (lane == 0) {
        int ic = si[i], jc = sj[j];
        while (i != 0) {
            INDEL(ic);
            PUT(pi, ni, ic); PUT(pj, nj, POY_GAP);
            --i; ic = si[i];
        }
        while (j != 0) {
            INDEL(jc);
            PUT(pi, ni, POY_GAP); PUT(pj, nj, jc);
            --j; jc = sj[j];
        }
        PUT(pi, ni, POY_GAP); PUT(pj, nj, POY_GAP); PUT(pw, nw, POY_GAP);
        if (first_m != POY_GAP) PUT_M(POY_GAP);
        if (out_len) {
            out_len[4 * J.pair + 0] = nm; out_len[4 * J.pair + 1] = nw;
            out_len[4 * J.pair + 2] = ni; out_len[4 * J.pair + 3] = nj;
        }
    }
#undef PUT
#undef PUT_M
#undef INDEL
}

cudaError_t launch_band_generic(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                PairState *d_state, int *d_ebrow, uint8_t *d_dir, int *d_work, size_t work_stride,
                                int blocks) {
    if (njobs <= 0) return cudaSuccess;
    k_band_generic<<<blocks, 256, 0, ctx->stream>>>(cm->d, pool->d_rowp, pool->d_colp, pool->d_h0, d_jobs, njobs, d_state,
                                                     d_ebrow, d_dir, d_work, work_stride);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_traceback(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                             const uint8_t *d_done, const uint8_t *d_dir, const int64_t *d_out_off, uint8_t *d_median,
                             uint8_t *d_medianwg, uint8_t *d_resi, uint8_t *d_resj, int *d_out_len) {
    if (njobs <= 0) return cudaSuccess;
    k_traceback<<<(njobs + 3) / 4, 128, 0, ctx->stream>>>(cm->d, pool->d_data, d_jobs, njobs, d_done, d_dir, d_out_off, d_median,
                                                              d_medianwg, d_resi, d_resj, d_out_len);
    ctx->launches++;
    return cudaGetLastError();
}
