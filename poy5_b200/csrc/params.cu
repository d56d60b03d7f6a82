// Per-base gap parameters for every sequence of a pool, both roles.
//
// The reference recomputes these per alignment: the row side inside the row loop
// (src/algn.c:2058-2065, 2273-2283), the column side in a pre-loop over sj
// (src/algn.c:2026-2035, 2230-2236) after cm_precalc_4algn (src/cm.c:1334-1368)
// has gathered prepend[sj[j]] into `prec` row 0.  They depend only on the
// sequence and the cost model, so here they are computed once per pool upload
// (O(total bases), one warp per sequence) and every alignment that uses the
// sequence reads them back as one 16-byte load per base.
#include "common.cuh"

__device__ __forceinline__ int gap_opening_at(int idx, int prev, int cur, int go) {
    // HAS_GAP_OPENING, src/algn.c:1240-1253 (bitset alphabet branch)
    if (idx == 1 && (cur & POY_GAP)) return 0;
    if (idx > 1 && !(prev & POY_GAP) && (cur & POY_GAP)) return 0;
    return go;
}

__global__ void __launch_bounds__(128) k_params(const DevCM *__restrict__ cm, const uint8_t *__restrict__ data,
                                                const int64_t *__restrict__ off, int nseq, int4 *__restrict__ rowp,
                                                unsigned *__restrict__ rowpk, int4 *__restrict__ colp, int *__restrict__ h0, int *__restrict__ g0,
                                                uint8_t *__restrict__ gapfree) {
    __shared__ int s_prepend[32], s_gapext[32];
    if (threadIdx.x < 32) {
        s_prepend[threadIdx.x] = cm->prepend[threadIdx.x];
        s_gapext[threadIdx.x] = cm->gapext[threadIdx.x];
    }
    __syncthreads();
    const int go = cm->gap_open;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < nseq; s += gridDim.x * warps_per_block) {
        const int64_t base = off[s];
        const int len = (int)(off[s + 1] - base);
        int carry_h = 0, carry_g = 0, carry_r = 0;
        unsigned anygap = 0;
        for (int x0 = 0; x0 < len; x0 += 32) {
            const int x = x0 + lane;
            int hl = 0, ge_c = 0, ge_row = 0, code15 = 0;
            if (x < len) {
                const int code = data[base + x] & 31;
                const int prev = x > 0 ? (data[base + x - 1] & 31) : 0;
                const int curgap = code & POY_GAP, prevgap = prev & POY_GAP;
                const int flags = (code & POY_NOGAP) | (curgap ? PF_HASGAP : 0) | (prevgap ? PF_PREVGAP : 0);
                int4 r = make_int4(0, 0, 0, flags), c = make_int4(0, 0, 0, flags);
                if (x >= 1) {
                    const int gopen = gap_opening_at(x, prev, code, go);
                    const int ge_r = s_gapext[code];
                    ge_row = ge_r;
                    ge_c = s_prepend[code];
                    r.x = (x > 1 && prevgap && !curgap) ? gopen + ge_r : ge_r;
                    r.y = gopen + ge_r;
                    r.z = gopen;
                    hl = (prevgap && !curgap) ? gopen + ge_c : ge_c;  // in-loop value (row 0 of the banded fill)
                    c.x = (x == 1) ? ge_c : hl;                       // after the hext[1] overwrite
                    c.y = gopen + ge_c;
                    c.z = gopen;
                    if (curgap) anygap = 1;
                }
                code15 = code & POY_NOGAP;
                // bits 6-31 of the flags word: the window-entry fields of the band kernel (band2.cu, row_entry /
                // col_entry), ready to be masked out.  Surcharge class = {symbol has the gap bit, previous symbol has
                // it, gap opening is free here}; the column side keeps PF_PREVGAP as its bit 0 and also carries the
                // class the column has on the left border (previous symbol := the symbol itself).
                const int hg = curgap ? 1 : 0, pv = prevgap ? 1 : 0, goz = (r.z == 0) ? 1 : 0;
                r.w |= ((hg | (pv << 1) | (goz << 2)) << PF_ROW_CLASS_SHIFT) | (code15 << PF_ROW_GF_SHIFT) | (code15 << PF_ROW_TAB_SHIFT);
                c.w |= (hg << 6) | (goz << 7) | ((hg | (hg << 1) | (goz << 2)) << PF_COL_LB_SHIFT) | (code15 << PF_COL_TAB_SHIFT);
                rowp[base + x] = r;
                colp[base + x] = c;
            }
            // inclusive warp scans of hl, ge_c and the row-role gap extension
            int sh = hl, sg = ge_c, sr = ge_row;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int th = __shfl_up_sync(0xffffffffu, sh, d);
                int tg = __shfl_up_sync(0xffffffffu, sg, d);
                int tr = __shfl_up_sync(0xffffffffu, sr, d);
                if (lane >= d) { sh += th; sg += tg; sr += tr; }
            }
            if (x < len) {
                h0[base + x] = carry_h + sh;
                g0[base + x] = carry_g + sg;
                // gap-free cost kernel: R_i = sum of cost[s_r][gap], r = 1..i, in the low 28 bits (the domain check
                // keeps it below HIGH_NUM), row index of the cost table in the top 4
                rowpk[base + x] = ((unsigned)(carry_r + sr) & 0x0FFFFFFFu) | ((unsigned)code15 << 28);
            }
            carry_h += __shfl_sync(0xffffffffu, sh, 31);
            carry_g += __shfl_sync(0xffffffffu, sg, 31);
            carry_r += __shfl_sync(0xffffffffu, sr, 31);
        }
        anygap = __any_sync(0xffffffffu, anygap);
        if (lane == 0) gapfree[s] = anygap ? 0 : 1;
    }
}

// sequences [s0, s1) of the pool (the node store appends sequences and only the new ones need parameters)
cudaError_t launch_params(poy_ctx *ctx, const poy_cm *cm, poy_pool *pool, int s0, int s1) {
    const int n = s1 - s0;
    if (n <= 0) return cudaSuccess;
    int blocks = (n + 3) / 4;
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_params<<<blocks, 128, 0, ctx->stream>>>(cm->d, pool->d_data, pool->d_off + s0, n, pool->d_rowp, pool->d_rowpk, pool->d_colp,
                                               pool->d_h0, pool->d_g0, pool->d_gapfree + s0);
    ctx->launches++;
    return cudaGetLastError();
}

// Sequence.is_empty (src/sequence.ml:241-251: every symbol equals the gap code) and Sequence.count_gaps (symbols that
// carry the gap bit, seq_CAML_count, src/seq.c:644-669) per sequence, from the device bytes: x = empty, y = gap count.
__global__ void __launch_bounds__(128) k_seq_flags(const uint8_t *__restrict__ data, const int64_t *__restrict__ off, int nseq,
                                                   int2 *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < nseq; s += gridDim.x * warps_per_block) {
        const int64_t base = off[s];
        const int len = (int)(off[s + 1] - base);
        int nongap = 0, gapbit = 0;
        for (int x = lane; x < len; x += 32) {
            const int code = data[base + x];
            nongap += code != POY_GAP;
            gapbit += (code & POY_GAP) != 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            nongap += __shfl_xor_sync(0xffffffffu, nongap, d);
            gapbit += __shfl_xor_sync(0xffffffffu, gapbit, d);
        }
        if (lane == 0) flags[s] = make_int2(nongap == 0, gapbit);
    }
}

cudaError_t launch_seq_flags(poy_ctx *ctx, const poy_pool *pool, int s0, int s1, int2 *d_flags) {
    const int n = s1 - s0;
    if (n <= 0) return cudaSuccess;
    int blocks = (n + 3) / 4;
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_seq_flags<<<blocks, 128, 0, ctx->stream>>>(pool->d_data, pool->d_off + s0, n, d_flags);
    ctx->launches++;
    return cudaGetLastError();
}
