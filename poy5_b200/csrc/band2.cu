// Register-resident Ukkonen-band fill: k_band2<D, NW, WPB, GF, DIR>.
//
// Same algorithm and bit-exact results as the fallback in band_affine.cu (which keeps the documentation of the
// band geometry and the reference citations: algn_newkk_test_aff / algn_newkk_fill_a_row_aff / ASSIGN_MINIMUM /
// algn_fill_gapnum, src/algn.c:2186-2306, 2113-2180, 1936-1981, 126-176).  What changes is how the work is mapped
// to the machine:
//
//  * one CTA of NW warps per pair; thread t owns the D diagonals [t*D, (t+1)*D) and keeps the latest cell of each
//    in registers, so a CTA covers bands up to NW*32*D diagonals (D=8, NW=16: 4096).  Strip edges cross lanes by
//    one __shfl per sub-step and cross warps through 12-byte shared-memory mailboxes guarded by pairwise named
//    barriers (producer bar.arrive, consumer bar.sync); NW == 1 packs 8 independent warps into a CTA;
//  * the tie logic is branch free and mostly compare free: all values are carried times 256 and the low byte
//    holds tie-break tags, so the todo / ALIGN_TO codes and the END_* bits of the direction byte are the low bits
//    of the minima themselves (cell_gf / cell_gen);
//  * per-row / per-column gap parameters ride in sliding register windows (an anti-diagonal step moves every
//    thread one row down and one column right), so a cell does no global loads;
//  * the two unsigned-short gap counters travel packed in one register and are updated with the DPX 16x2 max
//    (__vimax3_u16x2); exact while len_i + len_j < 65535 (longer pairs take the fallback kernel);
//  * the 16x16 cost table is replicated once per shared-memory bank (32 KB) so the per-cell lookup is conflict
//    free whatever symbols the 32 lanes hold;
//  * the anti-diagonals a < delta+k+2, where some cell of the band still lies outside the matrix, run a
//    predicated prologue; afterwards every cell with d < B either is valid or lies beyond the last row/column,
//    where garbage can no longer reach a valid cell (the right-border rule cuts the only path), so the
//    steady-state loop carries no validity predicates at all;
//  * GF = both sequences free of gap-bit symbols: the EB state and everything that depends on gap bits
//    disappear (EB >= INF can neither win nor tie a minimum there), and the gap extensions move into the table
//    (shifted domain, see k_band2);
//  * DIR = false is the probe fill: same states and gap counters, no direction bytes.
#include <stdlib.h>
#include <type_traits>
#include <cooperative_groups.h>
#include "common.cuh"
namespace cg = cooperative_groups;

namespace {

template <int N, class F>
__device__ __forceinline__ void sfor(F &&f) {
    if constexpr (N > 0) {
        sfor<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

// window entries.
// Gap-free pairs: meta = shared-memory byte offset of the cost row / column.
// Pairs with gap-bit symbols: meta = (byte offset of the cost row / column) << 16 | (index of the surcharge class) << 5,
// so that ONE add of a row and a column entry yields both the table address (high half) and the address of the
// 32-byte surcharge record of this (row class, column class) in s_lut (bits 5-11); bits 12-14 of a column entry hold
// the class the column has when it sits on the left border (cell_gen).  ext already carries the END_* tag of the
// direction byte (DIR fills).
struct Ent { int ext, opn, meta; };

// The class / table-offset fields come ready-made in the flags word of the per-base parameters (k_params, PF_ROW_* /
// PF_COL_* in common.cuh): an entry's meta is one masked OR.  Surcharge class of a base: {symbol has the gap bit,
// previous symbol has it, gap opening is free here}; rows order the bits (has, prev, free), columns (prev, has, free).
template <bool GF, bool DIR>
__device__ __forceinline__ Ent row_entry(const int4 v, int swoff) {
    Ent e;
    if (GF) { e.ext = 0; e.opn = 0; e.meta = v.w & PF_ROW_GF_MASK; }
    else {
        e.ext = (v.x << 8) + (DIR ? 16 : 0); e.opn = (v.y << 8);     // x256: see cell_gf / cell_gen
        e.meta = (v.w & PF_ROW_MASK) | swoff;
    }
    return e;
}
template <bool GF, bool DIR>
__device__ __forceinline__ Ent col_entry(const int4 v, int lane) {
    Ent e;
    if (GF) { e.ext = 0; e.opn = 0; e.meta = (((v.w & 15) << 7) + (lane << 2)); }
    else {
        e.ext = (v.x << 8) + (DIR ? 32 : 0); e.opn = (v.y << 8);
        // bits 12-14: at the left border the reference's "previous column symbol" is the column symbol itself
        e.meta = (v.w & PF_COL_MASK) | (lane << 18);
    }
    return e;
}

// One band cell of a pair with gap-bit symbols (4 states).  On entry CB/EV/EH/EB/G hold the diagonal predecessor
// (i-1,j-1), on exit the new cell.  l* = (i,j-1), u* = (i-1,j).  EDGE adds the j == 0 handling of the prologue.
// Same value-times-256 + tag scheme as cell_gf below (read that comment first); here the ALIGN_TO_* code cannot be
// taken from the predecessor's todo code (the gap-bit surcharges differ), so it is a second tagged minimum, and
// the block state EB brings its own END_BLOCK tag.  Everything that depends on the gap bits of the two symbols and
// of their predecessors (src/algn.c:1376-1490: the block-diagonal costs, the three surcharges of ALIGN_TO_*) is ONE
// 32-byte record of s_lut, selected by the sum of the row and column entries: {v, h, block surcharge (with their
// ALIGN_TO tags), block opening, block extension (with the END_BLOCK tag)}; every minimum is a chain of
// add-then-min (VIADDMNMX).  Byte (tagged format): bits 0-1 todo code, bits 2-3 ALIGN_TO code (both: 0/1 = the two
// gap directions in priority order, 2 = block, 3 = align), bits 4/5/6 = NOT END_VERTICAL / END_HORIZONTAL /
// END_BLOCK, bit 7 = HORIZONTAL_EQ_VERTICAL.
#define INF256 (POY_INF << 8)
template <bool EDGE, bool DIR>
__device__ __forceinline__ unsigned cell_gen(int &CB, int &EV, int &EH, int &EB, unsigned &G, int lCB, int lEH, unsigned lG,
                                             int uCB, int uEV, unsigned uG, const Ent r, const Ent c, const char *s_tab,
                                             const char *s_lut, bool lb, bool rb, bool jpos, int tagV, int tagH) {
    // extend horizontal / vertical: ties take the opening (END_* flag = the tag carried by ext)
    int eH = __viaddmin_s32(lEH, c.ext, lCB + c.opn);
    int eV = __viaddmin_s32(uEV, r.ext, uCB + r.opn);
    if (lb) eH = INF256 + (DIR ? 32 : 0);
    if (rb) eV = INF256 + (DIR ? 16 : 0);
    const int nEH = DIR ? (eH & ~255) : eH, nEV = DIR ? (eV & ~255) : eV;
    int cm = c.meta;
    cm ^= (cm ^ (cm >> 7)) & (lb ? 0xE0 : 0);     // left border: the class of bits 12-14 replaces the one of bits 5-7
    const int sum = r.meta + cm;
    const int diag = *(const int *)(s_tab + ((unsigned)sum >> 16));
    const char *q = s_lut + (sum & 0xFE0);
    const int4 sur = *(const int4 *)q;            // x: EV -> CB, y: EH -> CB, z: EB -> CB, w: CB -> EB (block opening)
    const int dgx = *(const int *)(q + 16);       // EB -> EB (block extension)
    int eB = __viaddmin_s32(EB, dgx, CB + sur.w);
    // ALIGN_TO_*: tagged minimum over the four ways into CB (tags in bits 2-3)
    const int mk = __viaddmin_s32(EB, sur.z, __viaddmin_s32(EH, sur.y, __viaddmin_s32(EV, sur.x, DIR ? CB + 12 : CB)));
    int nCB = (DIR ? (mk & ~255) : mk) + diag;
    if (EDGE && !jpos) { nCB = INF256; eB = INF256 + (DIR ? 64 : 0); }
    const int nEB = DIR ? (eB & ~255) : eB;
    // final minimum, its todo code and its tie set
    const int k = DIR ? __viaddmin_s32(nEB, 2, __viaddmin_s32(nEH, tagH, __viaddmin_s32(nEV, tagV, nCB + 3)))
                      : min(__vimin3_s32(nEV, nEH, nCB), nEB);
    const bool fV = nEV <= k, fH = nEH <= k, fA = nCB <= k, fD = nEB <= k;
    // gap counters: component-wise max over the chosen predecessors (+1 on the side that gaps)
    const unsigned cD = (fA || fD) ? G : 0u;
    const unsigned cL = fH ? lG + 1u : 0u;
    const unsigned cU = fV ? uG + 0x10000u : 0u;
    G = __vimax3_u16x2(cD, cL, cU);
    unsigned w = (unsigned)(k | mk | eH) | (unsigned)(eV | eB);
    if (fH && fV) w |= 128u;
    CB = nCB; EV = nEV; EH = nEH; EB = nEB;
    return w;      // only the low byte is meaningful: pack_dir picks it
}

// One band cell of a gap-free pair.  All DP values are carried times 256 (so a finite value stays below 2^28 and
// an "infinite" one below 2^30), which leaves the low byte free for tags that turn tie-breaking into plain minima:
//   * EH = min(EH_left + 32, CB_left + GO): ties take the opening, so bit 5 of the winner says "the extension won
//     strictly" = no END_HORIZONTAL; likewise bit 4 for EV;
//   * the final choice is min3 of (EV + tagV, EH + tagH, CB + 3): its low two bits ARE the todo code, with the
//     V-before-H (or, for swapped operands, H-before-V) priority built in (tagV, tagH = 0, 1 or 1, 0); "X is in
//     the tie set" is X <= that minimum;
//   * ALIGN_TO_* of cell (i,j) are the DO_* flags of cell (i-1,j-1) (same three values, no gap-bit surcharges), so
//     nothing is stored for them: after an align step the traceback reads the next cell's todo code, and CB needs
//     only the carried minimum K of the diagonal predecessor.
// Byte (gap-free format, k_traceback decodes it when bit 6 of BandJob::swaped is set): bits 0-1 todo code (0/1 =
// the two gap directions in priority order, 3 = align), bit 4 / 5 = NOT END_VERTICAL / END_HORIZONTAL, bit 7 =
// HORIZONTAL_EQ_VERTICAL.
template <bool EDGE, bool DIR>
__device__ __forceinline__ unsigned cell_gf(int &CB, int &EV, int &EH, int &K, unsigned &G, int lCB, int lEH, unsigned lG,
                                            int uCB, int uEV, unsigned uG, int diag, int GO256, bool lb, bool rb, bool jpos,
                                            int tagV, int tagH) {
    // (a probe fill writes no byte: no END_* tags, nothing to clear)
    int eH = min(lEH + (DIR ? 32 : 0), lCB + GO256);
    int eV = min(uEV + (DIR ? 16 : 0), uCB + GO256);
    if (lb) eH = INF256 + (DIR ? 32 : 0);
    if (rb) eV = INF256 + (DIR ? 16 : 0);
    const int nEH = DIR ? (eH & ~255) : eH, nEV = DIR ? (eV & ~255) : eV;
    int nCB = (DIR ? (K & ~255) : K) + diag;
    if (EDGE && !jpos) nCB = INF256;
    const int k = DIR ? __vimin3_s32(nEV + tagV, nEH + tagH, nCB + 3) : __vimin3_s32(nEV, nEH, nCB);
    const bool fV = nEV <= k, fH = nEH <= k, fA = nCB <= k;
    // gap counters: component-wise max over the chosen predecessors (+1 on the side that gaps)
    const unsigned cD = fA ? G : 0u;
    const unsigned cL = fH ? lG + 1u : 0u;
    const unsigned cU = fV ? uG + 0x10000u : 0u;
    G = __vimax3_u16x2(cD, cL, cU);
    unsigned w = (unsigned)(k | eH | eV);
    if (fH && fV) w |= 128u;
    CB = nCB; EV = nEV; EH = nEH; K = k;
    return w;      // only the low byte is meaningful: pack_dir picks it
}

// Named-barrier producer / consumer pair (the PTX manual's bar.arrive / bar.sync idiom): the producing warp
// arrives without waiting, the consuming warp waits until both have reached the barrier.
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// 16-byte mailbox records between the CTAs of a cluster (generic addresses: the local copy is polled, the neighbour's
// copy is written through distributed shared memory); one vector access each, the sequence number travels in .w
__device__ __forceinline__ void st_mailbox(volatile int4 *p, int x, int y, int z, int w) {
    asm volatile("st.volatile.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ int4 ld_mailbox(const volatile int4 *p) {
    int4 v;
    asm volatile("ld.volatile.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// bounded poll (a lost hand-over must end as a wrong result that the parity tests catch, never as a hung GPU)
__device__ __forceinline__ int4 wait_mailbox(const volatile int4 *p, int expect, bool &failed) {
    int4 v = ld_mailbox(p);
    if (failed) return v;
    for (int spin = 0; v.w != expect && spin < (1 << 16); ++spin) v = ld_mailbox(p);
    failed = v.w != expect;           // give up for the rest of this pair: ~1 ms lost once, not per sub-step
    return v;
}

// Progress words of speculative fills (global memory, one per fill): rows completed by the thread that owns the two
// leftmost diagonals, SPEC_DONE at the end, SPEC_FAILED if the fill gave up.  Polls are bounded: a predecessor that is
// not running (kernels serialised by a tool, a GPU shared with another process) must end as a fill the host repeats,
// never as a hung GPU.
__device__ __forceinline__ void spec_publish(int *p, int v) {
    __threadfence();
    *(volatile int *)p = v;
}
// Two bounds.  BEFORE a fill starts (max_spin = 2^18, ~0.1 s) giving up is harmless: nothing was written, the host repeats
// the doubling.  IN MID-RUN the fill has already stored into the pair's stale row, so giving up would leave a state the
// plain schedule never produces; there the predecessor is known to be running (it reported progress), it can only be
// waiting for ITS predecessor, and the head of a chain waits for nobody -- the wait always ends, and the bound (2^26 polls,
// tens of seconds) is only a net under a bug.
__device__ __forceinline__ int spec_wait(const int *p, int need, int max_spin) {     // returns the last value read: >= need unless it gave up
    int v = *(const volatile int *)p;
    for (int spin = 0; v >= 0 && v < need && spin < max_spin; ++spin) { __nanosleep(200); v = *(const volatile int *)p; }
    __threadfence();
    return v;
}
// what a fill publishes when it is over: status 1 = gave up, 2 = abandoned (an earlier doubling stopped), else complete
__device__ __forceinline__ int spec_final(int status, const PairState *st) {
    if (status == 1) return SPEC_FAILED;
    if (status == 2) return SPEC_ABORT;
    return *(const volatile int *)&st->done ? SPEC_STOP : SPEC_DONE;
}

// low bytes of H words -> one little-endian word (PRMT: three byte permutes for four cells)
template <int H>
__device__ __forceinline__ unsigned pack_dir(const unsigned (&b)[H]) {
    if constexpr (H == 1) return b[0] & 0xFFu;
    else if constexpr (H == 2) return __byte_perm(b[0], b[1], 0x0040) & 0xFFFFu;
    else return __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
}

template <int H>
__device__ __forceinline__ void store_dir(uint8_t *p, unsigned packed) {
    static_assert(H == 1 || H == 2 || H == 4, "one to four direction bytes per thread and sub-step");
    if (H == 1) *p = (uint8_t)packed;
    else if (H == 2) *(uint16_t *)p = (uint16_t)packed;
    else *(uint32_t *)p = packed;
}

}  // namespace

// NW warps cooperate on one pair.  NW == 1: a CTA holds WPB independent warps (each with its own
// pair) that only share the replicated cost table.
// DIR = false is the "probe" fill: same states, gap counters and stale-row side effects, but no direction bytes
// (the whole byte assembly is dead code then).  The host uses it for fills it expects not to be the last one.
// resident CTAs the register allocation aims at: the replicated table (32 KB) allows 6 per SM; the 4-warp class of
// gap-free pairs is the one where a few registers decide between 4 and 5
template <int NW, bool GFK> struct MinBlocks {
    // pairs with gap-bit symbols: 128 registers per thread (no spills since the surcharge records), i.e. two resident
    // CTAs of 8 warps / four of 4 warps per SM instead of one / three
    static constexpr int v = GFK ? (NW == 4 ? 5 : 1)
                                 : (NW == 8 ? 2 : NW == 4 ? 4 : NW == 3 ? 5 : NW == 5 ? 3 : NW == 6 ? 2 : NW == 2 ? 8 : NW == 1 ? 2 : 1);
};

// CL > 1: one pair per thread-block CLUSTER of CL CTAs of NW warps each ("cooperative tiled bands" for wide bands in
// latency-bound rounds): warp w of CTA c owns the diagonals of global warp c * NW + w.  Inside a CTA the warps hand their
// strip edges over exactly as before (shared-memory mailboxes + pairwise named barriers); between the last warp of CTA c
// and the first warp of CTA c + 1 the 16-byte mailbox record {CB, EV|EH, G, sequence number} is written into the
// NEIGHBOUR's shared memory (distributed shared memory, one vector store) and the consumer polls its own copy for the
// sequence number it expects -- no cluster-wide barrier inside a fill (a cluster.sync costs ~380 cycles and flushes
// L1; a DSMEM store ~215).  Two slots per direction (sequence parity): a producer cannot get more than one hand-over
// ahead of its consumer because it needs the consumer's reply for its own next sub-step.
template <int D, int NW, int WPB, bool GFK, bool DIR, int CL = 1>
__global__ void __launch_bounds__(WPB * 32, CL > 1 ? 1 : MinBlocks<NW, GFK>::v)
k_band2(const DevCM *__restrict__ cm, const int4 *__restrict__ rowp, const int4 *__restrict__ colp,
        const int *__restrict__ h0v, const int *__restrict__ g0v, const unsigned *__restrict__ rowpk,
        const BandJob *__restrict__ jobs, int njobs, int *counter, PairState *state, int *ebrow, uint8_t *dir, int *prog) {
    constexpr int H = D / 2;
    static_assert(CL == 1 || (NW > 1 && NW <= 8), "cluster shapes use the pairwise-barrier path inside each CTA");
    __shared__ __align__(16) int4 s_cx_left[2];    // CL > 1: written by the CTA to the left (its last warp's slot D-1)
    __shared__ __align__(16) int4 s_cx_right[2];   // CL > 1: written by the CTA to the right (its first warp's slot 0)
    // one array, so that both tables are addressed from one base register:
    //   s_tab_i: cost16 replicated per bank, entry e of lane l at [e*32 + l]
    //   s_lut_i: surcharge records of cell_gen, [swaped][row class][column class] x 32 B
    __shared__ __align__(16) int s_mem_i[256 * 32 + (GFK ? 8 : 128 * 8)];
    int *const s_tab_i = s_mem_i, *const s_lut_i = s_mem_i + 256 * 32;
    __shared__ int s_job;
    __shared__ int s_spec;                         // speculative fills: 1 = this fill gave up on its predecessor, 2 = it was abandoned (rank 0's copy counts)
    __shared__ int s_astop;                        // ... last anti-diagonal to compute once the fill is abandoned (every CTA's own copy)
    __shared__ __align__(16) int4 s_xe[NW];   // slot-0 state of lane 0 of every warp (read by the warp to its left): one 16-byte access
    __shared__ __align__(16) int4 s_xo[NW];   // slot-(D-1) state of lane 31 of every warp (read by the warp to its right)
    static_assert(NW == 1 || WPB == NW, "cooperating warps fill the whole CTA");
    int crank = 0;
    if constexpr (CL > 1) crank = (int)cg::this_cluster().block_rank();
    const int lane = threadIdx.x & 31, lw = (NW == 1) ? 0 : (threadIdx.x >> 5);      // lw: warp within the CTA
    const int warp = lw + crank * NW;                                                // warp within the group that owns the pair
    const int tid = (NW == 1) ? lane : (int)threadIdx.x + crank * NW * 32;           // thread index within that group
    volatile int4 *rem_left_cx_right = nullptr, *rem_right_cx_left = nullptr;        // the neighbours' mailboxes for this CTA
    const int *rem_job = &s_job;
    volatile int *rem_spec = &s_spec;
    int pub_slot = -1;                               // progress slot of the fill just completed (published at the next CTA-wide sync)
    int pub_stat = 0;                                // NW == 1: status of that fill (see spec_final)
    if constexpr (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (threadIdx.x < 2) { s_cx_left[threadIdx.x] = make_int4(0, 0, 0, -1); s_cx_right[threadIdx.x] = make_int4(0, 0, 0, -1); }
        if (crank > 0) rem_left_cx_right = (volatile int4 *)cluster.map_shared_rank(&s_cx_right[0], crank - 1);
        if (crank + 1 < CL) rem_right_cx_left = (volatile int4 *)cluster.map_shared_rank(&s_cx_left[0], crank + 1);
        rem_job = cluster.map_shared_rank(&s_job, 0);
        rem_spec = (volatile int *)cluster.map_shared_rank(&s_spec, 0);
    }
    // Gap-free launches work in a shifted domain: every state of cell (i,j) is carried minus S_j + R_i (the sums
    // of the column / row gap extensions up to j / i), which moves the "+ ge" of EH and EV into the table as
    // cost[a][b] - prepend[b] - cost[a][gap].  All comparisons of a cell are between states of the same shift, so
    // the direction bytes and gap counters do not change; the cost is shifted back when it is stored.
    for (int x = threadIdx.x; x < 256 * 32; x += WPB * 32) {
        const int e = x >> 5;
        s_tab_i[x] = (GFK ? cm->cost16[e] - cm->prepend[e & 15] - cm->gapext[e >> 4] : cm->cost16[e]) * 256;   // x256: see cell_gf
    }
    const int GO = cm->gap_open;
    const int GO256 = GO << 8;
    if (!GFK)
        for (int x = threadIdx.x; x < 128; x += WPB * 32) {
            const int sw = x >> 6, rc = (x >> 3) & 7, cc = x & 7;
            const bool hg_i = rc & 1, pv_i = rc & 2, goz_i = rc & 4, pv_j = cc & 1, hg_j = cc & 2, goz_j = cc & 4;
            const int tV = sw ? 1 : 0, tH = sw ? 0 : 1;
            const bool both = hg_i && hg_j, clean = !pv_i && !pv_j;
            int *e = s_lut_i + x * 8;
            e[0] = ((hg_i && !goz_j) ? GO256 : 0) + (DIR ? 4 * tV : 0);     // EV -> CB: + go_j if the row symbol has the gap bit
            e[1] = ((hg_j && !goz_i) ? GO256 : 0) + (DIR ? 4 * tH : 0);     // EH -> CB
            e[2] = ((goz_i && goz_j) ? 0 : GO256) + (DIR ? 8 : 0);          // EB -> CB: + max(go_i, go_j)
            e[3] = both ? (clean ? 0 : 2 * GO256) : INF256;                 // CB -> EB
            e[4] = (both ? 0 : INF256) + (DIR ? 64 : 0);                    // EB -> EB
            e[5] = e[6] = e[7] = 0;
        }
    __syncthreads();
    const char *s_tab = (const char *)s_tab_i;
    const char *s_lut = (const char *)s_lut_i;

    for (;;) {
        int job;
        // a speculative fill that is complete says so once every thread's stale-row stores are out (fence, CTA-wide sync)
        if (pub_slot >= 0) __threadfence();
        if (NW == 1) {
            job = 0;
            __syncwarp();
            if (lane == 0) {
                if (pub_slot >= 0) spec_publish(prog + pub_slot, spec_final(pub_stat, state + pub_slot));
                job = atomicAdd(counter, 1);
            }
            job = __shfl_sync(0xffffffffu, job, 0);
        } else if constexpr (CL > 1) {
            cg::this_cluster().sync();           // every CTA is done with the previous pair (and with its mailboxes)
            if (tid == 0) {
                if (pub_slot >= 0) spec_publish(prog + pub_slot, spec_final(s_spec, state + pub_slot));
                s_job = atomicAdd(counter, 1); s_spec = 0;
            }
            if (threadIdx.x == 0) s_astop = 0x7fffffff;
            cg::this_cluster().sync();
            job = *rem_job;
        } else {
            __syncthreads();
            if (tid == 0) {
                if (pub_slot >= 0) spec_publish(prog + pub_slot, spec_final(s_spec, state + pub_slot));
                s_job = atomicAdd(counter, 1); s_spec = 0; s_astop = 0x7fffffff;
            }
            __syncthreads();
            job = s_job;
        }
        pub_slot = -1; pub_stat = 0;
        if (job >= njobs) break;
        const BandJob J = jobs[job];
        if (J.lasti == 0) continue;
        // speculative fill: wait until the previous doubling's fill has left the stale EB entries this one starts from
        const bool spec = (J.swaped & 128) != 0;
        if (spec && J.dep >= 0) {
            int st0 = 0;                      // 1 = gave up, 2 = an earlier doubling stopped: nothing to do
            auto judge = [&](int v) { return v >= SPEC_STOP ? 2 : (v < J.need ? 1 : 0); };
            if (NW == 1) {
                if (lane == 0) st0 = judge(spec_wait(prog + J.dep, J.need, 1 << 18));
                st0 = __shfl_sync(0xffffffffu, st0, 0);
            } else if constexpr (CL > 1) {
                if (tid == 0) s_spec = judge(spec_wait(prog + J.dep, J.need, 1 << 18));
                cg::this_cluster().sync();
                st0 = *rem_spec;
            } else {
                if (tid == 0) s_spec = judge(spec_wait(prog + J.dep, J.need, 1 << 18));
                __syncthreads();
                st0 = s_spec;
            }
            if (st0 != 0) {                   // nothing was touched (1: the host runs this doubling again, on its own)
                if (tid == 0) spec_publish(prog + J.pair, st0 == 1 ? SPEC_FAILED : SPEC_ABORT);
                continue;
            }
        }
        if (spec) pub_slot = J.pair;
        // gap-free pairs (J.swaped bit 2, all jobs of a GFK launch) take the 3-state code path
        auto run = [&](auto gf_c) {
        constexpr bool GF = decltype(gf_c)::value;
        const int lasti = J.lasti, lastj = J.lastj, k = J.k;
        const bool swaped = (J.swaped & 1) != 0;
        const int delta = lastj - lasti, B = delta + 2 * k + 1;
        const int4 *rp = rowp + J.off_i;
        const int4 *cp = colp + J.off_j;
        asm("" : "+l"(rp)); asm("" : "+l"(cp));   // keep the per-pair bases whole: a window load is base + 16 * index (one IMAD.WIDE)
        const int *h0 = h0v + J.off_j;
        const int *g0 = g0v + J.off_j;
        int *eb = ebrow + J.eb_off;
        PairState *st = state + J.pair;
        uint8_t *dbase = dir + J.dir_off;
        const int stride = J.stride;
        const int d0 = tid * D;
        const int eh00 = spec ? J.eh00 : st->eh00;
        const int rbslot = (B - 1) - d0;  // slot holding the right border, if 0 <= rbslot < D

        int CB[D], EV[D], EH[D], EB[D];
        int K[D];                         // gap-free pairs: tagged minimum of the diagonal predecessor (cell_gf)
        unsigned G[D];
        const int tagV = swaped ? 1 : 0, tagH = swaped ? 0 : 1;
        sfor<D>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int d = d0 + u, j0 = d - k;
            if (d < B && j0 >= 0 && j0 <= lastj) {  // row 0 (src/algn.c:2222-2247)
                if (GF) {
                    CB[u] = (h0[j0] - g0[j0]) << 8;
                    EH[u] = j0 == 0 ? (eh00 << 8) : CB[u];
                    EV[u] = INF256;
                    K[u] = min(CB[u], EH[u]);
                } else {
                    CB[u] = h0[j0] << 8;
                    EH[u] = j0 == 0 ? (eh00 << 8) : CB[u];
                    EV[u] = INF256;
                }
                EB[u] = GF ? INF256 : (__ldcg(eb + j0) << 8);  // the stale EB row is kept unscaled in memory (L2: another SM may have written it)
                G[u] = (unsigned)j0 & 0xFFFFu;
            } else {
                CB[u] = EV[u] = EH[u] = EB[u] = K[u] = INF256; G[u] = 0u;
            }
        });

        // sliding windows: R[h] = row i0-h, C[h] = column j0+h
        int a = k & 1;
        int i0 = (a - d0 + k) >> 1, j0 = a - i0;
        Ent R[H], C[H + 1];
        const int swoff = swaped ? 2048 : 0;     // second half of s_lut: the ALIGN_TO tags of swapped operands
        auto load_row = [&](int i) { i = i < 0 ? 0 : (i > lasti ? lasti : i); return row_entry<GF, DIR>(__ldg(rp + i), swoff); };
        auto load_col = [&](int j) { j = j < 0 ? 0 : (j > lastj ? lastj : j); return col_entry<GF, DIR>(__ldg(cp + j), lane); };
        sfor<H>([&](auto hc) { constexpr int h = decltype(hc)::value; R[h] = load_row(i0 - h); });
        sfor<H + 1>([&](auto hc) { constexpr int h = decltype(hc)::value; C[h] = load_col(j0 + h); });

        // Neighbouring warps synchronise pairwise (P2P) instead of CTA wide: warp w waits for the odd sub-step of
        // warp w-1 on barrier O(w-1) and for the even sub-step of warp w+1 on barrier E(w+1); the producer side only
        // arrives.  Program order makes every arrive / wait pair up and keeps the 12-byte mailboxes race free
        // (DESIGN.md section 5).  Warps wholly right of the band take no part.  16 warps would need 30 barrier
        // ids, so that class keeps __syncthreads.
        constexpr bool P2P = (NW > 1 && NW <= 8);
        const bool warp_in_band = (NW == 1) || (warp * 32 * D < B);
        const bool has_left = P2P && lw > 0;
        const bool has_right = P2P && lw + 1 < NW && (warp + 1) * 32 * D < B;
        // cluster shapes: the neighbour across the CTA boundary
        const bool rem_left = CL > 1 && lw == 0 && crank > 0 && warp_in_band;
        const bool rem_right = CL > 1 && lw == NW - 1 && crank + 1 < CL && (warp + 1) * 32 * D < B;
        const int seqbase = (job + 1) << 16;                            // hand-over sequence numbers of this pair
        int it = 0;
        bool mb_failed = false;
        const int bar_E_mine = lw, bar_E_right = lw + 1;                // E(w) = id w, w = 1 .. NW-1
        const int bar_O_mine = NW + lw, bar_O_left = NW + lw - 1;       // O(w) = id NW + w, w = 0 .. NW-2
        if (NW > 1) {
            if (lane == 0) s_xe[lw] = make_int4(CB[0], EV[0], (int)G[0], 0);
            if (lane == 31) s_xo[lw] = make_int4(CB[D - 1], EH[D - 1], (int)G[D - 1], 0x7fffffff);
            if (CL > 1 && rem_right && lane == 31) st_mailbox(rem_right_cx_left, CB[D - 1], EH[D - 1], (int)G[D - 1], seqbase);
            __syncthreads();
            if (P2P && warp_in_band && has_right) pair_arrive(bar_O_mine);   // row 0 stands in for "odd sub-step -1"
        } else {
            __syncwarp();
        }

        const int a_end = lasti + lastj;
        int mid_stat = 0;                               // speculative fill, thread 0 (NW == 1: its warp): 1 = gave up on the predecessor in mid-run, 2 = abandoned
        int astop = 0x7fffffff;                         // ... last anti-diagonal to compute once the fill is abandoned
        uint8_t *dptr = dbase + (size_t)a * stride + tid * H;   // direction bytes of this thread on anti-diagonal a (advanced per iteration)
        int a_main = delta + k + 2;                     // first anti-diagonal whose band cells all have i >= 1, j >= 1
        if ((a_main ^ a) & 1) ++a_main;
        const int istar = lasti & ~1;                   // last even row: source of the stale EB row

        // one loop iteration = anti-diagonals a (even diagonals) and a+1 (odd diagonals)
        auto iteration = [&](auto edge_c) {
            constexpr bool EDGE = decltype(edge_c)::value;
            if (CL > 1 && spec) astop = min(astop, *(volatile int *)&s_astop);     // (inside a CTA the bound travels with the mailboxes)
            // next window entries: issued now, consumed after both sub-steps (hides the L1/L2 latency)
            int4 nrow = make_int4(0, 0, 0, 0), ncol = make_int4(0, 0, 0, 0);
            if (warp_in_band) {
                int ri = i0 + 1, cj = j0 + 1 + H;
                if (EDGE) {
                    ri = ri < 0 ? 0 : (ri > lasti ? lasti : ri);
                    cj = cj < 0 ? 0 : (cj > lastj ? lastj : cj);
                } else {
                    // past a_main every band cell has i >= 1, j >= 1: only the slots right of the band can still sit
                    // on a negative row, and what they load is never used -- one unsigned minimum keeps it in range
                    ri = (int)min((unsigned)ri, (unsigned)lasti);
                    cj = (int)min((unsigned)cj, (unsigned)lastj);
                }
                nrow = __ldg(rp + ri); ncol = __ldg(cp + cj);
            }
            // the stale EB row is written by the cells of even rows on the two leftmost diagonals
            // and of row istar: one test per iteration instead of one per cell
            const bool stale_hit = !GF && ((tid == 0 && !(i0 & 1)) || (unsigned)(i0 - istar) < (unsigned)H);
            // ---- even diagonals ----
            {
                if (has_left) pair_wait(bar_O_left);
                int sCB = __shfl_up_sync(0xffffffffu, CB[D - 1], 1);
                int sEH = __shfl_up_sync(0xffffffffu, EH[D - 1], 1);
                unsigned sG = __shfl_up_sync(0xffffffffu, G[D - 1], 1);
                if (NW > 1 && lane == 0 && lw > 0) { const int4 v = s_xo[lw - 1]; sCB = v.x; sEH = v.y; sG = (unsigned)v.z; }
                // a speculative fill that is being abandoned: the last anti-diagonal to compute travels from warp to warp
                // with the hand-over (one warp per iteration, the bound is set 32 iterations ahead)
                if (NW > 1 && spec && lw > 0) astop = min(astop, s_xo[lw - 1].w);
                if (CL > 1 && rem_left && lane == 0) {
                    const int4 v = wait_mailbox(&s_cx_left[it & 1], seqbase + it, mb_failed);
                    sCB = v.x; sEH = v.y; sG = (unsigned)v.z;
                }
                unsigned bw[H];
                sfor<H>([&](auto hc) { bw[decltype(hc)::value] = 0u; });
                if (warp_in_band)
                sfor<H>([&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr int u = 2 * h;
                    const int i = i0 - h, j = j0 + h, d = d0 + u;
                    bool valid = true;
                    if (EDGE) valid = (d < B) && (i >= 1) && (i <= lasti) && (j >= 0) && (j <= lastj);
                    if (valid) {
                        int lCB, lEH; unsigned lG;
                        if constexpr (u == 0) { lCB = sCB; lEH = sEH; lG = sG; }
                        else { lCB = CB[u > 0 ? u - 1 : 0]; lEH = EH[u > 0 ? u - 1 : 0]; lG = G[u > 0 ? u - 1 : 0]; }
                        bool lb = false;
                        if constexpr (u == 0) lb = (tid == 0);
                        if (EDGE) lb = lb || (j == 0);
                        unsigned b;
                        if constexpr (GF)
                            b = cell_gf<EDGE, DIR>(CB[u], EV[u], EH[u], K[u], G[u], lCB, lEH, lG, CB[u + 1], EV[u + 1], G[u + 1],
                                              *(const int *)(s_tab + R[h].meta + C[h].meta), GO256, lb, u == rbslot, j > 0, tagV, tagH);
                        else
                            b = cell_gen<EDGE, DIR>(CB[u], EV[u], EH[u], EB[u], G[u], lCB, lEH, lG, CB[u + 1], EV[u + 1], G[u + 1],
                                                    R[h], C[h], s_tab, s_lut, lb, u == rbslot, j > 0, tagV, tagH);
                        bw[h] = b;
                    }
                });
                if (NW > 1) {
                    if (lane == 0) s_xe[lw] = make_int4(CB[0], EV[0], (int)G[0], 0);
                    if (CL > 1 && rem_left && lane == 0) st_mailbox(rem_left_cx_right + (it & 1), CB[0], EV[0], (int)G[0], seqbase + it + 1);
                    if (P2P) { if (has_left) pair_arrive(bar_E_mine); }
                    else __syncthreads();
                }
                // (after the hand-over: the neighbour does not wait for this store)
                if (DIR && warp_in_band) store_dir<H>(dptr, pack_dir<H>(bw));
            }
            // ---- odd diagonals ----
            {
                if (has_right) pair_wait(bar_E_right);
                int sCB = __shfl_down_sync(0xffffffffu, CB[0], 1);
                int sEV = __shfl_down_sync(0xffffffffu, EV[0], 1);
                unsigned sG = __shfl_down_sync(0xffffffffu, G[0], 1);
                if (NW > 1 && lane == 31 && lw < NW - 1) { const int4 v = s_xe[lw + 1]; sCB = v.x; sEV = v.y; sG = (unsigned)v.z; }
                if (CL > 1 && rem_right && lane == 31) {
                    const int4 v = wait_mailbox(&s_cx_right[it & 1], seqbase + it + 1, mb_failed);
                    sCB = v.x; sEV = v.y; sG = (unsigned)v.z;
                }
                unsigned bw[H];
                sfor<H>([&](auto hc) { bw[decltype(hc)::value] = 0u; });
                if (warp_in_band)
                sfor<H>([&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr int u = 2 * h + 1;
                    const int i = i0 - h, j = j0 + h + 1, d = d0 + u;
                    bool valid = true;
                    if (EDGE) valid = (d < B) && (i >= 1) && (i <= lasti) && (j >= 0) && (j <= lastj);
                    if (valid) {
                        int uCB, uEV; unsigned uG;
                        if constexpr (u == D - 1) { uCB = sCB; uEV = sEV; uG = sG; }
                        else { constexpr int uu = u < D - 1 ? u + 1 : u; uCB = CB[uu]; uEV = EV[uu]; uG = G[uu]; }
                        bool lb = false;
                        if (EDGE) lb = (j == 0);
                        unsigned b;
                        if constexpr (GF)
                            b = cell_gf<EDGE, DIR>(CB[u], EV[u], EH[u], K[u], G[u], CB[u - 1], EH[u - 1], G[u - 1], uCB, uEV, uG,
                                              *(const int *)(s_tab + R[h].meta + C[h + 1].meta), GO256, lb, u == rbslot, j > 0, tagV, tagH);
                        else
                            b = cell_gen<EDGE, DIR>(CB[u], EV[u], EH[u], EB[u], G[u], CB[u - 1], EH[u - 1], G[u - 1], uCB, uEV, uG,
                                                    R[h], C[h + 1], s_tab, s_lut, lb, u == rbslot, j > 0, tagV, tagH);
                        bw[h] = b;
                    }
                });
                if (NW > 1) {
                    if (lane == 31) s_xo[lw] = make_int4(CB[D - 1], EH[D - 1], (int)G[D - 1], astop);
                    if (CL > 1 && rem_right && lane == 31)
                        st_mailbox(rem_right_cx_left + ((it + 1) & 1), CB[D - 1], EH[D - 1], (int)G[D - 1], seqbase + it + 1);
                    if (P2P) { if (has_right) pair_arrive(bar_O_mine); }
                    else __syncthreads();
                }
                if (DIR && warp_in_band) store_dir<H>(dptr + stride, pack_dir<H>(bw));
            }
            // the stale EB row: every slot was written once in this iteration (even slots in the first sub-step), so one
            // rare block serves both; it keeps the address arithmetic of the stale row out of the cells
            if (stale_hit && warp_in_band && !(spec && (NW == 1 ? mid_stat : *rem_spec) != 0))
            sfor<D>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                const int i = i0 - u / 2, j = j0 + (u + 1) / 2, d = d0 + u;
                if (!(i & 1) && (d <= 1 || i == istar) && d < B && i >= 1 && i <= lasti && j >= 0 && j <= lastj) eb[j] = EB[u] >> 8;
            });
            // speculative fills, every 16th iteration, the thread of the two leftmost diagonals (the one furthest down
            // the matrix): say how far the stale row is written, and stay behind the previous doubling's fill so that
            // the stores to the shared row land in the order of the doublings
            if (spec && (it & 15) == 15 && !GF) {
                if (tid == 0 && mid_stat == 0) {
                    spec_publish(prog + J.pair, i0);
                    if (J.dep >= 0) {
                        const int want = (i0 + 18 >= istar) ? SPEC_DONE : i0 + 18;
                        const int v = spec_wait(prog + J.dep, want, 1 << 26);
                        if (v >= SPEC_STOP) {
                            // an earlier doubling is the result: every warp leaves the loop after the same anti-diagonal,
                            // far enough ahead that all of them have seen the new bound by then
                            mid_stat = 2;
                            astop = a + (CL > 1 ? 192 : 64);
                            if (NW > 1) *rem_spec = 2;
                            if constexpr (CL > 1) { for (int r = 1; r < CL; ++r) *(volatile int *)cg::this_cluster().map_shared_rank(&s_astop, r) = astop; }
                        } else if (v < want) {
                            mid_stat = 1;
                            if (NW > 1) *rem_spec = 1;
                        }
                    }
                }
                // (the warp of thread 0 agrees on what its lane 0 decided; the other warps learn the bound from the mailboxes)
                if (NW == 1 || (lw == 0 && crank == 0)) { mid_stat = __shfl_sync(0xffffffffu, mid_stat, 0); astop = __shfl_sync(0xffffffffu, astop, 0); }
            }
            // the window entries prefetched at the top are first touched here: the load had the whole iteration to land
            asm volatile("" : "+r"(nrow.x), "+r"(nrow.y), "+r"(nrow.w), "+r"(ncol.x), "+r"(ncol.y), "+r"(ncol.w));
            // ---- slide the windows one row down / one column right ----
            sfor<H - 1>([&](auto hc) { constexpr int h = H - 1 - decltype(hc)::value; R[h] = R[h - 1]; });
            sfor<H>([&](auto hc) { constexpr int h = decltype(hc)::value; C[h] = C[h + 1]; });
            ++i0; ++j0; ++it;
            dptr += 2 * (size_t)stride;
            R[0] = row_entry<GF, DIR>(nrow, swoff);
            C[H] = col_entry<GF, DIR>(ncol, lane);
        };

        if (!P2P || warp_in_band) {
            for (; a <= a_end && a < a_main && a <= astop; a += 2) iteration(std::true_type{});
            for (; a <= a_end && a <= astop; a += 2) iteration(std::false_type{});
            if (has_left) pair_wait(bar_O_left);   // drains the last arrive of the left neighbour
        }

        // result: the cell (lasti, lastj) is the latest cell of diagonal delta + k
        const int dstar = delta + k;
        if (tid == dstar / D) {
            sfor<D>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                if (u == dstar % D) {
                    int fin = __vimin3_s32(EH[u], EV[u], CB[u]);
                    if (!GF) fin = min(fin, EB[u]);
                    fin >>= 8;
                    if (GF) fin += g0[lastj] + (int)(rowpk[J.off_i + lasti] & 0x0FFFFFFFu);
                    st->cost = fin;
                    const int gapnum = max((int)(G[u] & 0xFFFFu), (int)(G[u] >> 16));
                    st->gapnum = gapnum;
                    if (spec) {     // the stop rule (algn_newkk_test_aff; k_band_finish evaluates the same on its own): later doublings look at it
                        const int p = (J.T - delta) / 2, newp = (2 * J.T - delta) / 2;
                        st->done = (gapnum < p || newp - lastj + 1 >= 0) ? 1 : 0;
                    }
                }
            });
        }
        if (tid == 0 && min(k, lasti) >= 2) st->eh00 = POY_INF;  // an even row wrote EH[.][0] = INF into row buffer 0
        if (NW == 1) pub_stat = mid_stat;
        };
        // one kind of pair per launch (GFK): the 3-state path needs ~35 fewer registers, i.e. one more resident CTA
        if constexpr (GFK) run(std::true_type{}); else run(std::false_type{});
    }
    if constexpr (CL > 1) cg::this_cluster().sync();     // no CTA may exit while a neighbour can still touch its shared memory
}

template <int D, int NW>
static cudaError_t launch_one(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                              bool gapfree, bool probe, int *d_counter, PairState *d_state, int *d_ebrow, uint8_t *d_dir, int *d_prog) {
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
    if (e != cudaSuccess) return e;
    constexpr int WPB = NW == 1 ? 8 : NW;
    const int groups_per_block = WPB / NW;
    int blocks = (njobs + groups_per_block - 1) / groups_per_block;
    const int cap = ctx->sm_count * 6;   // more CTAs than can be resident just queue behind the persistent ones
    if (blocks > cap) blocks = cap;
#define LAUNCH(GFV, DIRV) k_band2<D, NW, WPB, GFV, DIRV><<<blocks, WPB * 32, 0, ctx->stream>>>(cm->d, pool->d_rowp, pool->d_colp, \
        pool->d_h0, pool->d_g0, pool->d_rowpk, d_jobs, njobs, d_counter, d_state, d_ebrow, d_dir, d_prog)
    if (gapfree) { if (probe) LAUNCH(true, false); else LAUNCH(true, true); }
    else { if (probe) LAUNCH(false, false); else LAUNCH(false, true); }
#undef LAUNCH
    ctx->launches++;
    return cudaGetLastError();
}

// one pair per cluster of CL CTAs (NW warps each): persistent clusters, one CTA per SM
template <int D, int NW, int CL>
static cudaError_t launch_cluster(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs,
                                  bool gapfree, bool probe, int *d_counter, PairState *d_state, int *d_ebrow, uint8_t *d_dir, int *d_prog) {
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
    if (e != cudaSuccess) return e;
    const int nclusters = std::max(1, std::min(njobs, ctx->sm_count / CL));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * CL)); cfg.blockDim = dim3(NW * 32); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const DevCM *a0 = cm->d; const int4 *a1 = pool->d_rowp, *a2 = pool->d_colp; const int *a3 = pool->d_h0, *a4 = pool->d_g0;
    const unsigned *a5 = pool->d_rowpk;
#define LAUNCHC(GFV, DIRV) e = cudaLaunchKernelEx(&cfg, k_band2<D, NW, NW, GFV, DIRV, CL>, a0, a1, a2, a3, a4, a5, d_jobs, njobs, d_counter, \
                                                  d_state, d_ebrow, d_dir, d_prog)
    if (gapfree) { if (probe) LAUNCHC(true, false); else LAUNCHC(true, true); }
    else { if (probe) LAUNCHC(false, false); else LAUNCHC(false, true); }
#undef LAUNCHC
    ctx->launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

// class = number of diagonals one CTA covers.  Warps right of the band idle, so the in-between sizes (3, 5, 6
// warps) only buy occupancy: their registers would otherwise sit unused in a 4- or 8-warp CTA.
int band2_class_for(long long B) {
    const int classes[10] = { 64, 128, 256, 512, 768, 1024, 1280, 1536, 2048, 4096 };
    for (int c = 0; c < 10; ++c) if (B <= classes[c]) return classes[c];
    return 0;
}
// bytes per anti-diagonal: one per two diagonals; above 256 diagonals only the warps that reach into the
// band store (128 bytes each)
int band2_stride_for(int cls, long long B) {
    if (cls <= 256) return cls / 2;
    return (int)((B + 255) / 256) * 128;
}

// lowlat: shapes for rounds with so few pairs that the GPU is mostly idle and the round's duration is the wavefront
// latency of ONE pair (anti-diagonals x time per sub-step).  The same band is then spread over 2-4x as many warps
// with 2 or 4 diagonals per thread: the per-sub-step dependency chain shrinks accordingly.  The direction-byte
// layout (byte d/2 of anti-diagonal a) and the stride do not depend on the shape, so the traceback is unchanged.
cudaError_t launch_band2(poy_ctx *ctx, const poy_cm *cm, const poy_pool *pool, const BandJob *d_jobs, int njobs, int cls,
                         bool gapfree, bool probe, int *d_counter, PairState *d_state, int *d_ebrow, uint8_t *d_dir, bool lowlat, int *d_prog) {
    if (njobs <= 0) return cudaSuccess;
    // Cluster shapes (one pair over 2 or 4 SMs, 4 diagonals per thread): latency-bound rounds with bands of 1280 diagonals
    // and more, where a fill lasts as long as the issue rate of ONE SM allows.  POY_CLUSTER=0 turns them off, 2 forces
    // them for every class (test hook).
    {
        const char *ce = getenv("POY_CLUSTER");
        const bool off = ce && ce[0] == '0', forced = ce && ce[0] == '2';
        if (!off && njobs < 60000 && (forced || (lowlat && cls >= 1280))) {
            if (cls <= 2048) return launch_cluster<4, 8, 2>(ctx, cm, pool, d_jobs, njobs, gapfree, probe, d_counter, d_state, d_ebrow, d_dir, d_prog);
            if (cls == 4096) return launch_cluster<4, 8, 4>(ctx, cm, pool, d_jobs, njobs, gapfree, probe, d_counter, d_state, d_ebrow, d_dir, d_prog);
        }
    }
#define L1(DD, WW) return launch_one<DD, WW>(ctx, cm, pool, d_jobs, njobs, gapfree, probe, d_counter, d_state, d_ebrow, d_dir, d_prog)
    if (lowlat)
        switch (cls) {
            case 128: L1(2, 2);
            case 256: L1(2, 4);
            case 512: L1(2, 8);
            case 768: L1(4, 6);
            case 1024: L1(4, 8);
        }
    switch (cls) {
        case 64: L1(2, 1);
        case 128: L1(4, 1);
        case 256: L1(8, 1);
        case 512: L1(8, 2);    // 8 diagonals per thread measured faster than 16 (registers -> occupancy)
        case 768: L1(8, 3);
        case 1024: L1(8, 4);
        case 1280: L1(8, 5);
        case 1536: L1(8, 6);
        case 2048: L1(8, 8);
        case 4096: L1(8, 16);
    }
#undef L1
    return cudaErrorInvalidValue;
}
