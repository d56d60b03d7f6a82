"""Host mirror of ``Sequence.Align`` (src/sequence.ml:596-1826), batch-shaped.

Every function takes a Pool and arrays of pool indices: one call == the list of calls the
reference makes one at a time from ``SeqCS``/``DynamicCS`` while mapping over loci and
candidates.  All alignment work happens in libpoy5b200.so on the GPU."""
import ctypes as C
import numpy as np
from .api import _ptr


def _gap_counts(pool):
    """Sequence.count_gaps for every pool sequence: symbols that carry the gap bit, leading gap included
    (seq_CAML_count, src/seq.c:644-669)."""
    if getattr(pool, "_gapcnt", None) is None:
        has = (pool.data & 16) != 0
        cs = np.concatenate([[0], np.cumsum(has, dtype=np.int64)])
        pool._gapcnt = (cs[pool.offsets[1:]] - cs[pool.offsets[:-1]]).astype(np.int64)
    return pool._gapcnt


def _linear_args(pool, a, b, deltaw):
    """Shorter-first ordering and the deltaw of Sequence.Align.cost_2 (src/sequence.ml:868-925)."""
    a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
    la, lb = pool.lens[a], pool.lens[b]
    swaped = (la > lb)
    s1 = np.where(swaped, b, a).astype(np.int32); s2 = np.where(swaped, a, b).astype(np.int32)
    l1, l2 = pool.lens[s1], pool.lens[s2]
    gc = _gap_counts(pool)
    gaps = np.maximum(gc[a], gc[b])
    dif = l1 - l2
    lower = (l1.astype(np.float64) * 0.10).astype(np.int64)       # int_of_float (float s1len *. 0.10)
    if deltaw is None:
        dcalc = np.where(dif < lower, lower // 2, 2)
    else:
        dcalc = np.where(dif < lower, lower, int(deltaw))
    return s1, s2, (gaps + dcalc).astype(np.int32), swaped.astype(np.uint8)


class Align:
    @staticmethod
    def cost_2(ctx, cm, pool, a, b, deltaw=None):
        """Sequence.Align.cost_2 (src/sequence.ml:928-936).  Affine model: cost_2_affine =
        "algn_CAML_cost_affine_3", either order.  Linear / no-alignment model: c_cost_2 =
        "algn_CAML_simple_2" with the shorter sequence first and deltaw = gaps + deltaw_calc
        (src/sequence.ml:868-925).  Returns int32[n]."""
        if cm.host.cost_model_type != 1:
            s1, s2, dwh, _ = _linear_args(pool, a, b, deltaw)
            cost = np.zeros(len(s1), np.int32)
            ctx.check(ctx.L.poy_batch_cost_linear(ctx.h, cm.h, pool.h, len(s1), _ptr(s1), _ptr(s2), _ptr(dwh), _ptr(cost)))
            return cost
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        n = len(a)
        cost = np.zeros(n, np.int32)
        ctx.check(ctx.L.poy_batch_cost_affine(ctx.h, cm.h, pool.h, n, _ptr(a), _ptr(b), _ptr(cost)))
        return cost

    @staticmethod
    def align_2(ctx, cm, pool, a, b, stats=False):
        """Sequence.Align.align_2 (src/sequence.ml:1061-1082): affine -> align_affine_3; linear ->
        cost_2 + create_edited_2 ("algn_CAML_backtrace_2d").  Returns dict(cost, res_a, res_b)."""
        if cm.host.cost_model_type == 1:
            r = Align.align_affine_3(ctx, cm, pool, a, b, want=("resi", "resj"), stats=stats)
            return r
        s1, s2, dwh, swaped = _linear_args(pool, a, b, None)
        n = len(s1)
        caps = (pool.lens[s1] + pool.lens[s2]).astype(np.int64)
        out_off = np.zeros(n, np.int64)
        if n > 1:
            np.cumsum(caps[:-1], out=out_off[1:])
        total = int(caps.sum())
        r1 = np.zeros(total, np.uint8); r2 = np.zeros(total, np.uint8)
        cost = np.zeros(n, np.int32); out_len = np.zeros(2 * n, np.int32)
        st = np.zeros(4 * n, np.int32) if stats else None
        ctx.check(ctx.L.poy_batch_align_linear(ctx.h, cm.h, pool.h, n, _ptr(s1), _ptr(s2), _ptr(dwh), _ptr(swaped), _ptr(out_off),
                                               _ptr(cost), _ptr(r1), _ptr(r2), _ptr(out_len), _ptr(st)))
        out_len = out_len.reshape(n, 2)
        ends = out_off + caps
        x1 = [r1[ends[p] - out_len[p, 0]:ends[p]] for p in range(n)]
        x2 = [r2[ends[p] - out_len[p, 1]:ends[p]] for p in range(n)]
        res = {"cost": cost, "swaped": swaped,
               "res_a": [x2[p] if swaped[p] else x1[p] for p in range(n)],
               "res_b": [x1[p] if swaped[p] else x2[p] for p in range(n)]}
        if stats:
            res["stats"] = st.reshape(n, 4)
        return res

    @staticmethod
    def full_median_2(ctx, cm, pool, a, b):
        """Sequence.Align.full_median_2 (src/sequence.ml:1136-1142): affine -> the median of align_affine_3; otherwise
        align_2 followed by median_2 of the two rows.  -> list of sequences"""
        if cm.host.cost_model_type == 1:
            return Align.align_affine_3(ctx, cm, pool, a, b, want=("median",))["median"]
        r = Align.align_2(ctx, cm, pool, a, b)
        return median_2(ctx, cm, r["res_a"], r["res_b"], False)

    @staticmethod
    def align_affine_3(ctx, cm, pool, a, b, want=("median", "medianwg", "resi", "resj"), stats=False):
        """Sequence.Align.align_affine_3 (src/sequence.ml:633-649): puts the shorter sequence first,
        passes `swaped`, and un-swaps the two aligned rows on return.  Returns a dict with
        cost[n] and, per requested output, a list of uint8 arrays (a's row is 'res_a', b's 'res_b')."""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        n = len(a)
        la, lb = pool.lens[a], pool.lens[b]
        swaped = (la > lb).astype(np.uint8)           # len1 <= len2 -> 0 (src/sequence.ml:636)
        si = np.where(swaped == 1, b, a).astype(np.int32)
        sj = np.where(swaped == 1, a, b).astype(np.int32)
        caps = (la + lb + 2).astype(np.int64)
        out_off = np.zeros(n, np.int64)
        if n > 1:
            np.cumsum(caps[:-1], out=out_off[1:])
        total = int(caps.sum())
        bufs = {k: (np.zeros(total, np.uint8) if k in want else None) for k in ("median", "medianwg", "resi", "resj")}
        cost = np.zeros(n, np.int32)
        out_len = np.zeros(4 * n, np.int32)
        st = np.zeros(4 * n, np.int32) if stats else None
        any_out = any(v is not None for v in bufs.values())     # cost only: no traceback, no direction bytes
        ctx.check(ctx.L.poy_batch_align_affine(ctx.h, cm.h, pool.h, n, _ptr(si), _ptr(sj), _ptr(swaped), _ptr(out_off),
                                               _ptr(cost), _ptr(bufs["median"]), _ptr(bufs["medianwg"]),
                                               _ptr(bufs["resi"]), _ptr(bufs["resj"]), _ptr(out_len) if any_out else None, _ptr(st)))
        out_len = out_len.reshape(n, 4)
        res = {"cost": cost, "swaped": swaped, "out_len": out_len}
        ends = out_off + caps

        def cut(buf, col):
            return [buf[ends[p] - out_len[p, col]:ends[p]] for p in range(n)]
        if bufs["median"] is not None:
            res["median"] = cut(bufs["median"], 0)
        if bufs["medianwg"] is not None:
            res["medianwg"] = cut(bufs["medianwg"], 1)
        if bufs["resi"] is not None:
            ri = cut(bufs["resi"], 2)
        if bufs["resj"] is not None:
            rj = cut(bufs["resj"], 3)
        if bufs["resi"] is not None and bufs["resj"] is not None:
            res["res_a"] = [rj[p] if swaped[p] else ri[p] for p in range(n)]
            res["res_b"] = [ri[p] if swaped[p] else rj[p] for p in range(n)]
        if stats:
            res["stats"] = st.reshape(n, 4)
        return res


class DevicePool:
    """A pool whose bytes already live in HBM (poy_pool_from_device): `d_data`/`d_off` are raw device
    pointers (ints), `offsets` the host copy of the offsets."""

    def __init__(self, ctx, d_data, d_off, offsets):
        self.ctx = ctx
        self.offsets = np.ascontiguousarray(offsets, np.int64)
        self.nseq = len(self.offsets) - 1
        self.lens = np.diff(self.offsets)
        h = C.c_void_p()
        ctx.check(ctx.L.poy_pool_from_device(ctx.h, C.c_void_p(d_data), C.c_void_p(d_off), _ptr(self.offsets), self.nseq,
                                             C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.poy_pool_free(self.ctx.h, self.h)
            self.h = None


def cost_2_dev(ctx, cm, pool, n, d_a, d_b, d_cost):
    """poy_batch_cost_affine_dev: index and result arrays are device pointers (ints); asynchronous."""
    ctx.check(ctx.L.poy_batch_cost_affine_dev(ctx.h, cm.h, pool.h, int(n), C.c_void_p(d_a), C.c_void_p(d_b),
                                              C.c_void_p(d_cost)))


def align_affine_3_dev(ctx, cm, pool, h_si, h_sj, d_swaped, d_out_off, d_cost, d_median, d_medianwg, d_resi, d_resj,
                       d_out_len):
    """poy_batch_align_affine_dev: outputs stay in HBM.  h_si/h_sj (host int32, shorter first) drive the
    band schedule; every d_* is a raw device pointer or 0."""
    h_si = np.ascontiguousarray(h_si, np.int32); h_sj = np.ascontiguousarray(h_sj, np.int32)

    def vp(x):
        return C.c_void_p(x) if x else None
    ctx.check(ctx.L.poy_batch_align_affine_dev(ctx.h, cm.h, pool.h, len(h_si), None, None, vp(d_swaped), _ptr(h_si),
                                               _ptr(h_sj), vp(d_out_off), vp(d_cost), vp(d_median), vp(d_medianwg),
                                               vp(d_resi), vp(d_resj), vp(d_out_len), None))


# ---- column-wise helpers over aligned rows -------------------------------------------------------------------
def _pack_rows(rows_a, rows_b):
    lens = np.fromiter((len(r) for r in rows_a), np.int32, len(rows_a))
    off = np.zeros(len(rows_a), np.int64)
    if len(rows_a) > 1:
        np.cumsum(lens[:-1], out=off[1:])
    cat = lambda rows: (np.concatenate([np.asarray(r, np.uint8) for r in rows]) if len(rows) and lens.sum() else np.zeros(1, np.uint8))
    return np.ascontiguousarray(cat(rows_a)), np.ascontiguousarray(cat(rows_b)), off, lens


def median_2(ctx, cm, rows_a, rows_b, with_gaps):
    """Sequence.median_2 / median_2_with_gaps (src/sequence.ml:473-492) over lists of aligned rows."""
    a, b, off, lens = _pack_rows(rows_a, rows_b)
    n = len(lens)
    out_off = off + np.arange(n, dtype=np.int64)
    out = np.zeros(int(lens.sum()) + n + 1, np.uint8)
    out_len = np.zeros(n, np.int32)
    ctx.check(ctx.L.poy_batch_median_2(ctx.h, cm.h, n, _ptr(a), _ptr(b), _ptr(off), _ptr(lens), int(with_gaps), _ptr(out_off),
                                       _ptr(out), _ptr(out_len)))
    return [out[out_off[p]:out_off[p] + out_len[p]] for p in range(n)]


def median_3(ctx, cm3, rows_a, rows_b, rows_c):
    """The column-wise three-way median of Sequence.Align.align_3_powell_inter (src/sequence.ml:1342-1369) over lists of
    three aligned rows: -> (medians with the leading gap restored, medians with gaps, sum of Three_D.cost per triple)."""
    a, b, off, lens = _pack_rows(rows_a, rows_b)
    c, _, _, _ = _pack_rows(rows_c, rows_c)
    n = len(lens)
    out_off = off + np.arange(n, dtype=np.int64)
    tot = int(lens.sum()) + n + 1
    med = np.zeros(tot, np.uint8); medwg = np.zeros(tot, np.uint8)
    out_len = np.zeros(n, np.int32); cost3 = np.zeros(n, np.int32)
    ctx.check(ctx.L.poy_batch_median_3(ctx.h, cm3.h, n, _ptr(a), _ptr(b), _ptr(c), _ptr(off), _ptr(lens), _ptr(out_off), _ptr(med),
                                       _ptr(medwg), _ptr(out_len), _ptr(cost3)))
    return ([med[out_off[p]:out_off[p] + out_len[p]] for p in range(n)],
            [medwg[out_off[p]:out_off[p] + lens[p]] for p in range(n)], cost3)


def union(ctx, rows_a, rows_b):
    """Sequence.Align.union (src/sequence.ml:1150-1177 -> algn_CAML_union)."""
    a, b, off, lens = _pack_rows(rows_a, rows_b)
    out = np.zeros(max(1, int(lens.sum())), np.uint8)
    ctx.check(ctx.L.poy_batch_union(ctx.h, len(lens), _ptr(a), _ptr(b), _ptr(off), _ptr(lens), _ptr(out)))
    return [out[off[p]:off[p] + lens[p]] for p in range(len(lens))]


def aligned_cost(ctx, cm, rows_a, rows_b, worst):
    """algn_CAML_worst_2 (worst=True: Sequence.Align.max_cost_2 without its empty-sequence shortcut) or
    algn_CAML_verify_2."""
    a, b, off, lens = _pack_rows(rows_a, rows_b)
    cost = np.zeros(len(lens), np.int32)
    ctx.check(ctx.L.poy_batch_aligned_cost(ctx.h, cm.h, len(lens), _ptr(a), _ptr(b), _ptr(off), _ptr(lens), int(bool(worst)), _ptr(cost)))
    return cost


def ancestor_2(ctx, cm, rows_a, rows_b):
    """Sequence.Align.ancestor_2 (algn_CAML_ancestor_2)."""
    a, b, off, lens = _pack_rows(rows_a, rows_b)
    n = len(lens)
    out_off = off + np.arange(n, dtype=np.int64)
    out = np.zeros(int(lens.sum()) + n + 1, np.uint8)
    out_len = np.zeros(n, np.int32)
    ctx.check(ctx.L.poy_batch_ancestor_2(ctx.h, cm.h, n, _ptr(a), _ptr(b), _ptr(off), _ptr(lens), _ptr(out_off), _ptr(out), _ptr(out_len)))
    return [out[out_off[p]:out_off[p] + out_len[p]] for p in range(n)]


def closest_columns(ctx, cm, rows_parent, rows_mine):
    """The column map of Sequence.Align.closest: get_closest cm parent.(i) mine.(i) for every column of the two
    aligned rows, gaps squeezed out, leading gap restored (src/sequence.ml:1217-1228, 209-222)."""
    a, b, off, lens = _pack_rows(rows_parent, rows_mine)
    n = len(lens)
    out_off = off + np.arange(n, dtype=np.int64)
    out = np.zeros(int(lens.sum()) + n + 1, np.uint8)
    out_len = np.zeros(n, np.int32)
    ctx.check(ctx.L.poy_batch_closest(ctx.h, cm.h, n, _ptr(a), _ptr(b), _ptr(off), _ptr(lens), _ptr(out_off), _ptr(out), _ptr(out_len)))
    return [out[out_off[p]:out_off[p] + out_len[p]] for p in range(n)]


def closest(ctx, cm, pool, parent, mine):
    """Batch Sequence.Align.closest parent mine cm (src/sequence.ml:1180-1237), bitset alphabets with combinations
    (DNA): the single-assignment of `mine` closest to `parent`.  Per pair: empty `mine` -> (mine, 0); identical
    sequences -> gap bits stripped, columns mapped, cost 0; otherwise align_2, map the columns through
    get_closest, squeeze the gaps out and RE-COST the result against the parent with cost_2 ("the set distance
    calculation is an upper bound in the affine gap cost model").  Returns (list of sequences, int64 costs)."""
    from .api import Pool
    parent = np.ascontiguousarray(parent, np.int32); mine = np.ascontiguousarray(mine, np.int32)
    n = len(parent)
    seqs = [pool.data[pool.offsets[s]:pool.offsets[s + 1]] for s in range(pool.nseq)]
    out = [None] * n
    cost = np.zeros(n, np.int64)
    same, diff = [], []
    for p in range(n):
        m = seqs[mine[p]]
        if len(m) == 0 or np.all(m == 16):                      # Sequence.is_empty
            out[p] = m.copy()
        elif len(seqs[parent[p]]) == len(m) and np.array_equal(seqs[parent[p]], m):
            same.append(p)
        else:
            diff.append(p)
    if same:
        strip = lambda s: np.concatenate([s[:1], s[1:] & 15]).astype(np.uint8)
        rows = [strip(seqs[mine[p]]) for p in same]
        for p, r in zip(same, closest_columns(ctx, cm, rows, rows)):
            out[p] = r
    if diff:
        d = np.asarray(diff)
        r = Align.align_2(ctx, cm, pool, parent[d], mine[d])
        res = closest_columns(ctx, cm, r["res_a"], r["res_b"])
        for p, s in zip(diff, res):
            out[p] = s
        # re-cost: a pool holding the parents and the new single-assignment sequences
        both = [seqs[parent[p]] for p in diff] + res
        lens = np.fromiter((len(s) for s in both), np.int64, len(both))
        off = np.zeros(len(both) + 1, np.int64); np.cumsum(lens, out=off[1:])
        p2 = Pool(ctx, data=np.concatenate(both).astype(np.uint8), offsets=off)
        k = len(diff)
        cost[d] = Align.cost_2(ctx, cm, p2, np.arange(k, dtype=np.int32), np.arange(k, 2 * k, dtype=np.int32))
        p2.close()
    return out, cost


def _algn_median(ctx, cm, seqs, ia, ib):
    """`algn s1 s2` of Sequence.readjust (src/sequence.ml:2099-2120): cost and median of one alignment; affine ->
    align_affine_3's median, otherwise align_2 + median_2.  One batch over a temporary pool of `seqs`."""
    from .api import Pool
    pool = Pool(ctx, seqs)
    try:
        if cm.host.cost_model_type == 1:
            r = Align.align_affine_3(ctx, cm, pool, ia, ib, want=("median",))
            return r["cost"].astype(np.int64), r["median"]
        r = Align.align_2(ctx, cm, pool, ia, ib)
        return r["cost"].astype(np.int64), median_2(ctx, cm, r["res_a"], r["res_b"], False)
    finally:
        pool.close()


def readjust(ctx, cm, pool, a, b, parent):
    """Batch Sequence.readjust a b _ cm parent (src/sequence.ml:2097-2156, the `Algn_Normal` branch): the
    approximate three-way re-optimisation of an interior node.  Per node: the medians ab, bc, ac (c = parent), each
    re-aligned with the third sequence; the cheapest composition decides which pair of (sequence, median) goes
    through Align.closest; the new sequence is then aligned with both children and the parent.
    Returns dict(cost3, cost2, sequence, aligned_mp) like the reference's tuple (its last three members are the
    same aligned row)."""
    from .api import Pool
    a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32); c = np.ascontiguousarray(parent, np.int32)
    n = len(a)
    sq = lambda idx: [pool.seq(int(x)) for x in idx]
    A, B, Cc = sq(a), sq(b), sq(c)
    ar = np.arange(n, dtype=np.int32)
    # stage 1: ab, bc, ac in one batch over [A | B | C]
    c1, m1 = _algn_median(ctx, cm, A + B + Cc, np.concatenate([ar, ar + n, ar]), np.concatenate([ar + n, ar + 2 * n, ar + 2 * n]))
    cab, cbc, cac = c1[:n], c1[n:2 * n], c1[2 * n:]
    ab, bc, ac = m1[:n], m1[n:2 * n], m1[2 * n:]
    # stage 2: (ab, c), (bc, a), (ac, b) over [ab | bc | ac | C | A | B]
    c2, _ = _algn_median(ctx, cm, list(ab) + list(bc) + list(ac) + Cc + A + B, np.arange(3 * n, dtype=np.int32),
                         np.arange(3 * n, 6 * n, dtype=np.int32))
    cabc, cbca, cacb = c2[:n] + cab, c2[n:2 * n] + cbc, c2[2 * n:] + cac
    # make_center's choice: closest c ab | closest b ac | closest a bc
    first, second = [], []
    for p in range(n):
        if cabc[p] <= cbca[p]:
            pick = (Cc[p], ab[p]) if cabc[p] <= cacb[p] else (B[p], ac[p])
        else:
            pick = (A[p], bc[p]) if cbca[p] < cacb[p] else (B[p], ac[p])
        first.append(pick[0]); second.append(pick[1])
    p3 = Pool(ctx, first + second)
    new, _ = closest(ctx, cm, p3, ar, ar + n)
    p3.close()
    # stage 3: align_2 of the new sequence with both children and the parent over [A | B | C | new]
    p4 = Pool(ctx, A + B + Cc + list(new))
    r = Align.align_2(ctx, cm, p4, np.arange(3 * n, dtype=np.int32), np.concatenate([ar, ar, ar]) + 3 * n)
    p4.close()
    cost = r["cost"].astype(np.int64)
    cost2 = cost[:n] + cost[n:2 * n]
    return {"cost3": cost2 + cost[2 * n:], "cost2": cost2, "sequence": list(new), "aligned_mp": r["res_b"][2 * n:]}


def select_one(seq):
    """Sequence.select_one (src/sequence.ml:2065-2085) for the bitset alphabets of this path (combine = 1, no levels):
    every symbol is replaced by the smallest element of its set, i.e. its lowest set bit."""
    s = np.asarray(seq, np.uint8)
    return (s & (~s + 1)).astype(np.uint8)


class NewkkAlign:
    """Sequence.NewkkAlign (src/sequence.ml:1831-2062): the diagonal-storage Ukkonen alignment of src/newkkonen.c,
    affine cost model (the reference's non-affine entry point is broken, see include/poy5_b200.h)."""

    @staticmethod
    def _order(pool, a, b):
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        exchange = pool.lens[a] > pool.lens[b]          # NewkkAlign.align_2: shorter first, swaped = 1 if exchanged
        s1 = np.where(exchange, b, a).astype(np.int32); s2 = np.where(exchange, a, b).astype(np.int32)
        return s1, s2, exchange

    @staticmethod
    def cost_2(ctx, cm, pool, a, b):
        """NewkkAlign.cost_2 (src/sequence.ml:1944-1990): newkk_cost2_affine on (shorter, longer).  int32[n]"""
        s1, s2, _ = NewkkAlign._order(pool, a, b)
        n = len(s1)
        cost = np.zeros(n, np.int32)
        ctx.check(ctx.L.poy_batch_newkk_align(ctx.h, cm.h, pool.h, n, _ptr(s1), _ptr(s2), None, None, _ptr(cost), None, None, None, None))
        return cost

    @staticmethod
    def align_2(ctx, cm, pool, a, b, stats=False):
        """NewkkAlign.align_2 (src/sequence.ml:1879-1942, first_gap = true): -> dict(cost, res_a, res_b[, stats]) with
        the aligned rows in the caller's operand order."""
        s1, s2, exchange = NewkkAlign._order(pool, a, b)
        n = len(s1)
        caps = (pool.lens[s1] + pool.lens[s2]).astype(np.int64)
        out_off = np.zeros(n, np.int64)
        if n > 1:
            np.cumsum(caps[:-1], out=out_off[1:])
        total = int(caps.sum())
        r1 = np.zeros(total, np.uint8); r2 = np.zeros(total, np.uint8)
        cost = np.zeros(n, np.int32); out_len = np.zeros(2 * n, np.int32)
        st = np.zeros(4 * n, np.int32) if stats else None
        sw = exchange.astype(np.uint8)
        ctx.check(ctx.L.poy_batch_newkk_align(ctx.h, cm.h, pool.h, n, _ptr(s1), _ptr(s2), _ptr(sw), _ptr(out_off), _ptr(cost),
                                              _ptr(r1), _ptr(r2), _ptr(out_len), _ptr(st)))
        out_len = out_len.reshape(n, 2)
        ends = out_off + caps
        x1 = [r1[ends[p] - out_len[p, 0]:ends[p]] for p in range(n)]
        x2 = [r2[ends[p] - out_len[p, 1]:ends[p]] for p in range(n)]
        res = dict(cost=cost, res_a=[x2[p] if exchange[p] else x1[p] for p in range(n)],
                   res_b=[x1[p] if exchange[p] else x2[p] for p in range(n)])
        if stats:
            res["stats"] = st.reshape(n, 4)
        return res

    @staticmethod
    def full_median_2(ctx, cm, pool, a, b):
        """NewkkAlign.full_median_2 (src/sequence.ml:1992-2000): align_2 then Sequence.median_2 of the two rows."""
        r = NewkkAlign.align_2(ctx, cm, pool, a, b)
        return median_2(ctx, cm, r["res_a"], r["res_b"], False)      # Sequence.median_2 = seq_CAML_median_2_no_gaps (src/sequence.ml:483-490)
