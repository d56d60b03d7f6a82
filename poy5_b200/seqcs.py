"""Host mirror of ``SeqCS.DOS`` (src/seqCS.ml), batch-shaped: the per-locus direct-optimization
character.  One call == the ``Array_ops.map`` over loci / the Parmap map over candidates that the
reference performs one alignment at a time (src/seqCS.ml:2232-2345, src/ptree.ml:1356-1408)."""
import numpy as np
from . import sequence
from .api import _ptr
from .sequence import Align


class Heuristic:
    """`h` of src/seqCS.ml:51-55: the pair (c2_full, c2_original) as device-resident cost models."""

    def __init__(self, c2_full, c2_original):
        self.c2_full = c2_full
        self.c2_original = c2_original


def _is_empty(pool):
    """Sequence.is_empty (src/sequence.ml:241-251): every symbol equals the gap code."""
    if getattr(pool, "_empty", None) is None:
        if pool.nseq == 0:
            pool._empty = np.zeros(0, bool)
        elif (pool.lens > 0).all():
            # per-sequence "any symbol differs from the gap code" (segments are non-empty: reduceat is exact)
            pool._empty = np.maximum.reduceat((pool.data != 16).view(np.uint8), pool.offsets[:-1]) == 0
        else:
            cs = np.concatenate([[0], np.cumsum(pool.data != 16, dtype=np.int64)])
            pool._empty = (cs[pool.offsets[1:]] - cs[pool.offsets[:-1]]) == 0
    return pool._empty


class DOS:
    @staticmethod
    def distance(ctx, h, pool, a, b, missing_distance=0):
        """DOS.distance (src/seqCS.ml:701-774): cost-only alignment under c2_ORIGINAL; an empty sequence
        on either side returns missing_distance.  deltaw = max(|len a - len b|, 8) as in the reference."""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        empty = _is_empty(pool)
        skip = empty[a] | empty[b]
        res = np.full(len(a), missing_distance, np.int64)
        idx = np.flatnonzero(~skip)
        if len(idx):
            cm = h.c2_original
            if cm.host.cost_model_type == 1:
                res[idx] = Align.cost_2(ctx, cm, pool, a[idx], b[idx])
            else:
                # one deltaw per pair: run the linear entry point with the per-pair value
                la, lb = pool.lens[a[idx]], pool.lens[b[idx]]
                dw = np.maximum(np.abs(la - lb), 8)
                out = np.zeros(len(idx), np.int64)
                for v in np.unique(dw):
                    m = dw == v
                    out[m] = Align.cost_2(ctx, cm, pool, a[idx][m], b[idx][m], deltaw=int(v))
                res[idx] = out
        return res

    @staticmethod
    def median_cost(ctx, h, pool, a, b):
        """The part of DOS.median a tree pass consumes -- the median sequence and cost2 (affine model; same
        empty-child rule) -- without reading the aligned rows / median_wg back or computing cost2_max.
        Returns (list of sequences, int64 costs)."""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        n = len(a)
        empty = _is_empty(pool)
        ea, eb = empty[a], empty[b]
        seqs, cost = [None] * n, np.zeros(n, np.int64)
        for p in np.flatnonzero(ea | eb):
            seqs[p] = pool.seq(b[p] if ea[p] else a[p]).copy()
        idx = np.flatnonzero(~(ea | eb))
        if len(idx):
            r = Align.align_affine_3(ctx, h.c2_full, pool, a[idx], b[idx], want=("median",))
            for q, p in enumerate(idx):
                seqs[p] = r["median"][q]
            cost[idx] = r["cost"]
        return seqs, cost

    @staticmethod
    def median(ctx, h, pool, a, b):
        """DOS.median (src/seqCS.ml:985-1084) for zero-diagonal (identity) matrices: an empty child yields
        the other child with cost 0; otherwise align under c2_FULL (affine: align_affine_3; linear:
        align_2 + ancestor_2 + median_2_with_gaps) and max_cost_2 of the aligned rows.
        Returns dict(sequence, aligned_a, aligned_b, median_wg, cost2, cost2_max)."""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        n = len(a)
        empty = _is_empty(pool)
        ea, eb = empty[a], empty[b]
        out = dict(sequence=[None] * n, aligned_a=[None] * n, aligned_b=[None] * n, median_wg=[None] * n,
                   cost2=np.zeros(n, np.int64), cost2_max=np.zeros(n, np.int64))
        for p in np.flatnonzero(ea | eb):
            keep, lost = (b[p], a[p]) if ea[p] else (a[p], b[p])
            out["sequence"][p] = pool.seq(keep).copy()
            out["aligned_a"][p] = pool.seq(lost).copy(); out["aligned_b"][p] = pool.seq(lost).copy()
            out["median_wg"][p] = pool.seq(keep).copy()
        idx = np.flatnonzero(~(ea | eb))
        if len(idx) == 0:
            return out
        cm = h.c2_full
        if cm.host.cost_model_type == 1:
            r = Align.align_affine_3(ctx, cm, pool, a[idx], b[idx])
            med, mwg = r["median"], r["medianwg"]
        else:
            r = Align.align_2(ctx, cm, pool, a[idx], b[idx])
            med = sequence.ancestor_2(ctx, cm, r["res_a"], r["res_b"])
            mwg = sequence.median_2(ctx, cm, r["res_a"], r["res_b"], True)
        mx = sequence.aligned_cost(ctx, cm, r["res_a"], r["res_b"], worst=True)
        for q, p in enumerate(idx):
            out["sequence"][p] = med[q]; out["median_wg"][p] = mwg[q]
            out["aligned_a"][p] = r["res_a"][q]; out["aligned_b"][p] = r["res_b"][q]
        out["cost2"][idx] = r["cost"]; out["cost2_max"][idx] = mx
        return out


def dist_2(ctx, h, pool, n, a, b, use_ukk=False):
    """DOS.dist_2 h n a b use_ukk (src/seqCS.ml:1181-1198): the cost of joining node `n` between `a` and `b`:
    tmp = n itself if a is empty, else Sequence.Align.full_median_2 a b under c2_FULL; cost = Sequence.Align.cost_2 n tmp
    (Sequence.NewkkAlign.cost_2 when use_ukk) under c2_FULL.  n, a, b: pool indices.  -> int64[len(n)]"""
    from .api import Pool
    from .sequence import NewkkAlign
    n = np.ascontiguousarray(n, np.int32); a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
    m = len(n)
    empty = _is_empty(pool)
    ea = empty[a]
    tmp = [None] * m
    for p in np.flatnonzero(ea):
        tmp[p] = pool.seq(int(n[p]))
    idx = np.flatnonzero(~ea)
    if len(idx):
        for q, med in zip(idx, Align.full_median_2(ctx, h.c2_full, pool, a[idx], b[idx])):
            tmp[q] = med
    scratch = Pool(ctx, [pool.seq(int(x)) for x in n] + [np.asarray(t, np.uint8) for t in tmp])
    i0 = np.arange(m, dtype=np.int32); i1 = i0 + m
    cost = (NewkkAlign.cost_2 if use_ukk else Align.cost_2)(ctx, h.c2_full, scratch, i0, i1)
    scratch.close()
    return np.asarray(cost, np.int64)


def median_3_union(ctx, cm2, pool, parent, aligned_a, aligned_b):
    """DOS.median_3_union (src/seqCS.ml:1151-1178): the live three-sequence path used for final-state
    assignment.  For every node: union of its two aligned children
    (Sequence.Align.union -> algn_CAML_union), ONE pairwise alignment parent x union under cm2
    (Sequence.Align.align_2), median_2 of the two aligned rows (gap-free, leading gap restored) and
    max_cost_2.  `parent` are pool indices, aligned_a / aligned_b lists of equal-length aligned rows.
    Returns dict(sequence, cost, cost_max, aligned_parent, aligned_union)."""
    from .api import Pool
    parent = np.ascontiguousarray(parent, np.int32)
    n = len(parent)
    un = sequence.union(ctx, aligned_a, aligned_b)
    # the unions become sequences of a scratch pool next to the parents
    seqs = [pool.seq(int(p)) for p in parent] + [np.asarray(u, np.uint8) for u in un]
    tmp = Pool(ctx, seqs)
    ia = np.arange(n, dtype=np.int32); ib = ia + n
    r = Align.align_2(ctx, cm2, tmp, ia, ib)
    med = sequence.median_2(ctx, cm2, r["res_a"], r["res_b"], False)
    mx = sequence.aligned_cost(ctx, cm2, r["res_a"], r["res_b"], worst=True)
    tmp.close()
    return dict(sequence=med, cost=np.asarray(r["cost"], np.int64), cost_max=mx.astype(np.int64),
                aligned_parent=r["res_a"], aligned_union=r["res_b"])


def _two_child(ctx, cm, pool, a, b):
    """readjust_algn_two_child (src/seqCS.ml:783-815, use_ukk = false): the median of one alignment, made unambiguous
    with Sequence.select_one.  Returns (list of sequences, int64 costs)."""
    if cm.host.cost_model_type == 1:
        r = Align.align_affine_3(ctx, cm, pool, a, b, want=("median",))
        med = r["median"]
    else:
        r = Align.align_2(ctx, cm, pool, a, b)
        med = sequence.median_2(ctx, cm, r["res_a"], r["res_b"], False)
    return [sequence.select_one(m) for m in med], r["cost"].astype(np.int64)


def readjust(ctx, h, pool, ch1, ch2, parent, mine, mine_costs, ch_sum_cost, use_ukk=False):
    """DOS.readjust (src/seqCS.ml:820-947), `ApproxD mode, DNA (non-custom) alphabet, batched over nodes: the approximate
    iterative re-optimisation of interior nodes given both children and the parent.

    ch1, ch2, parent, mine: pool indices (int32[n]); mine_costs: int64[n, 3] = (cost2, cost3, sum_cost) of `mine`;
    ch_sum_cost: int64[n] = ch1.sum_cost + ch2.sum_cost.  Returns dict:
      changed bool[n]; sequence (list); aligned (list of the row stored three times in aligned_children, or None when
      the node keeps another record's aligned_children); from_record int8[n] (-1: new record, 0/1/2: the node becomes a
      copy of ch1 / ch2 / parent with sum_cost replaced, the "two of three are empty" cases);
      cost2, cost2_max, cost3, sum_cost int64[n] (for from_record >= 0 only sum_cost is replaced in the copied record:
      cost2 / cost3 here are the 0, 0 the reference returns beside it).
    The five cases follow the reference's match on (is_empty ch1, is_empty ch2, is_empty parent)."""
    if use_ukk:
        raise NotImplementedError("DOS.readjust with use_ukk: route the alignments through sequence.NewkkAlign")
    from .api import Pool
    ch1 = np.ascontiguousarray(ch1, np.int32); ch2 = np.ascontiguousarray(ch2, np.int32)
    parent = np.ascontiguousarray(parent, np.int32); mine = np.ascontiguousarray(mine, np.int32)
    mc = np.asarray(mine_costs, np.int64).reshape(-1, 3)
    chs = np.asarray(ch_sum_cost, np.int64)
    n = len(mine)
    empty = _is_empty(pool)
    e1, e2, ep = empty[ch1], empty[ch2], empty[parent]
    out = dict(changed=np.zeros(n, bool), sequence=[None] * n, aligned=[None] * n, from_record=np.full(n, -1, np.int8),
               cost2=np.zeros(n, np.int64), cost2_max=np.zeros(n, np.int64), cost3=np.zeros(n, np.int64),
               sum_cost=np.zeros(n, np.int64))
    differs = lambda p, s: not np.array_equal(pool.seq(int(mine[p])), s)
    cm = h.c2_full
    # -- nobody empty: Sequence.readjust + max_cost_2 of the aligned row with itself
    idx = np.flatnonzero(~e1 & ~e2 & ~ep)
    if len(idx):
        r = sequence.readjust(ctx, cm, pool, ch1[idx], ch2[idx], parent[idx])
        mx = sequence.aligned_cost(ctx, cm, r["aligned_mp"], r["aligned_mp"], worst=True)
        for q, p in enumerate(idx):
            c2, c3 = int(r["cost2"][q]), int(r["cost3"][q])
            out["sequence"][p] = r["sequence"][q]; out["aligned"][p] = r["aligned_mp"][q]
            out["cost2"][p] = c2; out["cost3"][p] = c3; out["cost2_max"][p] = int(mx[q]); out["sum_cost"][p] = c2 + chs[p]
            out["changed"][p] = (c2 + chs[p] != mc[p, 2]) or c3 != mc[p, 1] or c2 != mc[p, 0] or differs(p, r["sequence"][q])
    # -- at least two of the three empty: the node becomes the record the reference's match binds to `r`
    for p in np.flatnonzero((e1 & e2) | (e1 & ep) | (e2 & ep)):
        which = 2 if (e1[p] and e2[p]) else (1 if (e1[p] and ep[p]) else 0)
        src = (ch1, ch2, parent)[which][p]
        out["from_record"][p] = which
        out["sequence"][p] = pool.seq(int(src)).copy()
        out["sum_cost"][p] = chs[p]
        out["changed"][p] = chs[p] != mc[p, 2] or mc[p, 1] != 0 or mc[p, 0] != 0 or differs(p, out["sequence"][p])
    # -- only the parent empty: the two children decide
    idx = np.flatnonzero(~e1 & ~e2 & ep)
    if len(idx):
        med, cost = _two_child(ctx, cm, pool, ch1[idx], ch2[idx])
        for q, p in enumerate(idx):
            c2 = int(cost[q])
            out["sequence"][p] = med[q]; out["cost2"][p] = c2; out["cost3"][p] = c2; out["sum_cost"][p] = c2 + chs[p]
            out["changed"][p] = (c2 + chs[p] != mc[p, 2]) or c2 != mc[p, 1] or c2 != mc[p, 0] or differs(p, med[q])
    # -- one child empty: the other child and the parent decide; the costs are DOS.distance to each of them
    for child, mask in ((ch1, ~e1 & e2 & ~ep), (ch2, e1 & ~e2 & ~ep)):
        idx = np.flatnonzero(mask)
        if len(idx) == 0:
            continue
        med, _ = _two_child(ctx, cm, pool, child[idx], parent[idx])
        k = len(idx)
        tmp = Pool(ctx, [pool.seq(int(x)) for x in child[idx]] + [pool.seq(int(x)) for x in parent[idx]] + list(med))
        ar = np.arange(k, dtype=np.int32)
        d = DOS.distance(ctx, h, tmp, np.concatenate([ar, ar + 2 * k]), np.concatenate([ar + 2 * k, ar + k]), 0)
        tmp.close()
        for q, p in enumerate(idx):
            c2 = int(d[q]); c3 = c2 + int(d[k + q])
            out["sequence"][p] = med[q]; out["cost2"][p] = c2; out["cost3"][p] = c3; out["sum_cost"][p] = c2 + chs[p]
            out["changed"][p] = (c2 + chs[p] != mc[p, 2]) or c3 != mc[p, 1] or c2 != mc[p, 0] or differs(p, med[q])
    return out


def to_single(ctx, h, pool, parent, mine):
    """DOS.to_single (src/seqCS.ml:950-982): the single-assignment sequence of `mine` given its parent's.  An empty
    `mine` stays empty (cost 0); an empty parent is replaced by `mine` itself; otherwise
    Sequence.Align.closest parent mine c2_FULL.  Returns (list of sequences, int64 costs)."""
    parent = np.ascontiguousarray(parent, np.int32).copy(); mine = np.ascontiguousarray(mine, np.int32)
    empty = _is_empty(pool)
    parent[empty[parent]] = mine[empty[parent]]
    return sequence.closest(ctx, h.c2_full, pool, parent, mine)


# ---- the same policy inside the C ABI (poy5_b200/csrc/dos.cu) -------------------------------------------------------
def dos_distance(ctx, h, pool, a, b, missing_distance=0):
    """poy_dos_distance: DOS.distance with the empty-sequence / ordering / deltaw rules applied in C++.  int64[n]."""
    a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
    cost = np.zeros(len(a), np.int32)
    ctx.check(ctx.L.poy_dos_distance(ctx.h, h.c2_original.h, pool.h, len(a), _ptr(a), _ptr(b), int(missing_distance), _ptr(cost)))
    return cost.astype(np.int64)


def dos_median(ctx, h, pool, a, b):
    """poy_dos_median (affine model): (list of median sequences, int64 cost2) with the empty-child / shorter-first /
    swaped rules applied in C++."""
    a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
    n = len(a)
    caps = (pool.lens[a] + pool.lens[b] + 2).astype(np.int64)
    out_off = np.zeros(n, np.int64)
    if n > 1:
        np.cumsum(caps[:-1], out=out_off[1:])
    buf = np.zeros(max(1, int(caps.sum())), np.uint8)
    cost = np.zeros(n, np.int32); out_len = np.zeros(n, np.int32)
    ctx.check(ctx.L.poy_dos_median(ctx.h, h.c2_full.h, pool.h, n, _ptr(a), _ptr(b), _ptr(out_off), _ptr(cost), _ptr(buf), _ptr(out_len)))
    ends = out_off + caps
    return [buf[ends[p] - out_len[p]:ends[p]] for p in range(n)], cost.astype(np.int64)
