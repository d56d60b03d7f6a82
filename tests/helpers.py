"""Shared test inputs: seeded pair sets that cover the reference's edge cases."""
import numpy as np
from poy5_b200 import synth

REGIMES = dict(synth.REGIMES, R4=(1, 1, 1), R5=(3, 2, 4))


def edge_pairs(seed, n=300, maxlen=40):
    """tiny / ragged / empty / decorated pairs (lengths 0..maxlen bases)."""
    rng = np.random.default_rng(seed)
    seqs = []
    for p in range(n):
        L = int(rng.integers(0, maxlen))
        anc = synth.random_seq(rng, L)
        a = synth.evolve(rng, anc, 0.15, 0.06)
        b = synth.evolve(rng, anc, 0.15, 0.06)
        if p % 3 == 0:
            a, b = synth.decorate(rng, a, 0.1, 0.1), synth.decorate(rng, b, 0.1, 0.1)
        if p % 7 == 0:
            b = synth.random_seq(rng, int(rng.integers(0, maxlen + 10)))
        if p % 11 == 0:
            a = np.zeros(0, np.uint8)
        if p % 13 == 0:
            a = np.full(int(rng.integers(1, 6)), 16, np.uint8)          # a run of pure gap codes
        seqs += [synth.with_gap(a), synth.with_gap(b)]
    idx = np.arange(n, dtype=np.int32)
    return seqs, 2 * idx, 2 * idx + 1


def witness_pairs():
    """F5 / F6 witnesses of SURVEY.md 8c."""
    return {
        "R3": ([16, 4, 4, 4, 4, 2, 1, 2, 1], [16, 4, 4, 1, 4, 4, 2, 1, 2]),
        "R2": ([16, 8, 8, 8, 4, 2, 8, 2, 2, 8, 4, 4, 4, 8, 1, 4, 4, 4, 4, 2],
               [16, 1, 2, 1, 4, 4, 1, 4, 4, 2, 1, 2, 1, 8, 1, 8, 2, 2, 8, 4, 4, 4]),
    }


def oracle_align(P, pc, a, b):
    """Sequence.Align.align_affine_3 semantics on top of the oracle: shorter first + swaped,
    rows un-swapped on return.  -> (cost, median, medianwg, res_a, res_b)"""
    sw = int(len(a) > len(b))
    si, sj = (b, a) if sw else (a, b)
    c, m, w, ri, rj = P.align_affine(pc, si, sj, sw)
    return (c, m, w, rj, ri) if sw else (c, m, w, ri, rj)


def oracle_closest(P, cmo, pc, m, parent, mine, linear_align=None):
    """Sequence.Align.closest parent mine (src/sequence.ml:1180-1237) restated on top of the oracle: align_2,
    get_closest per column, remove_gaps2, re-cost with cost_2.  `m` is the python cost-matrix object of
    oracle.cost_matrix_oracle, `pc` its handle in the C oracle, `linear_align(a, b) -> (cost, ra, rb)` the
    Align.align_2 of the linear-gap model.  -> (sequence, cost)"""
    parent = np.asarray(parent, np.uint8); mine = np.asarray(mine, np.uint8)
    if len(mine) == 0 or bool((mine == 16).all()):
        return mine.copy(), 0
    same = len(parent) == len(mine) and bool((parent == mine).all())
    if same:
        ra = np.concatenate([parent[:1], parent[1:] & 15]).astype(np.uint8)
        rb = ra.copy()
    elif m.cost_model_type == 1:
        _, _, _, ra, rb = oracle_align(P, pc, parent, mine)
    else:
        _, ra, rb = linear_align(parent, mine)
    col = [cmo.get_closest(m, int(a), int(b)) for a, b in zip(ra, rb)]
    res = np.array([16] + [c for c in col if c != 16], np.uint8)
    if same:
        return res, 0
    if m.cost_model_type == 1:
        return res, int(P.cost_affine(pc, parent, res))
    return res, int(linear_align(parent, res)[0])


def oracle_readjust(P, cmo, pc, m, a, b, parent, linear_align=None):
    """Sequence.readjust (src/sequence.ml:2097-2156, Algn_Normal) restated on top of the oracle.
    -> (cost3, cost2, new sequence, aligned row of the new sequence against the parent)"""
    def algn(s1, s2):
        if m.cost_model_type == 1:
            c, med, _, _, _ = oracle_align(P, pc, s1, s2)
            return int(c), med
        c, r1, r2 = linear_align(s1, s2)
        return int(c), P.median_2(pc, r1, r2, False)

    def align_2(s1, s2):
        if m.cost_model_type == 1:
            c, _, _, r1, r2 = oracle_align(P, pc, s1, s2)
            return r1, r2, int(c)
        c, r1, r2 = linear_align(s1, s2)
        return r1, r2, int(c)
    c = parent
    cab, ab = algn(a, b); cbc, bc = algn(b, c); cac, ac = algn(a, c)
    cabc = algn(ab, c)[0] + cab; cbca = algn(bc, a)[0] + cbc; cacb = algn(ac, b)[0] + cac
    if cabc <= cbca:
        x, y = (c, ab) if cabc <= cacb else (b, ac)
    else:
        x, y = (a, bc) if cbca < cacb else (b, ac)
    new, _ = oracle_closest(P, cmo, pc, m, x, y, linear_align)
    _, _, c1 = align_2(a, new)
    _, _, c2 = align_2(b, new)
    _, amp, c3 = align_2(parent, new)
    return c1 + c2 + c3, c1 + c2, new, amp


def oracle_dos_readjust(P, cmo, pc, m, ch1, ch2, parent, mine, mine_costs, ch_sum, distance, linear_align=None):
    """SeqCS.DOS.readjust (src/seqCS.ml:820-947; `ApproxD, DNA alphabet, use_ukk = false) restated on top of the oracle,
    one node.  `distance(a, b)` = DOS.distance with missing_distance 0.
    -> (changed, sequence, from_record, cost2, cost2_max, cost3, sum_cost, aligned row or None)"""
    is_empty = lambda x: bool((np.asarray(x) == 16).all())
    e1, e2, ep = is_empty(ch1), is_empty(ch2), is_empty(parent)
    mc2, mc3, msum = (int(x) for x in mine_costs)
    same = lambda x: len(x) == len(mine) and bool((np.asarray(x) == np.asarray(mine)).all())

    def two_child(a, b):
        if m.cost_model_type == 1:
            c, med, _, _, _ = oracle_align(P, pc, a, b)
        else:
            c, r1, r2 = linear_align(a, b)
            med = P.median_2(pc, r1, r2, False)
        med = np.asarray(med, np.uint8)
        return (med & (~med + 1)).astype(np.uint8), int(c)
    if not (e1 or e2 or ep):
        c3, c2, new, amp = oracle_readjust(P, cmo, pc, m, ch1, ch2, parent, linear_align)
        mx = int(P.worst_2(pc, amp, amp))
        ch = (c2 + ch_sum != msum) or c3 != mc3 or c2 != mc2 or not same(new)
        return ch, new, -1, c2, mx, c3, c2 + ch_sum, amp
    if (e1 and e2) or (e1 and ep) or (e2 and ep):
        which = 2 if (e1 and e2) else (1 if (e1 and ep) else 0)
        r = (ch1, ch2, parent)[which]
        ch = ch_sum != msum or mc3 != 0 or mc2 != 0 or not same(r)
        return ch, np.asarray(r, np.uint8), which, 0, 0, 0, ch_sum, None
    if ep:
        new, c2 = two_child(ch1, ch2)
        ch = (c2 + ch_sum != msum) or c2 != mc3 or c2 != mc2 or not same(new)
        return ch, new, -1, c2, 0, c2, c2 + ch_sum, None
    child = ch1 if e2 else ch2
    new, _ = two_child(child, parent)
    c2 = int(distance(child, new)); c3 = c2 + int(distance(new, parent))
    ch = (c2 + ch_sum != msum) or c3 != mc3 or c2 != mc2 or not same(new)
    return ch, new, -1, c2, 0, c3, c2 + ch_sum, None
