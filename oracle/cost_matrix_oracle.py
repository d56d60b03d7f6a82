"""TEST INFRASTRUCTURE ONLY -- restatement of POY5's ``Cost_matrix.Two_D`` table
construction (reference: ``src/cost_matrix.ml``; OCaml cannot be run in this
image, SURVEY.md F1, so this is a port; it is pinned only indirectly, by the
hand-checkable identities in tests/test_cost_matrix.py -- "parity unpinned by
the reference" for this file).

Tables are bit-indexed ``[a, b]`` with ``a, b in 1 .. 2**a_sz - 1`` (``lcm = a_sz``,
``src/cm.c:903-911``); row/column 0 stay zero as in the ``calloc``'d reference
arrays (``src/cm.c:627``).
"""
import numpy as np

MAX_INT = (2**31 - 1) >> 1  # src/cost_matrix.ml:38


def _bits(v, a_sz):
    """BitSet.Int.list_of_packed_max (src/bitSet.ml:343-353): highest bit first."""
    return [1 << b for b in range(a_sz - 1, -1, -1) if v & (1 << b)]


class CostMatrix2D:
    """Plain container mirroring ``struct cm`` (src/cm.h:33-76) for level = 0."""

    def __init__(self, a_sz=5, cost_model_type=0, gap_open=0, all_elements=31):
        n = 1 << a_sz
        self.a_sz_letters = a_sz
        self.lcm = a_sz
        self.n = n
        self.gap = 1 << (a_sz - 1)          # src/cm.c:592
        self.a_sz = n - 1                   # cm_combinations_of_alphabet
        self.cost_model_type = cost_model_type  # 0 linear, 1 affine, 2 no-align
        self.gap_open = gap_open
        self.all_elements = all_elements
        self.combinations = 1
        self.is_metric = 0
        self.is_identity = 0
        self.cost = np.zeros((n, n), np.int32)
        self.worst = np.zeros((n, n), np.int32)
        self.median = np.zeros((n, n), np.uint8)
        self.prepend = np.zeros(n, np.int32)
        self.tail = np.zeros(n, np.int32)

    def clone(self):
        c = CostMatrix2D.__new__(CostMatrix2D)
        c.__dict__.update(self.__dict__)
        for k in ("cost", "worst", "median", "prepend", "tail"):
            setattr(c, k, getattr(self, k).copy())
        return c

    def min_non0_cost(self):
        """cm_get_min_non0_cost (src/cm.c:1063-1089): scans 2*n*n ints; the
        second half of the calloc'd array is all zero."""
        pos = self.cost[self.cost > 0]
        return int(pos.min()) if pos.size else (2**31 - 1) // 2


def _cleanup(m):
    """src/cost_matrix.ml:700-716 (level = 0 branch)."""
    if m.combinations == 0 or m.cost_model_type != 1:
        return lambda item: item
    gap = m.gap
    return lambda item: gap if (item != gap and (item & gap)) else item


def fill_all_combinations(m):
    """fill_best_cost_and_median_for_all_combinations (src/cost_matrix.ml:862-897)
    with test_combinations (:479-506).  Reads singleton costs from ``m`` itself."""
    a_sz = m.a_sz_letters
    gap, go = m.gap, m.gap_open
    cleanup = _cleanup(m)
    ncomb = (1 << a_sz) - 1
    cost = m.cost
    for i in range(1, ncomb + 1):
        li = _bits(i, a_sz)
        for j in range(1, ncomb + 1):
            lj = _bits(j, a_sz)
            best, c, w = 0, MAX_INT, 0
            for a in li:
                for b in lj:
                    for k in range(a_sz):
                        v = 1 << k
                        goa = go if (m.cost_model_type == 1 and v == gap and (a & gap) and (b & gap)) else 0
                        tc = int(cost[a, v]) + int(cost[v, b]) + goa
                        if tc < c:
                            c, best = tc, v
                        elif tc == c:
                            best |= v
                    cab = int(cost[a, b])
                    if cab > w:
                        w = cab
            if len(li) == 1 and len(lj) == 1:
                m.median[i, j] = i | j
            else:
                m.cost[i, j] = c
                m.median[i, j] = cleanup(best)
            m.worst[i, j] = w


def fill_all_combinations_bitwise(m, create_original=False):
    """fill_best_cost_and_median_for_all_combinations_bitwise
    (src/cost_matrix.ml:721-804).  Reads from a clone of ``m``."""
    a_sz = m.lcm
    old = m.cost.copy()
    ncomb = (1 << a_sz) - 1

    def aux(acc, i, j):
        best, med, worst = acc
        cost1 = int(old[i, i]) + int(old[i, j])
        cost2 = int(old[i, j]) + int(old[j, j])
        if cost1 == cost2:
            costij, medij = (cost1 - int(old[i, i]) if create_original else cost1), i | j
        elif cost1 > cost2:
            costij, medij = (cost2 - int(old[j, j]) if create_original else cost2), j
        else:
            costij, medij = (cost1 - int(old[i, i]) if create_original else cost1), i
        if costij < best:
            best, med = costij, medij
        elif costij == best:
            med = medij | med
        if costij > worst:
            worst = costij
        return best, med, worst

    def process(l1, l2, acc):
        for i in l1:
            for j in l2:
                acc = aux(acc, i, j)
        return acc

    for i in range(1, ncomb + 1):
        li = _bits(i, a_sz)
        for j in range(1, ncomb + 1):
            lj = _bits(j, a_sz)
            if len(li) == 1 and len(lj) == 1:
                if i == j:
                    cii = int(old[i, i])
                    if not create_original:
                        cii *= 2
                    median, best, worst = i, cii, cii
                else:
                    cost1 = int(old[i, j]) + int(old[j, j])
                    cost2 = int(old[i, i]) + int(old[i, j])
                    if cost1 == cost2:
                        costij, med = cost1, i | j
                    elif cost1 > cost2:
                        costij, med = cost2, i
                    else:
                        costij, med = cost1, j
                    if create_original:
                        costij = int(old[i, j])
                    median, best, worst = med, costij, costij
            else:
                best, median, worst = process(li, lj, process(lj, li, (MAX_INT, 0, 0)))
            m.median[i, j] = median
            m.cost[i, j] = best
            m.worst[i, j] = worst


def fill_default_prepend_tail(m):
    """src/cost_matrix.ml:994-1001."""
    for i in range(1, m.a_sz + 1):
        m.tail[i] = m.cost[i, m.gap]
        m.prepend[i] = m.cost[m.gap, i]


def _input_is_metric(arr):
    """input_is_metric (src/cost_matrix.ml:1112-1138): positive && symmetric && zero diagonal."""
    arr = np.asarray(arr)
    ispos = bool((arr >= 0).all())
    issym = bool((arr == arr.T).all())
    iside = bool((np.diag(arr) == 0).all())
    return (ispos and issym and iside), iside


def fill_cost_matrix(rows, all_elements=31, create_original=False):
    """fill_cost_matrix (src/cost_matrix.ml:1140-1189), use_comb=true, level=0:
    matrices are always created Linnear with gap_opening 0 (``use_cost_model``)."""
    rows = np.asarray(rows, dtype=np.int64)
    a_sz = rows.shape[0]
    m = CostMatrix2D(a_sz, 0, 0, all_elements)
    for e1 in range(a_sz):
        for e2 in range(a_sz):
            m.cost[1 << e1, 1 << e2] = rows[e1, e2]
    ismetric, iside = _input_is_metric(rows)
    if iside:
        m.is_identity = 1
    if ismetric:
        m.is_metric = 1
        fill_all_combinations(m)
    else:
        fill_all_combinations_bitwise(m, create_original=create_original)
    fill_default_prepend_tail(m)
    return m


def of_list(rows, all_elements=31):
    """Cost_matrix.Two_D.of_list -> (c2_full, c2_original) (src/cost_matrix.ml:1257-1266)."""
    return fill_cost_matrix(rows, all_elements, False), fill_cost_matrix(rows, all_elements, True)


def of_transformations_and_gaps(trans, gaps, alph_size=5, all_elements=31):
    """src/cost_matrix.ml:1333-1344."""
    rows = [[0 if x == p else (gaps if (x == alph_size - 1 or p == alph_size - 1) else trans)
             for x in range(alph_size)] for p in range(alph_size)]
    return of_list(rows, all_elements)


def set_cost_model(m, model, go=0):
    """set_cost_model (src/cost_matrix.ml:1003-1016): c_set_aff then the
    ``_bitwise`` refill.  model: 0 linear, 1 affine, 2 no-alignment."""
    m.cost_model_type = model
    m.gap_open = go if model == 1 else 0
    fill_all_combinations_bitwise(m)
    return m


def dna_matrices(subst, indel, gap_open=None):
    """What `transform(tcm:(subst,indel)[, gap_opening:go])` leaves in
    (c2_full, c2_original) for nucleotide data (src/data.ml:5937-5964)."""
    full, orig = of_transformations_and_gaps(subst, indel)
    if gap_open is not None:
        full = set_cost_model(full.clone(), 1, gap_open)
        orig = set_cost_model(orig.clone(), 1, gap_open)
    return full, orig


def get_closest(m, a, b):
    """Cost_matrix.Two_D.get_closest (src/cost_matrix.ml:1387-1428), level = 0."""
    gap = m.gap
    if m.combinations == 0:
        return b
    if a == gap or b == gap:
        pass
    elif (a & gap) and (b & gap):
        b = gap
    else:
        b = b & ~gap
    bits = list(reversed(_bits(b, m.a_sz_letters)))  # states_of_code: ascending
    if not bits:
        raise ValueError("~ No bits on?")
    best, cur = a, MAX_INT
    for x in bits:
        nc = int(m.cost[a, x])
        if nc < cur:
            best, cur = x, nc
    return best


def three_d_of_two_dim_comb(m):
    """Cost_matrix.Three_D.of_two_dim_comb (src/cost_matrix.ml:1605-1652): -> (cost int32[32,32,32], median uint8[32,32,32]).
    Restated from the OCaml (cannot run here): parity unpinned by the reference, like the 2-D fill."""
    n, lcm, gap = m.n, m.lcm, m.gap
    cost = np.zeros((n, n, n), np.int32); med = np.zeros((n, n, n), np.uint8)
    for i in range(1, n):
        for j in range(1, n):
            for k in range(1, n):
                best, mm = MAX_INT, 0
                for l in range(lcm):
                    inter = 1 << l
                    shared = int(bool(i & inter)) + int(bool(j & inter)) + int(bool(k & inter))
                    if m.is_metric or shared >= 2 or inter != gap:
                        c = int(m.cost[inter, i]) + int(m.cost[inter, j]) + int(m.cost[inter, k])
                    else:
                        c = MAX_INT
                    if c < best:
                        best, mm = c, inter
                    elif c == best:
                        mm |= inter
                pick = 1
                while not (pick & mm):
                    pick <<= 1
                cost[i, j, k] = best; med[i, j, k] = pick
    return cost, med
