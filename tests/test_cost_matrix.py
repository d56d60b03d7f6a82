"""Cost_matrix.Two_D table construction: C++ product side (poy_cm_fill) vs the python restatement,
plus hand-checkable identities (the reference ships no fixtures for this)."""
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from poy5_b200.cost_matrix import Two_D, tables, min_non0, get_closest

CASES = [(1, 1, 3), (2, 1, 5), (1, 2, 0), (1, 1, None), (3, 2, None), (4, 7, 2)]


@pytest.mark.parametrize("s,g,go", CASES)
def test_fill_matches_python_restatement(s, g, go):
    t = Two_D.of_transformations_and_gaps(s, g, go)
    f, o = cmo.dna_matrices(s, g, go)
    for mine, ref in ((t.full, f), (t.original, o)):
        tb = tables(mine)
        for k in ("cost", "worst", "median", "prepend", "tail"):
            assert np.array_equal(tb[k], getattr(ref, k)), k
        assert tb["gap_open"] == ref.gap_open and tb["cost_model_type"] == ref.cost_model_type
        assert min_non0(mine) == ref.min_non0_cost()


def test_non_metric_input_uses_bitwise_fill():
    rows = [[0, 1, 2, 3, 4], [2, 0, 1, 1, 2], [1, 1, 0, 5, 2], [3, 1, 5, 1, 2], [4, 2, 2, 2, 0]]
    t = Two_D.of_list(rows, 4)
    f, o = cmo.of_list(rows)
    f = cmo.set_cost_model(f.clone(), 1, 4); o = cmo.set_cost_model(o.clone(), 1, 4)
    for mine, ref in ((t.full, f), (t.original, o)):
        tb = tables(mine)
        for k in ("cost", "worst", "median"):
            assert np.array_equal(tb[k], getattr(ref, k)), k


def test_identities_default_matrix():
    tb = tables(Two_D.of_transformations_and_gaps(1, 1, 3).full)
    c, m = tb["cost"], tb["median"]
    assert np.array_equal(c, c.T)
    for a in range(1, 32):
        assert c[a, a] == 0
        for b in range(1, 32):
            assert c[a, b] == (0 if a & b else 1)          # bitsets that intersect cost nothing
    for a in (1, 2, 4, 8, 16):
        for b in (1, 2, 4, 8, 16):
            assert m[a, b] == (a if a == b else a | b)     # singleton medians
    assert (tb["prepend"][1:] == c[16, 1:]).all() and (tb["tail"][1:] == c[1:, 16]).all()
    assert (c[0, :] == 0).all() and (c[:, 0] == 0).all()  # row/column 0 stay calloc-zero (A5)


def test_get_closest():
    t = Two_D.of_transformations_and_gaps(2, 1, 5)
    f, _ = cmo.dna_matrices(2, 1, 5)
    for a in range(1, 32):
        for b in range(1, 32):
            assert get_closest(t.full, a, b) == cmo.get_closest(f, a, b)


@pytest.mark.parametrize("s,g,go", [(1, 1, 3), (2, 1, 5), (1, 1, None)])
def test_three_d_fill_matches_python_restatement(s, g, go):
    """Cost_matrix.Three_D.of_two_dim_comb: C++ product side vs the python restatement, plus identities"""
    from poy5_b200.cost_matrix import Three_D
    t = Two_D.of_transformations_and_gaps(s, g, go)
    f, _ = cmo.dna_matrices(s, g, go)
    c, m = cmo.three_d_of_two_dim_comb(f)
    tb = Three_D.of_two_dim(t.full).tables()
    assert np.array_equal(tb["cost"], c) and np.array_equal(tb["median"], m)
    c2 = tables(t.full)["cost"]
    for a in (1, 2, 4, 8, 16):
        assert tb["cost"][a, a, a] == 0 and tb["median"][a, a, a] == a
        for b in (1, 2, 4, 8, 16):
            assert tb["cost"][a, a, b] == c2[a, b] and tb["median"][a, a, b] == a      # two against one: the majority symbol
    assert tb["median"][3, 5, 9] == 1 and tb["cost"][3, 5, 9] == 0                       # the shared bit wins


@pytest.mark.gpu
def test_median_3_columns(ctx):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Three_D
    from poy5_b200 import sequence
    t = Two_D.of_transformations_and_gaps(1, 1, 3)
    f, _ = cmo.dna_matrices(1, 1, 3)
    c, m = cmo.three_d_of_two_dim_comb(f)
    cm3 = pb.CostModel3D(ctx, Three_D.of_two_dim(t.full))
    rng = np.random.default_rng(9)
    rows = [[rng.integers(1, 32, size=int(L)).astype(np.uint8) for L in (0, 1, 31, 32, 33, 500, 1777)] for _ in range(3)]
    med, medwg, cost3 = sequence.median_3(ctx, cm3, rows[0], rows[1], rows[2])
    for p in range(len(rows[0])):
        a, b, cc = rows[0][p], rows[1][p], rows[2][p]
        wg = m[a, b, cc] if len(a) else np.zeros(0, np.uint8)
        assert np.array_equal(medwg[p], wg)
        assert np.array_equal(med[p], np.concatenate([[16], wg[wg != 16]]).astype(np.uint8))
        assert cost3[p] == int(c[a, b, cc].sum()) if len(a) else cost3[p] == 0
    cm3.close()
