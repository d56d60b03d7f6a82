"""Implied alignment (poy5_b200/implied_alignment.py): properties every implied alignment has, on the
CPU checker; the GPU-driven composition equals the checker-driven one (GPU box)."""
import numpy as np
import pytest

from oracle import cost_matrix_oracle as cmo
from poy5_b200 import implied_alignment as ia, synth, treesearch
from tests.oracle_backend import OracleBackend
from tests.helpers import oracle_align


def loci_taxa(seed, n, L):
    rng = np.random.default_rng(seed)
    pool = [synth.random_seq(rng, L)]
    while len(pool) < n:
        p = pool.pop(int(rng.integers(0, len(pool))))
        pool += [synth.evolve(rng, p, 0.08, 0.02), synth.evolve(rng, p, 0.08, 0.02)]
    return [synth.with_gap(s) for s in pool[:n]]


def build(port, n, L, seed, go=3):
    full, orig = cmo.dna_matrices(1, 1, go)
    b = OracleBackend(port, full, orig)
    taxa = loci_taxa(seed, n, L)
    tree = treesearch.wagner_build(taxa, b)
    root = tree.edges()[0]
    _, single_cost, singles = treesearch.single_assignment(tree, [taxa], b, root)
    seqs = {x: (taxa[x] if x < n else np.asarray(singles[0][x], np.uint8)) for x in tree.adj}
    return full, tree, root, taxa, seqs, single_cost


def check_properties(mat, leaves, taxa, tree, root, seqs, align2, cost):
    # every row spells its taxon's sequence, the leading gap column included; no all-gap column besides column 0
    for row, x in enumerate(leaves):
        r = mat[row]
        assert r[0] == 16 and np.array_equal(r[r != 16], taxa[x][1:])
    assert not (mat[:, 1:] == 16).all(axis=0).any()
    # two bases of one column are connected through a chain of aligned columns of the tree's edge alignments: the
    # number of columns is at most the total length and at least the longest sequence
    assert max(len(t) for t in taxa) <= mat.shape[1] <= sum(len(t) - 1 for t in taxa) + 1


def test_implied_alignment_properties(port):
    for seed, n, L in [(3, 6, 40), (4, 9, 70), (5, 12, 55)]:
        full, tree, root, taxa, seqs, _ = build(port, n, L, seed)
        pc = port.cm(full)

        def align2(pairs):
            out = []
            for a, b in pairs:
                _, _, _, ra, rb = oracle_align(port, pc, a, b)
                out.append((ra, rb))
            return out
        mat, leaves, redone = ia.implied_alignment(tree, root, seqs, full.cost, align2)
        assert leaves == list(range(n)) and redone == 0
        check_properties(mat, leaves, taxa, tree, root, seqs, align2, full.cost)
        # identical sequences give a gap-free alignment of that sequence
    same = [taxa[0]] * 5
    b = OracleBackend(port, *cmo.dna_matrices(1, 1, 3))
    tree = treesearch.wagner_build(same, b)
    seqs = {x: same[0] for x in tree.adj}
    full, _ = cmo.dna_matrices(1, 1, 3)
    pc = port.cm(full)
    mat, leaves, _ = ia.implied_alignment(tree, tree.edges()[0], seqs, full.cost,
                                          lambda pairs: [oracle_align(port, pc, a, b)[3:5] for a, b in pairs])
    assert mat.shape == (5, len(same[0])) and all(np.array_equal(r, same[0]) for r in mat)


def test_ancestor_drops_a_gap_median_column():
    """a column whose two symbols both carry the gap bit costs less with the gaps (0) than without: both positions keep
    their own codes and the ancestor loses the position (src/impliedAlignment.ml:456-466, 485-503)"""
    full, _ = cmo.dna_matrices(1, 1, 3)
    c = iter(range(1, 100))
    cg = lambda: next(c)
    a = ia.create_ias(np.array([16, 1, 17, 4], np.uint8), cg)
    b = ia.create_ias(np.array([16, 1, 18, 4], np.uint8), cg)
    r = ia.ancestor(a, b, (a.seq.copy(), b.seq.copy()), full.cost)
    assert list(r.seq) == [16, 1, 4] and len(r.order) == 4 and sorted(len(v) for v in r.hom.values()) == [1, 1, 2, 2]


@pytest.mark.gpu
def test_gpu_composition_equals_checker(ctx, port):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    full, tree, root, taxa, seqs, _ = build(port, 14, 300, 11)
    pc = port.cm(full)
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, 3).full)
    calls = []

    def align2_gpu(pairs):
        calls.append(len(pairs))
        flat = [x for ab in pairs for x in ab]
        pool = pb.Pool(ctx, flat)
        n = len(pairs)
        r = Align.align_2(ctx, cm, pool, np.arange(0, 2 * n, 2, dtype=np.int32), np.arange(1, 2 * n, 2, dtype=np.int32))
        pool.close()
        return list(zip(r["res_a"], r["res_b"]))
    m1, l1, redone = ia.implied_alignment(tree, root, seqs, full.cost, align2_gpu)
    m2, l2, _ = ia.implied_alignment(tree, root, seqs, full.cost, lambda pairs: [oracle_align(port, pc, a, b)[3:5] for a, b in pairs])
    assert calls == [2 * 14 - 3] and redone == 0          # ONE batch: every edge of the tree
    assert l1 == l2 and np.array_equal(m1, m2)
    check_properties(m1, l1, taxa, tree, root, seqs, None, full.cost)
    cm.close()


def test_symbols_and_fasta():
    """Alphabet.nucleotides code -> symbol (src/alphabet.ml:270-315) and the fasta writer"""
    import io
    from poy5_b200 import implied_alignment as ia
    m = np.array([[16, 1, 2, 16, 8, 15], [16, 4, 16, 16, 10, 17]], np.uint8)
    assert ia.to_strings(m) == ["AC-TN", "G--Y1"]
    assert ia.NUCLEOTIDE_SYMBOLS[8] == "T" and ia.NUCLEOTIDE_SYMBOLS[31] == "*" and ia.NUCLEOTIDE_SYMBOLS[26] == "0"
    out = io.StringIO()
    ia.write_fasta(out, m, ["t1", "t2"], width=3)
    assert out.getvalue() == ">t1\nAC-\nTN\n>t2\nG--\nY1\n"
