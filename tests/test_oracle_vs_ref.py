"""Pins the C restatement (oracle/do_oracle.c) against the UNMODIFIED reference C compiled from
/root/reference/src (oracle/_ref).  Runs only where that library exists (the build container)."""
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from tests.helpers import REGIMES, edge_pairs, witness_pairs
from poy5_b200 import synth


@pytest.mark.parametrize("rname", list(REGIMES))
def test_random_pairs(reflib, port, rname):
    s_, g_, go = REGIMES[rname]
    full, _ = cmo.dna_matrices(s_, g_, go)
    rc, pc = reflib.cm(full), port.cm(full)
    seqs, ia, ib = edge_pairs(100 + len(rname) + go, n=400, maxlen=60)
    rng = np.random.default_rng(5)
    for p in range(len(ia)):
        a, b = seqs[ia[p]], seqs[ib[p]]
        assert reflib.cost_affine(rc, a, b) == port.cost_affine(pc, a, b)
        si, sj = (a, b) if len(a) <= len(b) else (b, a)
        sw = int(rng.integers(0, 2))
        r1, r2 = reflib.align_affine(rc, si, sj, sw), port.align_affine(pc, si, sj, sw)
        assert r1[0] == r2[0]
        for x, y in zip(r1[1:], r2[1:]):
            assert np.array_equal(x, y)


def test_medium_pairs(reflib, port):
    full, _ = cmo.dna_matrices(1, 1, 3)
    rc, pc = reflib.cm(full), port.cm(full)
    seqs, ia, ib = synth.pair_batch(77, 12, 700, frac_decorated=0.5, jitter=0.3)
    for p in range(len(ia)):
        a, b = seqs[ia[p]], seqs[ib[p]]
        assert reflib.cost_affine(rc, a, b) == port.cost_affine(pc, a, b)
        si, sj = (a, b) if len(a) <= len(b) else (b, a)
        r1, r2 = reflib.align_affine(rc, si, sj, 0), port.align_affine(pc, si, sj, 0)
        assert r1[0] == r2[0] and all(np.array_equal(x, y) for x, y in zip(r1[1:], r2[1:]))


def test_witnesses(reflib, port):
    """cost-only and banded align disagree with each other on these; the
    restatement must reproduce each entry point separately."""
    w = witness_pairs()
    full, _ = cmo.dna_matrices(1, 2, 0)
    rc, pc = reflib.cm(full), port.cm(full)
    a, b = w["R3"]
    assert reflib.cost_affine(rc, a, b) == port.cost_affine(pc, a, b) == 5
    assert reflib.align_affine(rc, a, b)[0] == port.align_affine(pc, a, b)[0] == 4
