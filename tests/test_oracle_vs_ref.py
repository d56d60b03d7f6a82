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


@pytest.mark.parametrize("tcm", [(1, 1), (2, 1), (1, 2), (3, 2)])
def test_linear_and_helpers(reflib, port, tcm):
    """linear-gap align_2d (full plane and Ukkonen band) + column-wise helpers vs the compiled reference"""
    rng = np.random.default_rng(tcm[0] * 7 + tcm[1])
    for m in cmo.dna_matrices(tcm[0], tcm[1], None):
        rc, pc = reflib.cm(m), port.cm(m)
        modes = [0, 0]
        for it in range(400):
            L = int(rng.integers(0, 140)) if it % 5 else int(rng.integers(0, 12))
            anc = synth.random_seq(rng, L)
            a, b = synth.evolve(rng, anc, 0.15, 0.05), synth.evolve(rng, anc, 0.15, 0.05)
            if it % 3 == 0:
                a, b = synth.decorate(rng, a, 0.1, 0.1), synth.decorate(rng, b, 0.1, 0.1)
            if it % 7 == 0:
                b = synth.random_seq(rng, int(rng.integers(0, 200)))
            a, b = synth.with_gap(a), synth.with_gap(b)
            s1, s2 = (a, b) if len(a) <= len(b) else (b, a)
            dw, sw = int(rng.integers(0, 30)), int(rng.integers(0, 2))
            r1 = reflib.align_linear(rc, s1, s2, dw, sw)
            c2, st = port.cost_linear(pc, s1, s2, dw, with_stats=True)
            r2 = port.align_linear(pc, s1, s2, dw, sw)
            modes[1 if st.iterations > 0 else 0] += 1
            assert r1[0] == r2[0] == c2 and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
            x, y = r1[1], r1[2]
            for wg in (0, 1):
                assert np.array_equal(reflib.median_2(rc, x, y, wg), port.median_2(pc, x, y, wg))
            assert np.array_equal(reflib.union(x, y), port.union(x, y))
            assert reflib.worst_2(rc, x, y) == port.worst_2(pc, x, y) and reflib.verify_2(rc, x, y) == port.verify_2(pc, x, y)
            assert np.array_equal(reflib.ancestor_2(rc, x, y), port.ancestor_2(pc, x, y))
        assert modes[0] > 20 and modes[1] > 20   # both the full plane and the Ukkonen band were exercised


def test_affine_helpers(reflib, port):
    rng = np.random.default_rng(4)
    for reg in ((1, 1, 3), (2, 1, 5), (1, 2, 0)):
        m = cmo.dna_matrices(*reg)[0]
        rc, pc = reflib.cm(m), port.cm(m)
        for it in range(200):
            anc = synth.random_seq(rng, int(rng.integers(0, 80)))
            a = synth.with_gap(synth.decorate(rng, synth.evolve(rng, anc, 0.15, 0.06), 0.1, 0.1))
            b = synth.with_gap(synth.decorate(rng, synth.evolve(rng, anc, 0.15, 0.06), 0.1, 0.1))
            si, sj = (a, b) if len(a) <= len(b) else (b, a)
            r = reflib.align_affine(rc, si, sj, 0)
            x, y = r[3], r[4]
            assert reflib.worst_2(rc, x, y) == port.worst_2(pc, x, y) and reflib.verify_2(rc, x, y) == port.verify_2(pc, x, y)
            assert np.array_equal(reflib.ancestor_2(rc, x, y), port.ancestor_2(pc, x, y))
            for wg in (0, 1):
                assert np.array_equal(reflib.median_2(rc, x, y, wg), port.median_2(pc, x, y, wg))


def test_arbitrary_alphabet_property(reflib, port):
    """hypothesis: sequences over the WHOLE bitset alphabet 1..31 (any ambiguity, gap bits anywhere, runs of pure
    gaps), any length 0..48 on either side, all regimes: both entry points of the restatement equal the reference"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    cms = {}
    for rname, (s_, g_, go) in REGIMES.items():
        full, _ = cmo.dna_matrices(s_, g_, go)
        cms[rname] = (reflib.cm(full), port.cm(full))
    code = st.integers(min_value=1, max_value=31)
    seq = st.lists(code, min_size=0, max_size=48)

    @settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(seq, seq, st.sampled_from(sorted(cms)), st.integers(0, 1))
    def check(a, b, rname, sw):
        rc, pc = cms[rname]
        a = np.array([16] + a, np.uint8); b = np.array([16] + b, np.uint8)
        assert reflib.cost_affine(rc, a, b) == port.cost_affine(pc, a, b)
        si, sj = (a, b) if len(a) <= len(b) else (b, a)
        r1, r2 = reflib.align_affine(rc, si, sj, sw), port.align_affine(pc, si, sj, sw)
        assert r1[0] == r2[0] and all(np.array_equal(x, y) for x, y in zip(r1[1:], r2[1:]))
    check()
