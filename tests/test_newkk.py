"""Sequence.NewkkAlign (src/newkkonen.c, affine entry point): the plain-C restatement oracle/newkk_oracle.c against
the compiled reference (build container only) and against the committed reference-generated goldens; the CUDA path
(poy_batch_newkk_align through the C ABI) against both (GPU box)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import cost_matrix_oracle as cmo
from poy5_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "newkk_golden.npz")


def digest(arrs):
    h = hashlib.sha1()
    for x in arrs:
        h.update(np.ascontiguousarray(x, np.uint8).tobytes()); h.update(b"|")
    return np.frombuffer(h.digest(), np.uint8)


def golden_sets():
    g = np.load(GOLD)
    for rname in ("R1", "R2", "R3", "R5"):
        data, off = g[rname + "_data"], g[rname + "_off"]
        seqs = [data[off[s]:off[s + 1]] for s in range(len(off) - 1)]
        yield rname, tuple(int(x) for x in g[rname + "_regime"]), seqs, g[rname + "_cost"], g[rname + "_k"], g[rname + "_lens2"], g[rname + "_sha"]


def random_pairs(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        la = int(rng.integers(0, 90)) if t % 3 else int(rng.integers(0, 4))
        lb = int(rng.integers(0, 220)) if t % 5 else int(rng.integers(100, 450))
        a, b = synth.random_seq(rng, la), synth.random_seq(rng, lb)
        if t % 2:
            a = synth.decorate(rng, a, 0.15, 0.15)
        if (t // 2) % 2:
            b = synth.decorate(rng, b, 0.15, 0.15)
        if t % 4 == 0 and la > 5:       # related pair
            b = synth.evolve(rng, a & 15, 0.1, 0.03)
        out.append((synth.with_gap(a), synth.with_gap(b)))
    return out


def test_port_matches_golden(port):
    """the restatement reproduces the reference-generated vectors (cost, final k, both aligned rows)"""
    checked = 0
    for rname, reg, seqs, cost, kk, lens2, sha in golden_sets():
        full, _ = cmo.dna_matrices(*reg)
        pc = port.cm(full)
        for p in range(len(cost)):
            a, b = seqs[2 * p], seqs[2 * p + 1]
            sw = int(len(a) > len(b))
            s1, s2 = (b, a) if sw else (a, b)
            c, r1, r2, st = port.newkk_align(pc, s1, s2, sw, with_stats=True)
            assert c == cost[p], (rname, p)
            if kk[p] >= 0:
                assert st.final_k == kk[p], (rname, p)
            assert [len(r1), len(r2)] == list(lens2[p]) and np.array_equal(digest([r1, r2]), sha[p]), (rname, p)
            checked += 1
    assert checked == 4 * 46


def test_port_matches_reference(port):
    """build container only: restatement vs the unmodified src/newkkonen.c on random / ragged / decorated pairs, one
    newkkmat scratch reused across all calls (as NewkkAlign.default_ukkm is)"""
    from oracle import refbind
    if not refbind.available(True):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = refbind.RefLib(True)
    if not hasattr(R.lib, "ref_newkk_new"):
        pytest.skip("oracle/_ref predates the newkkonen driver")
    for reg in [(1, 1, 3), (2, 1, 5), (1, 2, 0), (3, 2, 4), (2, 2, 7)]:
        full, _ = cmo.dna_matrices(*reg)
        rc, pc = R.cm(full), port.cm(full)
        for a, b in random_pairs(sum(reg), 120):
            sw = int(len(a) > len(b))
            s1, s2 = (b, a) if sw else (a, b)
            c, r1, r2, k = R.newkk_align(rc, s1, s2, 1, sw)
            oc, o1, o2, st = port.newkk_align(pc, s1, s2, sw, with_stats=True)
            assert c == oc and np.array_equal(r1, o1) and np.array_equal(r2, o2)
            if not len(s1) * 100 < len(s2):
                assert k == st.final_k


def test_reference_non_affine_entry_point_is_broken():
    """why the product offers the affine entry point only: newkkonen_CAML_algn returns 0 and its traceback raises"""
    from oracle import refbind
    if not refbind.available(True):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = refbind.RefLib(True)
    if not hasattr(R.lib, "ref_newkk_new"):
        pytest.skip("oracle/_ref predates the newkkonen driver")
    full, _ = cmo.dna_matrices(1, 1, None)
    rc = R.cm(full)
    rng = np.random.default_rng(2)
    a = synth.with_gap(synth.random_seq(rng, 40)); b = a.copy(); b[7] ^= 3
    assert R.newkk_cost(rc, a, b, 0, 0) == 0
    with pytest.raises(RuntimeError):
        R.newkk_align(rc, a, b, 0, 0)


# ---- CUDA path ------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_cuda_matches_golden(ctx):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import NewkkAlign
    for rname, reg, seqs, cost, kk, lens2, sha in golden_sets():
        cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(*reg).full)
        pool = pb.Pool(ctx, seqs)
        n = len(cost)
        ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
        r = NewkkAlign.align_2(ctx, cm, pool, ia, ib, stats=True)
        c_only = NewkkAlign.cost_2(ctx, cm, pool, ia, ib)
        assert np.array_equal(r["cost"], cost) and np.array_equal(c_only, cost), rname
        for p in range(n):
            a, b = seqs[2 * p], seqs[2 * p + 1]
            sw = len(a) > len(b)
            r1, r2 = (r["res_b"][p], r["res_a"][p]) if sw else (r["res_a"][p], r["res_b"][p])
            assert [len(r1), len(r2)] == list(lens2[p]) and np.array_equal(digest([r1, r2]), sha[p]), (rname, p)
            if kk[p] >= 0:
                assert r["stats"][p, 2] == kk[p]
            else:
                assert r["stats"][p, 3] == 1
        cm.close(); pool.close()


@pytest.mark.gpu
def test_cuda_matches_port(ctx, port):
    """seeded random pairs, incl. bands wider than the shared-memory planes (global-scratch path) when forced small"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import NewkkAlign
    for reg in [(1, 1, 3), (2, 1, 5), (3, 2, 4)]:
        full, _ = cmo.dna_matrices(*reg)
        pc = port.cm(full)
        cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(*reg).full)
        pairs = random_pairs(100 + sum(reg), 150)
        rng = np.random.default_rng(5)
        for L in (2500, 5200):      # unrelated long pairs: W reaches len1 + len2 - 1 > 4096 -> global planes
            pairs.append((synth.with_gap(synth.random_seq(rng, L)), synth.with_gap(synth.random_seq(rng, L - 300))))
        seqs = [x for ab in pairs for x in ab]
        pool = pb.Pool(ctx, seqs)
        n = len(pairs)
        ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
        r = NewkkAlign.align_2(ctx, cm, pool, ia, ib, stats=True)
        for p, (a, b) in enumerate(pairs):
            sw = int(len(a) > len(b))
            s1, s2 = (b, a) if sw else (a, b)
            oc, o1, o2, st = port.newkk_align(pc, s1, s2, sw, with_stats=True)
            ra, rb = (o2, o1) if sw else (o1, o2)
            assert oc == r["cost"][p], (reg, p)
            assert np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), (reg, p)
            if not len(s1) * 100 < len(s2):
                assert st.final_k == r["stats"][p, 2] and st.iterations == r["stats"][p, 0]
        cm.close(); pool.close()


@pytest.mark.gpu
def test_cuda_newkk_errors(ctx):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import NewkkAlign
    from poy5_b200._lib import PoyError
    cm_lin = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, None).full)
    pool = pb.Pool(ctx, [np.array([16, 1, 2], np.uint8), np.array([16, 1, 2, 4], np.uint8)])
    with pytest.raises(PoyError):
        NewkkAlign.cost_2(ctx, cm_lin, pool, [0], [1])
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, 3).full)
    cost = np.zeros(1, np.int32)
    one, zero = np.array([1], np.int32), np.array([0], np.int32)
    st = ctx.L.poy_batch_newkk_align(ctx.h, cm.h, pool.h, 1, one.ctypes.data, zero.ctypes.data, None, None, cost.ctypes.data, None, None,
                                     None, None)
    assert st == -4     # POY_ERR_ORDER: "newkkonen.newkk_algn, s1 len > s2 len"


@pytest.mark.gpu
@pytest.mark.parametrize("use_ukk", [False, True])
def test_dos_dist_2(ctx, port, use_ukk):
    """DOS.dist_2 (src/seqCS.ml:1181-1198): full_median_2 of (a, b) then cost_2 -- Sequence.NewkkAlign.cost_2 when
    use_ukk -- of n against it, vs the same composition on the CPU checker"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic, dist_2
    from tests.helpers import oracle_align
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    full, orig = cmo.dna_matrices(1, 1, 3)
    pf = port.cm(full)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    rng = np.random.default_rng(12)
    seqs = []
    for t in range(40):
        anc = synth.random_seq(rng, int(rng.integers(20, 200)))
        trio = [synth.with_gap(synth.decorate(rng, synth.evolve(rng, anc, 0.1, 0.03), 0.05, 0.05 * (t % 2))) for _ in range(3)]
        if t % 9 == 0:
            trio[1] = np.array([16], np.uint8)          # empty a: tmp = n itself
        seqs += trio
    pool = pb.Pool(ctx, seqs)
    idx = np.arange(0, len(seqs), 3, dtype=np.int32)
    got = dist_2(ctx, h, pool, idx, idx + 1, idx + 2, use_ukk=use_ukk)
    for q, p in enumerate(idx):
        n_, a, b = seqs[p], seqs[p + 1], seqs[p + 2]
        tmp = n_ if (a == 16).all() else oracle_align(port, pf, a, b)[1]
        if use_ukk:
            s1, s2 = (tmp, n_) if len(n_) > len(tmp) else (n_, tmp)
            want = port.newkk_align(pf, s1, s2, 0)[0]
        else:
            want = port.cost_affine(pf, n_, tmp)
        assert got[q] == want, (q, use_ukk)
    pool.close()
