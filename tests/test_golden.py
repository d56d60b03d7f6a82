"""Golden vectors generated from the unmodified reference C (tests/golden/make_golden.py):
the oracle port (CPU) and the CUDA path (GPU) must both reproduce them bit for bit."""
import os
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from tests.helpers import oracle_align

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "affine_golden.npz"))
REG = sorted({k.split("_")[0] for k in G.files})


def load(r):
    data, off = G[r + "_data"], G[r + "_off"]
    seqs = [data[off[s]:off[s + 1]] for s in range(len(off) - 1)]
    blens = G[r + "_blens"].reshape(-1, 4)
    starts = np.concatenate([[0], np.cumsum(G[r + "_blens"])])
    blob = G[r + "_blob"]
    outs = [[blob[starts[4 * p + q]:starts[4 * p + q + 1]] for q in range(4)] for p in range(len(blens))]
    return seqs, G[r + "_ia"], G[r + "_ib"], G[r + "_cost"], G[r + "_acost"], outs, tuple(int(x) for x in G[r + "_regime"])


@pytest.mark.parametrize("r", REG)
def test_port_reproduces_reference(port, r):
    seqs, ia, ib, cost, acost, outs, (s_, g_, go) = load(r)
    full, _ = cmo.dna_matrices(s_, g_, go)
    pc = port.cm(full)
    for p in range(len(ia)):
        a, b = seqs[ia[p]], seqs[ib[p]]
        assert port.cost_affine(pc, a, b) == cost[p]
        sw = int(len(a) > len(b))
        si, sj = (b, a) if sw else (a, b)
        res = port.align_affine(pc, si, sj, sw)
        assert res[0] == acost[p]
        for x, y in zip(res[1:], outs[p]):
            assert np.array_equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("r", REG)
def test_cuda_reproduces_reference(ctx, r):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    seqs, ia, ib, cost, acost, outs, (s_, g_, go) = load(r)
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(s_, g_, go).full)
    pool = pb.Pool(ctx, seqs)
    got = Align.cost_2(ctx, cm, pool, ia, ib)
    assert np.array_equal(got, cost)
    res = Align.align_affine_3(ctx, cm, pool, ia, ib)
    assert np.array_equal(res["cost"], acost)
    for p in range(len(ia)):
        sw = res["swaped"][p]
        med, mwg, ri, rj = outs[p]
        assert np.array_equal(res["median"][p], med) and np.array_equal(res["medianwg"][p], mwg)
        ra, rb = (rj, ri) if sw else (ri, rj)
        assert np.array_equal(res["res_a"][p], ra) and np.array_equal(res["res_b"][p], rb)
    cm.close(); pool.close()


# ---- pins at the BASELINE lengths (2 kb, 10 kb), generated from libpoyref_long.so by make_golden_long.py ----
GL = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "affine_long_golden.npz"))
REGL = sorted({k.split("_")[0] for k in GL.files})


def _digest(arrs):
    import hashlib
    h = hashlib.sha1()
    for x in arrs:
        h.update(np.ascontiguousarray(x, np.uint8).tobytes()); h.update(b"|")
    return np.frombuffer(h.digest(), np.uint8)


def load_long(r):
    data, off = GL[r + "_data"], GL[r + "_off"]
    seqs = [data[off[s]:off[s + 1]] for s in range(len(off) - 1)]
    return seqs, GL[r + "_cost"], GL[r + "_acost"], GL[r + "_lens4"], GL[r + "_sha"], tuple(int(x) for x in GL[r + "_regime"])


@pytest.mark.parametrize("r", REGL)
def test_port_reproduces_reference_long(port, r):
    """the oracle restatement against reference-held vectors at 2 kb (all pairs) and 10 kb (one pair per regime on CPU)"""
    seqs, cost, acost, lens4, sha, (s_, g_, go) = load_long(r)
    full, _ = cmo.dna_matrices(s_, g_, go)
    pc = port.cm(full)
    pairs = [p for p in range(len(cost)) if len(seqs[2 * p]) < 3000] + [len(cost) - 3]
    for p in pairs:
        a, b = seqs[2 * p], seqs[2 * p + 1]
        assert port.cost_affine(pc, a, b) == cost[p]
        sw = int(len(a) > len(b))
        si, sj = (b, a) if sw else (a, b)
        res = port.align_affine(pc, si, sj, sw)
        assert res[0] == acost[p] and [len(x) for x in res[1:]] == list(lens4[p])
        assert np.array_equal(_digest(res[1:]), sha[p])


@pytest.mark.gpu
@pytest.mark.parametrize("r", REGL)
def test_cuda_reproduces_reference_long(ctx, r):
    """CUDA against reference-held vectors at every BASELINE length: multi-block cost-only path, band classes >= 1280,
    len1 + len2 > 16382, gap-bit / ambiguity symbols, unrelated sequences"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    seqs, cost, acost, lens4, sha, (s_, g_, go) = load_long(r)
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(s_, g_, go).full)
    pool = pb.Pool(ctx, seqs)
    n = len(cost)
    ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
    assert np.array_equal(Align.cost_2(ctx, cm, pool, ia, ib), cost)
    assert np.array_equal(Align.cost_2(ctx, cm, pool, ib[pool.lens[ia] != pool.lens[ib]], ia[pool.lens[ia] != pool.lens[ib]]),
                          cost[pool.lens[ia] != pool.lens[ib]])
    res = Align.align_affine_3(ctx, cm, pool, ia, ib)
    assert np.array_equal(res["cost"], acost)
    for p in range(n):
        sw = res["swaped"][p]
        ri, rj = (res["res_b"][p], res["res_a"][p]) if sw else (res["res_a"][p], res["res_b"][p])
        outs = [res["median"][p], res["medianwg"][p], ri, rj]
        assert [len(x) for x in outs] == list(lens4[p])
        assert np.array_equal(_digest(outs), sha[p]), (r, p)
    # the same pairs through the low-latency / probe / generic schedules must give the same bytes
    cm.close(); pool.close()
