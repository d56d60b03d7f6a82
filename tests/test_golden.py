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
