"""Golden vectors from the unmodified reference for the linear-gap path and the column-wise helpers
(tests/golden/make_golden_linear.py): oracle port on CPU, CUDA path on GPU."""
import os
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linear_golden.npz"))
CASES = sorted({k.split("_")[0] for k in G.files})


def load(c):
    data, off = G[c + "_data"], G[c + "_off"]
    seqs = [data[off[s]:off[s + 1]] for s in range(len(off) - 1)]
    starts = np.concatenate([[0], np.cumsum(G[c + "_blens"])])
    blob = G[c + "_blob"]
    n = len(G[c + "_ia"])
    parts = [[blob[starts[6 * p + q]:starts[6 * p + q + 1]] for q in range(6)] for p in range(n)]
    s_, g_, go = (int(x) for x in G[c + "_regime"])
    return seqs, G[c + "_ia"], G[c + "_ib"], G[c + "_cost"], G[c + "_dw"], parts, G[c + "_ints"].reshape(n, 2), (s_, g_, None if go < 0 else go)


@pytest.mark.parametrize("c", CASES)
def test_port_reproduces_reference(port, c):
    seqs, ia, ib, cost, dws, parts, ints, (s_, g_, go) = load(c)
    pc = port.cm(cmo.dna_matrices(s_, g_, go)[0])
    for p in range(len(ia)):
        a, b = seqs[ia[p]], seqs[ib[p]]
        sw = int(len(a) > len(b))
        s1, s2 = (b, a) if sw else (a, b)
        if go is None:
            c_, r1, r2 = port.align_linear(pc, s1, s2, int(dws[p]), sw)
        else:
            c_, _, _, r1, r2 = port.align_affine(pc, s1, s2, sw)
        assert c_ == cost[p]
        got = [r1, r2, port.median_2(pc, r1, r2, 0), port.median_2(pc, r1, r2, 1), port.union(r1, r2), port.ancestor_2(pc, r1, r2)]
        for x, y in zip(got, parts[p]):
            assert np.array_equal(x, y)
        assert port.worst_2(pc, r1, r2) == ints[p, 0] and port.verify_2(pc, r1, r2) == ints[p, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES)
def test_cuda_reproduces_reference(ctx, c):
    import poy5_b200 as pb
    from poy5_b200 import sequence
    from poy5_b200.api import _ptr
    from poy5_b200.cost_matrix import Two_D
    seqs, ia, ib, cost, dws, parts, ints, (s_, g_, go) = load(c)
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(s_, g_, go).full)
    pool = pb.Pool(ctx, seqs)
    n = len(ia)
    la, lb = pool.lens[ia], pool.lens[ib]
    sw = (la > lb).astype(np.uint8)
    s1 = np.where(sw == 1, ib, ia).astype(np.int32); s2 = np.where(sw == 1, ia, ib).astype(np.int32)
    if go is None:
        caps = (la + lb).astype(np.int64)
        out_off = np.concatenate([[0], np.cumsum(caps[:-1])]).astype(np.int64)
        r1b = np.zeros(int(caps.sum()), np.uint8); r2b = np.zeros(int(caps.sum()), np.uint8)
        got_cost = np.zeros(n, np.int32); out_len = np.zeros(2 * n, np.int32)
        dwh = np.ascontiguousarray(dws, np.int32)
        ctx.check(ctx.L.poy_batch_align_linear(ctx.h, cm.h, pool.h, n, _ptr(s1), _ptr(s2), _ptr(dwh), _ptr(sw), _ptr(out_off),
                                               _ptr(got_cost), _ptr(r1b), _ptr(r2b), _ptr(out_len), None))
        ends = out_off + caps
        rows1 = [r1b[ends[p] - out_len[2 * p]:ends[p]] for p in range(n)]
        rows2 = [r2b[ends[p] - out_len[2 * p + 1]:ends[p]] for p in range(n)]
    else:
        r = sequence.Align.align_affine_3(ctx, cm, pool, ia, ib)
        got_cost = r["cost"]
        rows1 = [r["res_b"][p] if sw[p] else r["res_a"][p] for p in range(n)]
        rows2 = [r["res_a"][p] if sw[p] else r["res_b"][p] for p in range(n)]
    assert np.array_equal(got_cost, cost)
    m0 = sequence.median_2(ctx, cm, rows1, rows2, False); m1 = sequence.median_2(ctx, cm, rows1, rows2, True)
    un = sequence.union(ctx, rows1, rows2); anc = sequence.ancestor_2(ctx, cm, rows1, rows2)
    wo = sequence.aligned_cost(ctx, cm, rows1, rows2, True); ve = sequence.aligned_cost(ctx, cm, rows1, rows2, False)
    for p in range(n):
        for x, y in zip([rows1[p], rows2[p], m0[p], m1[p], un[p], anc[p]], parts[p]):
            assert np.array_equal(x, y), p
        assert wo[p] == ints[p, 0] and ve[p] == ints[p, 1]
    cm.close(); pool.close()
