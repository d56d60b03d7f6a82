"""SeqCS.DOS policy inside the C ABI (poy_dos_distance / poy_dos_median, poy5_b200/csrc/dos.cu) against the Python
mirror of the same policy and against the CPU checker: empty sequences, either argument order, affine and linear
cost models."""
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from poy5_b200 import synth
from tests.oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


def _batch(seed):
    rng = np.random.default_rng(seed)
    seqs, ia, ib = synth.pair_batch(seed, 40, 120, frac_decorated=0.4, jitter=0.4)
    seqs = list(seqs)
    ia = list(ia); ib = list(ib)
    for q in range(0, 40, 2):                 # either order
        ia[q], ib[q] = ib[q], ia[q]
    e1 = len(seqs); seqs.append(np.array([16], np.uint8))                 # empty: only the leading gap
    e2 = len(seqs); seqs.append(np.array([16, 16, 16], np.uint8))         # empty: gaps only
    ia += [e1, 3, e2, e1]; ib += [5, e2, 7, e2]
    return seqs, np.asarray(ia, np.int32), np.asarray(ib, np.int32)


@pytest.mark.parametrize("go", [3, None])
def test_dos_distance_cabi(ctx, port, go):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import DOS, Heuristic, dos_distance
    t2d = Two_D.of_transformations_and_gaps(1, 2, go)
    full, orig = cmo.dna_matrices(1, 2, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    seqs, ia, ib = _batch(41)
    pool = pb.Pool(ctx, seqs)
    got = dos_distance(ctx, h, pool, ia, ib, missing_distance=7)
    assert np.array_equal(got, DOS.distance(ctx, h, pool, ia, ib, missing_distance=7))
    ob = OracleBackend(port, full, orig)
    ref = ob.distance([(seqs[a], seqs[b]) for a, b in zip(ia, ib)])
    nonempty = [p for p in range(len(ia)) if not (ob._empty(seqs[ia[p]]) or ob._empty(seqs[ib[p]]))]
    assert [int(got[p]) for p in nonempty] == [ref[p] for p in nonempty]
    assert all(int(got[p]) == 7 for p in range(len(ia)) if p not in nonempty)
    pool.close()


def test_dos_median_cabi(ctx, port):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import DOS, Heuristic, dos_median
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    full, orig = cmo.dna_matrices(1, 1, 3)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    seqs, ia, ib = _batch(43)
    pool = pb.Pool(ctx, seqs)
    med, cost = dos_median(ctx, h, pool, ia, ib)
    med2, cost2 = DOS.median_cost(ctx, h, pool, ia, ib)
    assert np.array_equal(cost, cost2)
    ref = OracleBackend(port, full, orig).median([(seqs[a], seqs[b]) for a, b in zip(ia, ib)])
    for p in range(len(ia)):
        assert np.array_equal(med[p], med2[p]), p
        assert np.array_equal(med[p], ref[p][0]) and int(cost[p]) == ref[p][1], p
    pool.close()
