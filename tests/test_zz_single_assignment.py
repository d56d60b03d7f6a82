"""Tree-level single assignment (treesearch.single_assignment: downpass + pre-order DOS.to_single) through the GPU
backend vs the replay of the same call sequence on the CPU checker."""
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from poy5_b200 import treesearch
from tests.oracle_backend import OracleBackend
from tests.test_treesearch import loci_taxa

pytestmark = pytest.mark.gpu


def test_single_assignment_gpu_matches_oracle_replay(ctx, port):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    full, orig = cmo.dna_matrices(1, 1, 3)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    loci = loci_taxa(57, 9, (110, 140))
    gb, ob = treesearch.GpuBackend(ctx, h), OracleBackend(port, full, orig)
    tree = treesearch.wagner_build(loci[0], ob)
    got = treesearch.single_assignment(tree, loci, gb)
    ref = treesearch.single_assignment(tree, loci, ob)
    assert got[:2] == ref[:2]
    for sg, sr in zip(got[2], ref[2]):
        assert set(sg) == set(sr)
        for x in sg:
            assert np.array_equal(sg[x], sr[x]), x
