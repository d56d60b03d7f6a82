"""Tree-level driver: the GPU backend and an oracle-backed replay must build the same Wagner
tree, find the same TBR move and report the same tree costs; the N>1 sharding of candidate batches is
exercised with gloo on CPU."""
import os
import numpy as np
import pytest
import torch.multiprocessing as mp
from oracle import cost_matrix_oracle as cmo
from poy5_b200 import synth, treesearch
from tests.oracle_backend import OracleBackend


def taxa(seed, n, L):
    """n taxa evolved along a random bifurcating history (5 % substitutions, 0.5 % indels per branch)"""
    rng = np.random.default_rng(seed)
    pool = [synth.random_seq(rng, L)]
    while len(pool) < n:
        p = pool.pop(int(rng.integers(0, len(pool))))
        pool += [synth.evolve(rng, p, 0.05, 0.005), synth.evolve(rng, p, 0.05, 0.005)]
    return [synth.with_gap(s) for s in pool[:n]]


def run(backend, leaves):
    tree = treesearch.wagner_build(leaves, backend)
    c0 = treesearch.tree_cost(tree, leaves, backend)
    est, move, ncand = treesearch.tbr_round(tree, leaves, backend)
    t2 = treesearch.apply_tbr(tree, move)
    c1 = treesearch.tree_cost(t2, leaves, backend)
    return tree.edges(), c0, est, move, ncand, t2.edges(), c1


def test_oracle_replay_is_self_consistent(port):
    full, orig = cmo.dna_matrices(1, 1, 3)
    leaves = taxa(3, 7, 60)
    r = run(OracleBackend(port, full, orig), leaves)
    assert len(r[0]) == 2 * 7 - 3 and len(r[5]) == 2 * 7 - 3     # unrooted binary trees
    assert r[4] > 50 and r[1] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("go", [3, None])
def test_gpu_matches_oracle_replay(ctx, port, go):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 1, go)
    full, orig = cmo.dna_matrices(1, 1, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    leaves = taxa(11, 9, 150)
    got = run(treesearch.GpuBackend(ctx, h), leaves)
    ref = run(OracleBackend(port, full, orig), leaves)
    assert got == ref


def _worker(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.port import Port
    full, orig = cmo.dna_matrices(1, 1, 3)
    leaves = taxa(5, 6, 50)
    b = treesearch.ShardedBackend(OracleBackend(Port(), full, orig))
    r = run(b, leaves)
    q.put((rank, r[1], r[2], r[6]))
    dist.destroy_process_group()


def test_sharded_candidates_gloo(port):
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    pn = 29700 + os.getpid() % 2000
    procs = [ctxm.Process(target=_worker, args=(r, world, pn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, orig = cmo.dna_matrices(1, 1, 3)
    single = run(OracleBackend(port, full, orig), taxa(5, 6, 50))
    for r in res:
        assert (r[1], r[2], r[3]) == (single[1], single[2], single[6])
