"""world_size-2 gloo test of the N>1 host logic: LPT sharding of a candidate batch + the packed MIN
all-reduce that replaces the reference's MPI/Parmap gather.  The per-candidate costs come from the
CPU oracle here (no GPU in this test); on the GPU box the same code path runs under NCCL."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from poy5_b200 import shard, synth


def _worker(rank, world, port_file, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_file)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.port import Port
    from oracle import cost_matrix_oracle as cmo
    P = Port()
    pc = P.cm(cmo.dna_matrices(1, 1, 3)[0])
    seqs, ia, ib = synth.pair_batch(21, 40, 60, frac_decorated=0.3, jitter=0.5)
    lens = np.array([len(s) for s in seqs])
    parts = shard.lpt_partition(synth.cells(lens, ia, ib), world)
    mine = parts[rank]
    costs = [P.cost_affine(pc, seqs[ia[p]], seqs[ib[p]]) for p in mine]
    best = shard.allreduce_best(shard.pack_best(costs, mine))
    q.put((rank, best, [int(x) for x in mine]))
    dist.destroy_process_group()


def test_two_rank_min_reduce():
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank sees the same global best; shards are disjoint and cover the batch
    assert res[0][1] == res[1][1]
    allidx = sorted(res[0][2] + res[1][2])
    assert allidx == list(range(40))
    from oracle.port import Port
    from oracle import cost_matrix_oracle as cmo
    P = Port(); pc = P.cm(cmo.dna_matrices(1, 1, 3)[0])
    seqs, ia, ib = synth.pair_batch(21, 40, 60, frac_decorated=0.3, jitter=0.5)
    costs = np.array([P.cost_affine(pc, seqs[ia[p]], seqs[ib[p]]) for p in range(40)])
    cost, idx = shard.unpack_best(res[0][1])
    assert cost == costs.min() and idx == int(np.flatnonzero(costs == costs.min())[0])


def test_lpt_balance():
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1000, size=500)
    parts = shard.lpt_partition(w, 8)
    loads = np.array([w[p].sum() for p in parts])
    assert loads.max() - loads.min() <= w.max()
    assert sorted(np.concatenate(parts).tolist()) == list(range(500))
