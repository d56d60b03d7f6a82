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


# ---- SPR round over several loci with incremental medians (BASELINE configs #4 / #5) ----

def loci_taxa(seed, n, lengths):
    return [taxa(seed + 17 * k, n, L) for k, L in enumerate(lengths)]


def run_spr(backend, loci):
    """Wagner tree of locus 0, downpass cost, one SPR round over all loci, exact cost of the rearranged tree"""
    tree = treesearch.wagner_build(loci[0], backend)
    c0, _ = treesearch.downpass(tree, loci, backend)
    est, move, ncand, naln = treesearch.spr_round(tree, loci, backend, chunk=5)
    t2 = treesearch.apply_spr(tree, move)
    c1, _ = treesearch.downpass(t2, loci, backend)
    return tree.edges(), c0, est, move, ncand, naln, t2.edges(), c1


def test_spr_round_oracle(port):
    full, orig = cmo.dna_matrices(1, 1, 3)
    loci = loci_taxa(7, 7, (50, 70))
    b = OracleBackend(port, full, orig)
    r = run_spr(b, loci)
    n = 7
    assert len(r[6]) == 2 * n - 3 and sorted(x for e in r[6] for x in e if x < n) == sorted(
        x for e in r[0] for x in e if x < n)
    assert r[4] > 20 and r[1] > 0 and r[7] > 0
    # downpass == sum over loci of the single-locus tree cost rooted on the same edge
    tree = treesearch.wagner_build(loci[0], b)
    assert r[1] == sum(treesearch.tree_cost(tree, ls, b) for ls in loci)
    # chunking does not change the neighbourhood or its best estimate
    est2, move2, ncand2, _ = treesearch.spr_round(tree, loci, b, chunk=1000)
    assert (est2, move2, ncand2) == (r[2], r[3], r[4])
    # the estimate of the winning move is what the driver says it is: distance(clade, edge median) + both sides
    dms = treesearch.all_directions(tree, loci, b)
    est3, move3, _, _ = treesearch.spr_round(tree, loci, b, dms=dms, prunings=[r[3][0]])
    assert (est3, move3) == (r[2], r[3])


@pytest.mark.gpu
def test_spr_gpu_matches_oracle_replay(ctx, port):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    full, orig = cmo.dna_matrices(1, 1, 3)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    loci = loci_taxa(23, 8, (120, 90, 150))
    got = run_spr(treesearch.GpuBackend(ctx, h), loci)
    ref = run_spr(OracleBackend(port, full, orig), loci)
    assert got == ref


def _spr_worker(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.port import Port
    full, orig = cmo.dna_matrices(1, 1, 3)
    b = treesearch.ShardedBackend(OracleBackend(Port(), full, orig), min_shard_medians=4)
    r = run_spr(b, loci_taxa(9, 6, (40, 55)))
    q.put((rank, r[1:6], r[7]))
    dist.destroy_process_group()


def test_sharded_spr_and_medians_gloo(port):
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    pn = 31700 + os.getpid() % 2000
    procs = [ctxm.Process(target=_spr_worker, args=(r, world, pn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, orig = cmo.dna_matrices(1, 1, 3)
    single = run_spr(OracleBackend(port, full, orig), loci_taxa(9, 6, (40, 55)))
    for r in res:
        assert (r[1], r[2]) == (single[1:6], single[7])


# ---- SPR neighbourhood strong-scaled by pruning (the swap-evaluation workload of bench.py) ----

def _spr_sharded_worker(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.port import Port
    full, orig = cmo.dna_matrices(1, 1, 3)
    b = OracleBackend(Port(), full, orig)
    loci = loci_taxa(9, 7, (40, 55))
    tree = treesearch.wagner_build(loci[0], b)
    dms = treesearch.all_directions(tree, loci, b)
    pr = treesearch.spr_prunings(tree, 7)
    r = treesearch.spr_round_sharded(tree, loci, b, dms, pr, chunk=3, rank=rank, world=world)
    q.put((rank, r))
    dist.destroy_process_group()


def test_spr_sharded_by_pruning_gloo(port):
    """world_size 2 over gloo: prunings dealt to the ranks, best candidate reduced with MIN all-reduces; every rank
    must report exactly what the single-process round reports (estimate, move, counters)"""
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    pn = 33700 + os.getpid() % 2000
    procs = [ctxm.Process(target=_spr_sharded_worker, args=(r, world, pn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, orig = cmo.dna_matrices(1, 1, 3)
    b = OracleBackend(port, full, orig)
    loci = loci_taxa(9, 7, (40, 55))
    tree = treesearch.wagner_build(loci[0], b)
    single = treesearch.spr_round(tree, loci, b, chunk=1000)
    for _, r in res:
        assert r == single


def test_spr_lanes_match_single(port):
    """several concurrent lanes on one rank (own backend each, prunings dealt like ranks): same result as one lane, with
    and without the edge medians merged into the level batches"""
    from oracle.port import Port
    full, orig = cmo.dna_matrices(1, 1, 3)
    b = OracleBackend(port, full, orig)
    loci = loci_taxa(9, 7, (40, 55))
    tree = treesearch.wagner_build(loci[0], b)
    dms = treesearch.all_directions(tree, loci, b)
    pr = treesearch.spr_prunings(tree, 7)
    single = treesearch.spr_round(tree, loci, b, dms=dms, prunings=pr, chunk=1000)
    merged = treesearch.spr_round(tree, loci, b, dms=dms, prunings=pr, chunk=4, merge_edges=True)
    assert merged == single
    lanes = [(OracleBackend(Port(), full, orig), loci, dms) for _ in range(3)]      # one checker scratch per thread
    got = treesearch.spr_round_sharded(tree, loci, b, dms, pr, chunk=2, rank=0, world=1, lanes=lanes)
    assert got == single


def test_vectorised_spr_round_matches_generic(port):
    """the id-array driver that runs on the node store (treesearch._spr_round_ids) issues the same medians / distances as
    the generic per-object driver: same estimate, move, counters, `where`; chunking does not matter"""
    from tests.oracle_backend import OracleStoreBackend
    full, orig = cmo.dna_matrices(1, 1, 3)
    host_loci = loci_taxa(17, 9, (40, 55, 48))
    ob = OracleBackend(port, full, orig)
    tree = treesearch.wagner_build(host_loci[0], ob)
    dms_h = treesearch.all_directions(tree, host_loci, ob)
    pr = treesearch.spr_prunings(tree, 9)
    w0 = []
    generic = treesearch.spr_round(tree, host_loci, ob, dms=dms_h, prunings=pr, chunk=1000, where=w0)
    sb = OracleStoreBackend(port, full, orig)
    loci = [sb.put(ls) for ls in host_loci]
    dms = treesearch.all_directions(tree, loci, sb)
    for chunk, merge in ((1000, False), (3, False), (4, True)):
        w1 = []
        got = treesearch.spr_round(tree, loci, sb, dms=dms, prunings=pr, chunk=chunk, where=w1, merge_edges=merge)
        assert got == generic and w1 == w0


@pytest.mark.gpu
@pytest.mark.parametrize("go", [3, None])
def test_store_backend_matches_oracle_replay(ctx, port, go):
    """device-resident node store: the whole SPR replay (Wagner build, downpass, SPR round with released temporaries,
    exact rejoin) through StoreBackend handles equals the CPU checker's replay; sampled medians equal byte for byte"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 1, go)
    full, orig = cmo.dna_matrices(1, 1, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    host_loci = loci_taxa(23, 8, (120, 90, 150))
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=1 << 16, cap_seqs=64)        # tiny: exercises the store's growth
    loci = [sb.put(ls) for ls in host_loci]
    got = run_spr(sb, loci)
    ref = run_spr(OracleBackend(port, full, orig), host_loci)
    assert got == ref
    ob = OracleBackend(port, full, orig)
    pairs = [(loci[0][0], loci[0][1]), (loci[1][2], loci[1][5]), (loci[2][7], loci[2][3])]
    res = sb.median(pairs)
    exp = ob.median([(host_loci[0][0], host_loci[0][1]), (host_loci[1][2], host_loci[1][5]), (host_loci[2][7], host_loci[2][3])])
    for (node, c2), (seq, oc), fetched in zip(res, exp, sb.fetch([r_[0] for r_ in res])):
        assert c2 == oc and len(node) == len(seq) and np.array_equal(fetched, seq)
    # medians of medians (inputs that never left the device), and an empty child
    empty = sb.put([np.array([16], np.uint8)])[0]
    res2 = sb.median([(res[0][0], res[1][0]), (empty, res[2][0]), (res[1][0], empty)])
    exp2 = ob.median([(exp[0][0], exp[1][0]), (np.array([16], np.uint8), exp[2][0]), (exp[1][0], np.array([16], np.uint8))])
    for (node, c2), (seq, oc), fetched in zip(res2, exp2, sb.fetch([r_[0] for r_ in res2])):
        assert c2 == oc and np.array_equal(fetched, seq)
    assert sb.distance([(res2[0][0], loci[0][4]), (empty, loci[0][1])]) == ob.distance([(exp2[0][0], host_loci[0][4]), (np.array([16], np.uint8), host_loci[0][1])])
    sb.close()


# ---- TBR round over several loci with incremental medians ----

def run_tbr_multi(backend, loci, chunk=3):
    tree = treesearch.wagner_build(loci[0], backend)
    c0, _ = treesearch.downpass(tree, loci, backend)
    est, move, ncand, naln = treesearch.tbr_round_multi(tree, loci, backend, chunk=chunk)
    t2 = treesearch.apply_tbr_multi(tree, move)
    c1, _ = treesearch.downpass(t2, loci, backend)
    return tree.edges(), c0, est, move, ncand, naln, t2.edges(), c1


def _is_binary_tree(edges, n):
    deg = {}
    for a, b in edges:
        deg[a] = deg.get(a, 0) + 1; deg[b] = deg.get(b, 0) + 1
    t = treesearch.Tree()
    for a, b in edges:
        t.add_edge(a, b)
    return (len(edges) == 2 * n - 3 and all(deg[x] == (1 if x < n else 3) for x in deg)
            and sorted(x for x in deg if x < n) == list(range(n)) and len(t.component(0, None)) == len(deg))


def test_tbr_multi_oracle(port):
    full, orig = cmo.dna_matrices(1, 1, 3)
    n = 7
    loci = loci_taxa(13, n, (45, 60))
    b = OracleBackend(port, full, orig)
    r = run_tbr_multi(b, loci)
    assert _is_binary_tree(r[0], n) and _is_binary_tree(r[6], n)
    tree = treesearch.wagner_build(loci[0], b)
    # the TBR neighbourhood contains the SPR neighbourhood (re-rooting one side only) and the unbroken tree is left out
    _, _, nspr, _ = treesearch.spr_round(tree, loci, b)
    assert r[4] > nspr / 2
    # chunking does not change the neighbourhood or its best estimate
    est2, move2, ncand2, _ = treesearch.tbr_round_multi(tree, loci, b, chunk=100)
    assert (est2, move2, ncand2) == (r[2], r[3], r[4])
    # every move of a small neighbourhood yields a valid tree
    for brk in tree.edges()[:4]:
        sides = [treesearch._side_plan(tree, s, t) for s, t in (brk, brk[::-1])]
        names = [[None] if pl is None else ["m"] + [(pl[3][c][0], c) for lvl in pl[2] for c in lvl] for pl in sides]
        for ea in names[0]:
            for eb in names[1]:
                assert _is_binary_tree(treesearch.apply_tbr_multi(tree, (brk, ea, eb)).edges(), n), (brk, ea, eb)
    # re-inserting at (merged, merged) restores the tree
    brk = [e for e in tree.edges() if e[0] >= n and e[1] >= n][0]
    assert treesearch.apply_tbr_multi(tree, (brk, "m", "m")).edges() == tree.edges()


@pytest.mark.gpu
def test_tbr_multi_store_backend_matches_oracle_replay(ctx, port):
    """TBR round over several loci with incremental medians on both sides of the break, through the device-resident node
    store: identical trees, estimate, move, counters and costs to the CPU checker's replay"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    full, orig = cmo.dna_matrices(1, 1, 3)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    host_loci = loci_taxa(29, 8, (110, 140))
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=1 << 18, cap_seqs=256)
    loci = [sb.put(ls) for ls in host_loci]
    got = run_tbr_multi(sb, loci)
    ref = run_tbr_multi(OracleBackend(port, full, orig), host_loci)
    assert got == ref
    sb.close()


# ---- single assignment (pre-order pass after the downpass) ----

def test_single_assignment_oracle(port):
    full, orig = cmo.dna_matrices(1, 1, 3)
    n = 8
    loci = loci_taxa(31, n, (50, 64))
    b = OracleBackend(port, full, orig)
    tree = treesearch.wagner_build(loci[0], b)
    cost, single_cost, singles = treesearch.single_assignment(tree, loci, b)
    assert cost == treesearch.downpass(tree, loci, b)[0] and single_cost > 0
    pf = port.cm(full)
    root = tree.edges()[0]
    for l, ls in enumerate(loci):
        assert set(singles[l]) == set(tree.adj) | {"root"}
        for x, s in singles[l].items():
            assert s[0] == 16 and all(bin(int(v)).count("1") == 1 for v in s[1:]), (l, x)     # one base per position
        for t in range(n):
            assert np.array_equal(singles[l][t], ls[t])            # an unambiguous leaf is its own single assignment
    # the single-assignment cost is the sum of the parent/child distances of the assigned sequences
    tot = 0
    for l in range(len(loci)):
        parent = {root[0]: "root", root[1]: "root"}
        order = [root[0], root[1]]
        for x in order:
            up = (root[1] if x == root[0] else root[0]) if parent[x] == "root" else parent[x]
            for y in tree.adj[x]:
                if y != up:
                    parent[y] = x; order.append(y)
        for x in order:
            ps, xs = singles[l][parent[x]], singles[l][x]
            tot += 0 if np.array_equal(ps, xs) else int(port.cost_affine(pf, ps, xs))
    assert tot == single_cost


@pytest.mark.gpu
def test_baseline_config_workloads_small(ctx):
    """poy5_b200.workloads (the configs[0] / [2] / [4] sub-records of bench.py) at toy sizes: every recorded sample
    replays on the CPU checker without a mismatch"""
    from poy5_b200 import workloads
    from tests.oracle_backend import replay_sample, replay_triplets
    r1, s1 = workloads.config1(ctx, taxa=7, L=120, check=8)
    assert r1["tbr_candidates"] > 20 and r1["cost_after_tbr"] <= r1["build_cost"] + 50
    p1 = replay_sample(s1[0], s1[1], (1, 1, 3))
    assert p1["checked"] > 0 and p1["mismatches"] == 0
    r3, s3 = workloads.config3(ctx, triplets=40, L=150, chunk=20, check=5)
    p3 = replay_triplets(s3, (1, 1, 3))
    assert p3["checked"] == 5 and p3["mismatches"] == 0 and r3["sum_cost"] > 0
    r5, s5 = workloads.config5(ctx, taxa=12, L=300, prunings=3, check=6)
    p5 = replay_sample(s5[0], s5[1], (1, 1, 3))
    assert r5["spr_candidates"] > 10 and p5["checked"] > 0 and p5["mismatches"] == 0
    from tests.oracle_backend import replay_newkk
    rn, sn = workloads.newkk(ctx, pairs=40, L=200, check=6)
    pn = replay_newkk(sn, (1, 1, 3))
    assert pn["checked"] == 6 and pn["mismatches"] == 0 and rn["alignments_per_s"] > 0
