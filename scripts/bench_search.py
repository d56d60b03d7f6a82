"""BASELINE configs #4 / #5: a tree-search round over dynamic-homology characters, all alignments batched.

  config #4   --taxa 200 --loci 5 --lmin 1000 --lmax 3000      (DO search round, candidates sharded over ranks)
  config #5   --taxa 1000 --loci 1 --lmin 10000 --lmax 10000   (downpass + SPR round at 1/2/4/8 GPUs)

Workload per run: a random (Yule) starting tree over the synthetic taxa, the all-direction medians of the tree
(level-synchronous, every locus in the same batches), the downpass cost, then one SPR neighbourhood
(poy5_b200.treesearch.spr_round: incremental medians after each break + one cost-only candidate batch per chunk of
prunings).  `--prunings K` evaluates an evenly spaced sample of K prunings instead of the whole neighbourhood
(the full config #5 neighbourhood is about 8e6 banded medians and 4e6 cost-only alignments of 10 kb).
Under torchrun every rank owns one GPU; candidate batches and wide median levels are sharded over the ranks
(strong scaling), the exchange is one all-reduce / all-gather per batch.  `--check S` replays a sample of S
candidates and S medians on the CPU checker (oracle/) and compares bit for bit.

    python scripts/bench_search.py --taxa 200 --loci 5 --prunings 64
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_search.py --taxa 1000 --lmin 10000 --lmax 10000 --prunings 32
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def make_taxa(seed, n, L):
    """n taxa evolved along a random bifurcating history (5 % substitutions, 0.5 % indels per branch)"""
    from poy5_b200 import synth
    rng = np.random.default_rng(seed)
    pool = [synth.random_seq(rng, L)]
    while len(pool) < n:
        p = pool.pop(int(rng.integers(0, len(pool))))
        pool += [synth.evolve(rng, p, 0.05, 0.005), synth.evolve(rng, p, 0.05, 0.005)]
    return [synth.with_gap(s) for s in pool[:n]]


def random_tree(seed, n):
    from poy5_b200.treesearch import Tree
    rng = np.random.default_rng(seed)
    t = Tree(); t.add_edge(0, 1)
    for leaf in range(2, n):
        edges = t.edges()
        u, v = edges[int(rng.integers(0, len(edges)))]
        w = max(max(t.adj) + 1, n)
        t.remove_edge(u, v); t.add_edge(u, w); t.add_edge(w, v); t.add_edge(w, leaf)
    return t


class Recorder:
    """keeps a sample of the (inputs, outputs) that went through the backend, for the CPU replay"""
    def __init__(self, b, every):
        self.b, self.every, self.med, self.dis, self.k = b, max(1, every), [], [], 0

    def median(self, pairs):
        r = self.b.median(pairs)
        for p, o in zip(pairs[::self.every], r[::self.every]):
            self.med.append((p, o))
        return r

    def distance(self, pairs):
        r = self.b.distance(pairs)
        for p, o in zip(pairs[::self.every], r[::self.every]):
            self.dis.append((p, o))
        return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=200)
    ap.add_argument("--loci", type=int, default=5)
    ap.add_argument("--lmin", type=int, default=1000)
    ap.add_argument("--lmax", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--prunings", type=int, default=0, help="sample size; 0 = the whole SPR neighbourhood")
    ap.add_argument("--chunk", type=int, default=32, help="prunings per candidate batch")
    ap.add_argument("--check", type=int, default=0, help="replay this many medians and candidates on the CPU checker")
    ap.add_argument("--full-median", action="store_true", help="read back the whole DOS.median record per node")
    ap.add_argument("--profile", action="store_true", help="cProfile of the SPR round to stderr")
    ap.add_argument("--tbr", action="store_true", help="TBR neighbourhood (both sides re-rooted) instead of SPR; --prunings samples breaks")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import poy5_b200 as pb
    from poy5_b200 import treesearch
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pb.Context(local)
    t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    rng = np.random.default_rng(a.seed)
    lens = [int(rng.integers(a.lmin, a.lmax + 1)) for _ in range(a.loci)]
    loci = [make_taxa(a.seed + 17 * k, a.taxa, L) for k, L in enumerate(lens)]
    tree = random_tree(a.seed, a.taxa)
    gb = treesearch.GpuBackend(ctx, h, lean=not a.full_median)
    back = treesearch.ShardedBackend(gb, device=torch.device("cuda", local), min_shard_medians=64 * world)
    rec = Recorder(back, 1) if a.check else back
    treesearch.downpass(random_tree(1, 6), [make_taxa(2, 6, 200)], back)         # warm-up

    def sync():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sync(); t0 = time.perf_counter()
    cost, _ = treesearch.downpass(tree, loci, rec)
    sync(); t1 = time.perf_counter()
    dms = treesearch.all_directions(tree, loci, rec)
    sync(); t2 = time.perf_counter()
    pr = treesearch.spr_prunings(tree, a.taxa)
    if a.prunings and a.prunings < len(pr):
        pr = [pr[i] for i in np.linspace(0, len(pr) - 1, a.prunings).astype(int)]
    m0, d0, c0 = gb.n_median, gb.n_distance, gb.cells_distance
    if a.check:
        rec.every = max(1, (len(pr) * 4 * a.taxa * a.loci) // max(1, a.check))
    if a.profile:
        import cProfile, pstats
        prof = cProfile.Profile(); prof.enable()
    if a.tbr:
        brk = tree.edges()
        if a.prunings and a.prunings < len(brk):
            brk = [brk[i] for i in np.linspace(0, len(brk) - 1, a.prunings).astype(int)]
        pr = brk
        est, move, ncand, naln = treesearch.tbr_round_multi(tree, loci, rec, dms=dms, breaks=brk, chunk=a.chunk)
        move_out = [list(move[0]), str(move[1]), str(move[2])] if move else None
    else:
        est, move, ncand, naln = treesearch.spr_round(tree, loci, rec, dms=dms, prunings=pr, chunk=a.chunk)
        move_out = [list(move[0]), list(move[1])] if move else None
    if a.profile:
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
    sync(); t3 = time.perf_counter()
    if world > 1:
        tot = torch.tensor([gb.n_median - m0, gb.n_distance - d0, gb.cells_distance - c0], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        nm, nd, cells = [int(x) for x in tot.tolist()]
    else:
        nm, nd, cells = gb.n_median - m0, gb.n_distance - d0, gb.cells_distance - c0
    out = dict(workload="%d taxa x %d loci (%s bp), random tree, downpass + %s round" % (a.taxa, a.loci, lens, "TBR" if a.tbr else "SPR"),
               n_gpus=world, tree_cost=cost, downpass_s=t1 - t0, downpass_medians=(a.taxa - 1) * a.loci,
               all_directions_s=t2 - t1, spr_prunings=len(pr), spr_candidates=ncand, spr_alignments=naln,
               spr_s=t3 - t2, spr_candidates_per_s=ncand / (t3 - t2), spr_medians=nm, spr_distances=nd,
               spr_distance_gcups=cells / (t3 - t2) / 1e9, best_estimate=est, move=move_out,
               data="synthetic", scaling="strong")
    if a.check and rank == 0:
        from oracle import cost_matrix_oracle as cmo
        from oracle.port import Port
        from tests.oracle_backend import OracleBackend
        full, orig = cmo.dna_matrices(1, 1, 3)
        ob = OracleBackend(Port(), full, orig)
        med = rec.med[:: max(1, len(rec.med) // a.check)][:a.check]
        dis = rec.dis[:: max(1, len(rec.dis) // a.check)][:a.check]
        bad = 0
        for (p, o), r in zip(med, ob.median([p for p, _ in med])):
            bad += not (np.array_equal(o[0], r[0]) and o[1] == r[1])
        for (p, o), r in zip(dis, ob.distance([p for p, _ in dis])):
            bad += int(o != r)
        out.update(checked_medians=len(med), checked_distances=len(dis), mismatches=int(bad))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
