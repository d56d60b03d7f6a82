"""The swap-evaluation workload (BASELINE configs #4 / #5; north star: ">= 7x from 1 to 8 GPUs on the swap-evaluation
workload"): ONE SPR neighbourhood of a fixed tree over dynamic-homology loci, strong-scaled over the ranks.

Reference seam: `Ptree` evaluates the join candidates of a break through `Parmap.parmap cost_fn`
(src/ptree.ml:1356-1408, 1226-1268), and `cost_fn` is `AllDirChar.cost_fn` -> `Node.distance` of the clade root
against the (lazily recomputed) median of every join edge (src/allDirChar.ml:2132-2177).  Here:

  * setup (untimed, replicated on every rank): synthetic taxa, a starting tree, the all-direction medians of the
    unbroken tree in the device-resident node store (what POY holds on the tree when a swap round starts);
  * timed: for every pruning of the sample, the incremental medians that see the cut, the edge median of every join
    edge, one cost-only distance per (join edge, locus) and the reduction of the best candidate -- prunings dealt to
    the ranks by estimated work, all loci of a candidate on one GPU, two 8-byte MIN all-reduces + one SUM at the end
    (treesearch.spr_round_sharded);
  * untimed: a sample of the medians and distances that went through the GPU is handed back to the caller, which
    replays it on the CPU checker (bench.py, tests) -> parity {checked, mismatches}.
"""
import time

import numpy as np


def make_taxa(seed, n, L):
    """n taxa evolved along a random bifurcating history (5 % substitutions, 0.5 % indels per branch)"""
    from . import synth
    rng = np.random.default_rng(seed)
    pool = [synth.random_seq(rng, L)]
    while len(pool) < n:
        p = pool.pop(int(rng.integers(0, len(pool))))
        pool += [synth.evolve(rng, p, 0.05, 0.005), synth.evolve(rng, p, 0.05, 0.005)]
    return [synth.with_gap(s) for s in pool[:n]]


def random_tree(seed, n):
    from .treesearch import Tree
    rng = np.random.default_rng(seed)
    t = Tree(); t.add_edge(0, 1)
    for leaf in range(2, n):
        edges = t.edges()
        u, v = edges[int(rng.integers(0, len(edges)))]
        w = max(max(t.adj) + 1, n)
        t.remove_edge(u, v); t.add_edge(u, w); t.add_edge(w, v); t.add_edge(w, leaf)
    return t


class SampleRecorder:
    """Wraps a StoreBackend; keeps every `every`-th (inputs, output) of the median / distance batches as HOST byte
    arrays (fetched when recorded: the store's temporaries are released chunk by chunk)."""

    def __init__(self, b, every_median, every_distance, cap=64):
        self.b, self.em, self.ed, self.cap = b, max(1, every_median), max(1, every_distance), cap
        self.med, self.dis, self.km, self.kd = [], [], 0, 0

    def __getattr__(self, name):
        return getattr(self.b, name)

    def median(self, pairs):
        r = self.b.median(pairs)
        if len(self.med) < self.cap:
            first = (-self.km) % self.em
            pick = list(range(first, len(pairs), self.em))[: self.cap - len(self.med)]
            if pick:
                got = self.b.fetch([x for q in pick for x in (pairs[q][0], pairs[q][1], r[q][0])])
                for j, q in enumerate(pick):
                    self.med.append((got[3 * j].copy(), got[3 * j + 1].copy(), got[3 * j + 2].copy(), int(r[q][1])))
        self.km += len(pairs)
        return r

    def median_ids(self, a, b):
        ids, ln, c2 = self.b.median_ids(a, b)
        if len(self.med) < self.cap:
            first = (-self.km) % self.em
            pick = list(range(first, len(a), self.em))[: self.cap - len(self.med)]
            if pick:
                got = self.b.store.read(np.array([x for q in pick for x in (a[q], b[q], ids[q])], np.int32))
                for j, q in enumerate(pick):
                    self.med.append((got[3 * j].copy(), got[3 * j + 1].copy(), got[3 * j + 2].copy(), int(c2[q])))
        self.km += len(a)
        return ids, ln, c2

    def distance_ids(self, a, b, la, lb):
        r = self.b.distance_ids(a, b, la, lb)
        if len(self.dis) < self.cap:
            first = (-self.kd) % self.ed
            pick = list(range(first, len(a), self.ed))[: self.cap - len(self.dis)]
            if pick:
                got = self.b.store.read(np.array([x for q in pick for x in (a[q], b[q])], np.int32))
                for j, q in enumerate(pick):
                    self.dis.append((got[2 * j].copy(), got[2 * j + 1].copy(), int(r[q])))
        self.kd += len(a)
        return r

    def distance(self, pairs):
        r = self.b.distance(pairs)
        if len(self.dis) < self.cap:
            first = (-self.kd) % self.ed
            pick = list(range(first, len(pairs), self.ed))[: self.cap - len(self.dis)]
            if pick:
                got = self.b.fetch([x for q in pick for x in (pairs[q][0], pairs[q][1])])
                for j, q in enumerate(pick):
                    self.dis.append((got[2 * j].copy(), got[2 * j + 1].copy(), int(r[q])))
        self.kd += len(pairs)
        return r


def replicate_lane(ctx_device, h_tables, host_loci, src_backend, dms, arena_bytes=None):
    """A further lane of this rank: own context (stream, scratch arenas), cost models and node store, holding a replica
    of the observed sequences and of the all-direction medians `dms` of `src_backend` (copied through the host once,
    untimed setup).  -> (backend, loci, dms) for treesearch.spr_round_sharded(lanes=...)"""
    import poy5_b200 as pb
    from . import treesearch
    from .seqcs import Heuristic
    c = pb.Context(ctx_device)
    if arena_bytes:
        c.set_arena_limit(int(arena_bytes))
    h = Heuristic(pb.CostModel(c, h_tables.full), pb.CostModel(c, h_tables.original))
    b = treesearch.StoreBackend(c, h, cap_bytes=max(1 << 26, src_backend.store.nbytes * 2), cap_seqs=1 << 17)
    loci = [b.put(ls) for ls in host_loci]
    new_dms = []
    for l, dm in enumerate(dms):
        keys = list(dm.keys())
        seqs = src_backend.fetch([dm[k][0] for k in keys])
        nodes = b.put([np.array(x, np.uint8) for x in seqs])
        new_dms.append({k: (nd, dm[k][1]) for k, nd in zip(keys, nodes)})
    return c, b, loci, new_dms


def run(ctx, rank=0, world=1, device=None, taxa=200, nloci=5, lmin=1000, lmax=3000, seed=4, prunings=128, chunk=32,
        check=24, regime=(1, 1, 3), backend=None, lanes=1, merge_edges=False):
    """One strong-scaled SPR neighbourhood sample.  Returns (record, sample): the `swap_eval` record on rank 0 (None
    elsewhere) and the recorded (medians, distances) sample for the caller's CPU-checker replay (bench.py /
    tests own the checker; nothing in this package touches it)."""
    import poy5_b200 as pb
    from . import treesearch
    from .cost_matrix import Two_D
    from .seqcs import Heuristic
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    t2d = Two_D.of_transformations_and_gaps(*regime)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    rng = np.random.default_rng(seed)
    lens = [int(rng.integers(lmin, lmax + 1)) for _ in range(nloci)]
    host_loci = [make_taxa(seed + 17 * k, taxa, L) for k, L in enumerate(lens)]
    tree = random_tree(seed, taxa)
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=max(1 << 26, 8 * taxa * nloci * (lmax + 64)), cap_seqs=1 << 17) if backend is None else backend
    loci = [sb.put(ls) for ls in host_loci]
    # warm-up: kernels, scratch arenas, pinned staging
    wt = random_tree(1, 8)
    wl = [sb.put(make_taxa(2, 8, 300))]
    treesearch.spr_round(wt, wl, sb, prunings=treesearch.spr_prunings(wt, 8)[:4], chunk=4)
    barrier(); t0 = time.perf_counter()
    cost, _ = treesearch.downpass(tree, loci, sb)
    barrier(); t1 = time.perf_counter()
    dms = treesearch.all_directions(tree, loci, sb)
    barrier(); t2 = time.perf_counter()
    pr = treesearch.spr_prunings(tree, taxa)
    n_all = len(pr)
    if prunings and prunings < len(pr):
        pr = [pr[i] for i in np.linspace(0, len(pr) - 1, prunings).astype(int)]
    # further lanes of this rank (concurrent host threads, each with its own context and node store)
    lane_list, lane_ctx = None, []
    if lanes > 1:
        import torch
        free_b, _ = torch.cuda.mem_get_info(device) if device is not None else (64 << 30, 0)
        arena = int(min(16 << 30, 0.5 * free_b / lanes))
        ctx.set_arena_limit(arena)
        lane_list = [None]
        for q in range(1, lanes):
            c, b, lc, dm = replicate_lane(ctx.device, t2d, host_loci, sb, dms, arena)
            lane_ctx.append((c, b)); lane_list.append((b, lc, dm))
    rec = sb
    if check and rank == 0:
        # expected batch volume on this rank: ~ (4 medians + 1 distance) per (candidate, locus)
        approx = max(1, len(pr) // world) * taxa * nloci
        rec = SampleRecorder(sb, max(1, 4 * approx // max(1, check)), max(1, approx // max(1, check)), cap=check)
    m0, d0, c0 = sb.n_median, sb.n_distance, sb.cells_distance
    if lane_list is not None:
        lane_list[0] = (rec, loci, dms)
    barrier(); t3 = time.perf_counter()
    timing = {}
    est, move, ncand, naln = treesearch.spr_round_sharded(tree, loci, rec, dms, pr, chunk=chunk, rank=rank, world=world, device=device,
                                                          lanes=lane_list, merge_edges=merge_edges, timing=timing)
    for c, _ in lane_ctx:
        c.synchronize()
    barrier(); t4 = time.perf_counter()
    secs = t4 - t3
    nm, nd, cells = sb.n_median - m0, sb.n_distance - d0, sb.cells_distance - c0
    for _, b in lane_ctx:
        nm += b.n_median; nd += b.n_distance; cells += b.cells_distance
    per_rank = None
    if dist is not None and timing:
        import torch
        mine_t = torch.tensor([timing["local_seconds"], float(timing["local_prunings"])], dtype=torch.float64, device=device)
        allt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        per_rank = [dict(seconds=round(float(x[0]), 3), prunings=int(x[1])) for x in allt]
    if dist is not None:
        import torch
        tt = torch.tensor([secs], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs = float(tt.item())
        tot = torch.tensor([nm, nd, cells], dtype=torch.int64, device=device)
        dist.all_reduce(tot)
        nm, nd, cells = [int(x) for x in tot.tolist()]
    # exact rejoin of the winner (the reference's acceptance step after `cc < b_delta`, src/ptree.ml:1410-1453): untimed
    new_cost = None
    if move is not None:
        new_cost, _ = treesearch.downpass(treesearch.apply_spr(tree, move), loci, sb)
    out = None
    if rank == 0:
        out = dict(workload="configs[3]-shaped swap evaluation: %d taxa x %d loci (%s bp), random starting tree, one SPR "
                            "neighbourhood sample of %d of %d prunings, every (pruning, join edge, locus) candidate" % (taxa, nloci, lens, len(pr), n_all),
                   scaling="strong", n_gpus=world, prunings=len(pr), chunk=chunk, lanes_per_gpu=lanes, candidates=int(ncand), alignments=int(naln),
                   medians=nm, distances=nd, seconds=secs, candidates_per_s=ncand / secs, alignments_per_s=naln / secs,
                   distance_gcups=cells / secs / 1e9, best_estimate=None if est is None else int(est),
                   move=None if move is None else [list(move[0]), list(move[1])], tree_cost=int(cost),
                   tree_cost_after_move=None if new_cost is None else int(new_cost),
                   setup=dict(downpass_s=t1 - t0, downpass_medians=(taxa - 1) * nloci, all_directions_s=t2 - t1,
                              note="replicated on every rank, not part of `seconds`"),
                   sharding="prunings dealt to ranks by LPT on rest-tree size; 2 MIN + 1 SUM all-reduce per round",
                   timed="incremental medians + edge medians + cost-only distances + best-candidate reduction, max over ranks (host clock "
                         "between device-synchronised barriers)",
                   node_store_bytes=sb.store.nbytes, per_rank=per_rank, merge_edges=bool(merge_edges))
    sample = (rec.med, rec.dis) if (check and rank == 0) else None
    for c, b in lane_ctx:
        b.close(); c.close()
    if backend is None:
        sb.close()
    return out, sample
