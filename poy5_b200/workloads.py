"""BASELINE.json configs #1, #3 and #5 as workloads behind the batch API (config #2 is bench.py's headline sweep,
config #4 the `swap_eval` sub-record): what POY does around the alignment stubs in each of them, restated as a
workload generator, so that the driver's bench line carries a measured, parity-sampled record
of every configuration.  Each function returns (record, sample); the sample -- recorded (inputs, outputs) of some of
the medians / distances that went through the GPU, as host arrays -- is replayed on the CPU checker by the CALLER
(bench.py / tests own the checker; nothing in this package touches it)."""
import time

import numpy as np

from . import synth, treesearch
from .swap_eval import SampleRecorder, make_taxa, random_tree


def _heuristic(ctx, regime):
    import poy5_b200 as pb
    from .cost_matrix import Two_D
    from .seqcs import Heuristic
    t2d = Two_D.of_transformations_and_gaps(*regime)
    return Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))


def config1(ctx, taxa=20, L=1800, regime=(1, 1, 3), check=12):
    """configs[0]: 20 synthetic taxa x 1.8 kb, dynamic homology, affine gaps, Wagner build + one TBR round, report cost
    (the POY script `build(); swap(tbr)` reduced to its alignment workload: src/ptree.ml:1144-1270, 1356-1453)."""
    h = _heuristic(ctx, regime)
    leaves_h = make_taxa(1, taxa, L)
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=1 << 24, cap_seqs=1 << 14)
    warm = sb.put(make_taxa(2, 6, 200))
    treesearch.tbr_round(treesearch.wagner_build(warm, sb), warm, sb)
    leaves = sb.put(leaves_h)
    rec = SampleRecorder(sb, 331, 307, cap=check) if check else sb
    m0, d0 = sb.n_median, sb.n_distance
    ctx.synchronize(); t0 = time.perf_counter()
    tree = treesearch.wagner_build(leaves, rec)
    c0 = treesearch.tree_cost(tree, leaves, rec)
    est, move, ncand = treesearch.tbr_round(tree, leaves, rec)
    c1 = treesearch.tree_cost(treesearch.apply_tbr(tree, move), leaves, rec) if move is not None else c0
    ctx.synchronize(); secs = time.perf_counter() - t0
    out = dict(workload="configs[0]: %d taxa x %d bp, Wagner build + one TBR round" % (taxa, L), seconds=secs, build_cost=int(c0),
               tbr_estimate=None if est is None else int(est), tbr_candidates=int(ncand), cost_after_tbr=int(c1),
               medians=sb.n_median - m0, distances=sb.n_distance - d0)
    sample = (rec.med, rec.dis) if check else None
    sb.close()
    return out, sample


def config3(ctx, triplets=50000, L=1500, chunk=10000, regime=(1, 1, 3), check=6):
    """configs[2]: three-sequence medians for final-state assignment through the reference's LIVE path
    (SeqCS.DOS.median_3_union, src/seqCS.ml:1151-1178: union of the two aligned children, one pairwise alignment
    parent x union, median_2); the 3-D cube of the config's wording is dead and wrong in the reference."""
    import poy5_b200 as pb
    from .seqcs import DOS, median_3_union
    h = _heuristic(ctx, regime)

    def make(seed, n, length):
        rng = np.random.default_rng(seed)
        seqs = []
        for _ in range(n):
            anc = synth.random_seq(rng, length)
            seqs += [synth.with_gap(synth.evolve(rng, anc, 0.05, 0.005)) for _ in range(3)]      # parent, child 1, child 2
        return seqs

    def run(seqs):
        n = len(seqs) // 3
        pool = pb.Pool(ctx, seqs)
        ip = np.arange(0, 3 * n, 3, dtype=np.int32)
        node = DOS.median(ctx, h, pool, ip + 1, ip + 2)
        got = median_3_union(ctx, h.c2_full, pool, ip, node["aligned_a"], node["aligned_b"])
        pool.close()
        return got
    run(make(1, 64, 200))
    chunks = [make(3000 + c, min(chunk, triplets - c), L) for c in range(0, triplets, chunk)]
    ctx.synchronize(); t0 = time.perf_counter()
    tot, first = 0, None
    for seqs in chunks:
        got = run(seqs)
        tot += int(got["cost"].sum())
        first = got if first is None else first
    ctx.synchronize(); secs = time.perf_counter() - t0
    out = dict(workload="configs[2]: %d three-sequence medians at %d bp (median_3_union)" % (triplets, L), seconds=secs,
               triplets_per_s=triplets / secs, sum_cost=tot)
    sample = [(chunks[0][3 * t], chunks[0][3 * t + 1], chunks[0][3 * t + 2], int(first["cost"][t]), np.array(first["sequence"][t], np.uint8))
              for t in range(min(check, len(chunks[0]) // 3))]
    return out, sample


def config5(ctx, taxa=1000, L=10000, prunings=6, regime=(1, 1, 3), check=8, seed=5):
    """configs[4]: 1000 taxa x 10 kb, one downpass (999 medians, level-synchronous) + an SPR sample (incremental medians,
    edge medians, one cost-only distance per candidate) on the device-resident node store."""
    h = _heuristic(ctx, regime)
    host = make_taxa(seed, taxa, L)
    tree = random_tree(seed, taxa)
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=max(1 << 26, 8 * taxa * (L + 64)), cap_seqs=1 << 16)
    loci = [sb.put(host)]
    ctx.synchronize(); t0 = time.perf_counter()
    cost, _ = treesearch.downpass(tree, loci, sb)
    ctx.synchronize(); t1 = time.perf_counter()
    dms = treesearch.all_directions(tree, loci, sb)
    ctx.synchronize(); t2 = time.perf_counter()
    pr = treesearch.spr_prunings(tree, taxa)
    pr = [pr[i] for i in np.linspace(0, len(pr) - 1, prunings).astype(int)]
    rec = SampleRecorder(sb, max(1, prunings * taxa // max(1, check)), max(1, prunings * taxa // (2 * max(1, check))), cap=check) if check else sb
    c0 = sb.cells_distance
    ctx.synchronize(); t3 = time.perf_counter()
    est, move, ncand, naln = treesearch.spr_round(tree, loci, rec, dms=dms, prunings=pr, chunk=prunings)
    ctx.synchronize(); t4 = time.perf_counter()
    out = dict(workload="configs[4]: %d taxa x %d bp, downpass + SPR sample of %d prunings" % (taxa, L, len(pr)),
               downpass_s=t1 - t0, downpass_medians=taxa - 1, tree_cost=int(cost), all_directions_s=t2 - t1,
               spr_seconds=t4 - t3, spr_candidates=int(ncand), spr_alignments=int(naln), spr_candidates_per_s=ncand / (t4 - t3),
               spr_distance_gcups=(sb.cells_distance - c0) / (t4 - t3) / 1e9, best_estimate=None if est is None else int(est))
    sample = (rec.med, rec.dis) if check else None
    sb.close()
    return out, sample


def newkk(ctx, pairs=2000, L=2000, regime=(1, 1, 3), check=8):
    """Sequence.NewkkAlign.align_2 (src/newkkonen.c, affine entry point) on interior-node-like pairs: cost + both aligned
    rows per pair, threshold doubling inside the kernel."""
    import poy5_b200 as pb
    from .cost_matrix import Two_D
    from .sequence import NewkkAlign
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(*regime).full)
    data, off = synth.pair_pool(4242, 0, pairs, L, subst=0.10, indel=0.01, decorated=0.5)
    seqs = [data[off[s]:off[s + 1]] for s in range(2 * pairs)]
    pool = pb.Pool(ctx, seqs)
    ia = np.arange(0, 2 * pairs, 2, dtype=np.int32); ib = ia + 1
    NewkkAlign.cost_2(ctx, cm, pool, ia[:64], ib[:64])          # warm-up
    ctx.synchronize(); t0 = time.perf_counter()
    r = NewkkAlign.align_2(ctx, cm, pool, ia, ib, stats=True)
    ctx.synchronize(); secs = time.perf_counter() - t0
    lens = np.diff(off)
    cells = int(((lens[ia] - 1) * (lens[ib] - 1)).sum())
    out = dict(workload="Sequence.NewkkAlign.align_2: %d pairs of %d bp" % (pairs, L), seconds=secs, alignments_per_s=pairs / secs,
               gcups_full_matrix_equivalent=cells / secs / 1e9, mean_doublings=float(r["stats"][:, 0].mean()),
               median_final_k=int(np.median(r["stats"][:, 2])))
    sample = [(seqs[2 * p], seqs[2 * p + 1], int(r["cost"][p]), np.array(r["res_a"][p], np.uint8), np.array(r["res_b"][p], np.uint8))
              for p in np.linspace(0, pairs - 1, check).astype(int)]
    pool.close(); cm.close()
    return out, sample


def main():
    """python -m poy5_b200.workloads --out FILE [--device D]: runs the three workloads in THIS process and pickles
    {name: (record, sample) | {"error": ...}} -- bench.py runs it as a child process so that a failure or a time-out of
    a side record can never take the headline JSON line with it."""
    import argparse
    import pickle
    import poy5_b200 as pb
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--regime", default="1,1,3")
    ap.add_argument("--small", action="store_true", help="toy sizes (tests)")
    a = ap.parse_args()
    regime = tuple(int(x) for x in a.regime.split(","))
    ctx = pb.Context(a.device)
    out = {}
    jobs = (("configs[0]", config1, dict(taxa=7, L=120) if a.small else {}),
            ("configs[2]", config3, dict(triplets=40, L=150, chunk=20) if a.small else {}),
            ("configs[4]", config5, dict(taxa=12, L=300, prunings=3) if a.small else {}),
            ("newkkonen", newkk, dict(pairs=40, L=200) if a.small else {}))
    for name, fn, kw in jobs:
        try:
            out[name] = fn(ctx, regime=regime, **kw)
        except Exception as e:
            out[name] = dict(error="%s: %s" % (type(e).__name__, e))
        with open(a.out, "wb") as f:          # rewritten after every workload: a later time-out keeps the earlier records
            pickle.dump(out, f)
    ctx.close()


if __name__ == "__main__":
    main()
