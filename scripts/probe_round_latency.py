"""Latency of ONE small batch of interior-node-like pairs through the banded align entry point (what a tree pass
issues per level).  Run under POY_TRACE=2 for the per-round split, or under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/probe_round_latency.py
for the per-kernel durations (serialised)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=40)
    ap.add_argument("--len", type=int, default=2500)
    ap.add_argument("--subst", type=float, default=0.2)
    ap.add_argument("--indel", type=float, default=0.02)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import poy5_b200 as pb
    from poy5_b200 import synth
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    ctx = pb.Context(0)
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, 3).full)
    data, off = synth.pair_pool(99, 0, a.pairs, a.len, subst=a.subst, indel=a.indel, decorated=1.0)
    seqs = [data[off[s]:off[s + 1]] for s in range(2 * a.pairs)]
    pool = pb.Pool(ctx, seqs)
    ia = np.arange(0, 2 * a.pairs, 2, dtype=np.int32); ib = ia + 1
    for r in range(a.reps):
        ctx.synchronize(); t = time.perf_counter()
        res = Align.align_affine_3(ctx, cm, pool, ia, ib)
        ctx.synchronize()
        print("rep %d: %d pairs of %d bp: %.2f ms, cost sum %d" % (r, a.pairs, a.len, 1e3 * (time.perf_counter() - t), int(np.sum(res["cost"]))), file=sys.stderr)
    print(ctx.stats(), file=sys.stderr)


if __name__ == "__main__":
    main()
