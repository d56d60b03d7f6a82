"""Timing probe: cost-only and banded-align batches at one length (run under gpurun / ncu)."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import poy5_b200 as pb
from poy5_b200 import synth, sequence
from poy5_b200.cost_matrix import Two_D

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=2000)
ap.add_argument("--pairs", type=int, default=20000)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mode", default="both")
ap.add_argument("--decorated", type=float, default=0.10)
ap.add_argument("--linear", action="store_true", help="linear-gap cost model (algn_CAML_simple_2 / align_2d) instead of affine")
a = ap.parse_args()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = pb.Context(0, stream.cuda_stream)
cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, None if a.linear else 3).full)
data, off = synth.pair_pool(1234 + a.L, 0, a.pairs, a.L, decorated=a.decorated)
n = a.pairs
lens = np.diff(off)
ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
cells = int(((lens[ia] - 1) * (lens[ib] - 1)).sum())
pool = pb.Pool(ctx, data=data, offsets=off)
def timed(fn):
    best = 1e30
    for _ in range(a.reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, r
if a.mode in ("both", "cost"):
    t, cost = timed(lambda: sequence.Align.cost_2(ctx, cm, pool, ia, ib))
    print("cost-only  L=%d pairs=%d: %.1f ms  %.1f GCUPS  %.0f aln/s" % (a.L, n, t * 1e3, cells / t / 1e9, n / t))
if a.linear:
    t, r = timed(lambda: sequence.Align.align_2(ctx, cm, pool, ia, ib))
    print("linear align_2 L=%d pairs=%d: %.1f ms  %.1f GCUPS-eq  %.0f aln/s" % (a.L, n, t * 1e3, cells / t / 1e9, n / t))
    sys.exit(0)
if a.mode in ("both", "align"):
    t, r = timed(lambda: sequence.Align.align_affine_3(ctx, cm, pool, ia, ib, want=("median",), stats=True))
    st = r["stats"]
    print("align      L=%d pairs=%d: %.1f ms  %.1f GCUPS-eq  %.0f aln/s" % (a.L, n, t * 1e3, cells / t / 1e9, n / t))
    print(" iterations: mean %.2f max %d; final k: median %d p90 %d max %d; band cells/pair mean %.3g (%.1f%% of full)" % (
        st[:, 0].mean(), st[:, 0].max(), np.median(st[:, 2]), np.percentile(st[:, 2], 90), st[:, 2].max(),
        st[:, 3].mean() * 1024, 100 * st[:, 3].sum() * 1024 / cells))
    B = (lens[ib] - lens[ia]).__abs__() + 2 * st[:, 2] + 1
    print(" final band width B: median %d p90 %d max %d; share >512: %.1f%%" % (np.median(B), np.percentile(B, 90), B.max(), 100 * (B > 512).mean()))
    print(" band G cells/s (all iterations): %.1f" % (st[:, 3].sum() * 1024 / t / 1e9))
