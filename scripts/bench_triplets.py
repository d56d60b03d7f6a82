"""BASELINE config #3: three-sequence medians for final-state assignment, through the reference's LIVE path
(SeqCS.DOS.median_3_union, src/seqCS.ml:1151-1178: union of the two aligned children, ONE pairwise alignment
parent x union, median_2).  The 3-D cube of the north star is dead and wrong in the reference (DESIGN.md section
7), so there is no reference number for it; the CPU figure printed with --cpu is the oracle replay of the same
composition on one core over a small sample."""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poy5_b200 as pb
from poy5_b200 import synth
from poy5_b200.cost_matrix import Two_D
from poy5_b200.seqcs import DOS, Heuristic, median_3_union

ap = argparse.ArgumentParser()
ap.add_argument("--triplets", type=int, default=50000)
ap.add_argument("--L", type=int, default=1500)
ap.add_argument("--chunk", type=int, default=10000)
ap.add_argument("--cpu", type=int, default=0, help="oracle replay of this many triplets (single core)")
a = ap.parse_args()
ctx = pb.Context(0)
t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))


def make(seed, n, L):
    rng = np.random.default_rng(seed)
    seqs = []
    for t in range(n):
        anc = synth.random_seq(rng, L)
        seqs += [synth.with_gap(synth.evolve(rng, anc, 0.05, 0.005)) for _ in range(3)]   # parent, child 1, child 2
    return seqs


def gpu(seqs):
    n = len(seqs) // 3
    pool = pb.Pool(ctx, seqs)
    ip = np.arange(0, 3 * n, 3, dtype=np.int32)
    node = DOS.median(ctx, h, pool, ip + 1, ip + 2)
    got = median_3_union(ctx, h.c2_full, pool, ip, node["aligned_a"], node["aligned_b"])
    pool.close()
    return node, got


gpu(make(1, 64, 200))                                   # warm-up
chunks = [make(3000 + c, min(a.chunk, a.triplets - c), a.L) for c in range(0, a.triplets, a.chunk)]
t0 = time.perf_counter()
tot = 0
for seqs in chunks:
    node, got = gpu(seqs)
    tot += int(got["cost"].sum())
tg = time.perf_counter() - t0
out = dict(triplets=a.triplets, L=a.L, gpu_seconds=tg, triplets_per_s=a.triplets / tg, sum_cost=tot)
if a.cpu:
    from oracle import cost_matrix_oracle as cmo
    from oracle.port import Port
    from tests.helpers import oracle_align
    P = Port(); pf = P.cm(cmo.dna_matrices(1, 1, 3)[0])
    seqs = chunks[0][:3 * a.cpu]
    node, got = gpu(seqs)
    t0 = time.perf_counter()
    same = True
    for t in range(a.cpu):
        p, c1, c2 = seqs[3 * t], seqs[3 * t + 1], seqs[3 * t + 2]
        _, _, _, ra, rb = oracle_align(P, pf, c1, c2)
        u = P.union(ra, rb)
        oc, _, _, xa, xb = oracle_align(P, pf, p, u)
        med = P.median_2(pf, xa, xb, False)
        same &= (oc == got["cost"][t]) and np.array_equal(med, got["sequence"][t])
    tc = time.perf_counter() - t0
    out.update(cpu_sample=a.cpu, cpu_triplets_per_s=a.cpu / tc, identical=bool(same))
print(json.dumps(out))
