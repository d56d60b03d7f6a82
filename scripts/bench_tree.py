"""BASELINE config #1: 20 synthetic taxa x 1.8 kb DNA, dynamic homology, affine gaps, Wagner build + one TBR
round, report cost.  GPU driver (poy5_b200.treesearch) vs the CPU replay of the identical call sequence
through the oracle (reference C via oracle/_ref when present, else the port)."""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poy5_b200 as pb
from poy5_b200 import synth, treesearch
from poy5_b200.cost_matrix import Two_D
from poy5_b200.seqcs import Heuristic
from oracle import cost_matrix_oracle as cmo
from oracle.port import Port
from tests.oracle_backend import OracleBackend
from tests.test_treesearch import taxa, run

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=20)
ap.add_argument("--L", type=int, default=1800)
ap.add_argument("--cpu", action="store_true", help="also run the CPU replay (slow: single core)")
a = ap.parse_args()
leaves = taxa(1, a.taxa, a.L)
ctx = pb.Context(0)
t2d = Two_D.of_transformations_and_gaps(1, 1, 3)
h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
gb = treesearch.GpuBackend(ctx, h)
run(gb, taxa(2, 6, 200))   # warm-up
gb = treesearch.GpuBackend(ctx, h)
t0 = time.perf_counter(); got = run(gb, leaves); tg = time.perf_counter() - t0
out = dict(taxa=a.taxa, L=a.L, build_cost=got[1], tbr_estimate=got[2], tbr_candidates=got[4], cost_after_tbr=got[6],
           gpu_seconds=tg, medians=gb.n_median, distances=gb.n_distance, distance_gcups=gb.cells_distance / tg / 1e9)
if a.cpu:
    full, orig = cmo.dna_matrices(1, 1, 3)
    t0 = time.perf_counter(); ref = run(OracleBackend(Port(), full, orig), leaves); tc = time.perf_counter() - t0
    out.update(cpu_seconds=tc, identical=(ref == got))
print(json.dumps(out))
