"""BASELINE configs[4] downpass alone (1000 taxa x 10 kb, random tree: 999 medians, most levels one to four pairs), twice;
POY_TRACE=1 prints one line per level batch.  Run on a GPU box from the repo root: python scripts/probe_cfg5_downpass.py"""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poy5_b200 as pb
from poy5_b200 import workloads, treesearch
from poy5_b200.workloads import _heuristic, make_taxa, random_tree
ctx = pb.Context(0)
h = _heuristic(ctx, (1, 1, 3))
taxa, L = 1000, 10000
host = make_taxa(5, taxa, L)
tree = random_tree(5, taxa)
for rep in range(2):
    sb = treesearch.StoreBackend(ctx, h, cap_bytes=max(1 << 26, 8 * taxa * (L + 64)), cap_seqs=1 << 16)
    loci = [sb.put(host)]
    ctx.synchronize(); t0 = time.perf_counter()
    cost, _ = treesearch.downpass(tree, loci, sb)
    ctx.synchronize(); t1 = time.perf_counter()
    print("rep", rep, "downpass", t1 - t0, "cost", cost, file=sys.stderr)
    sb.close()
print(ctx.stats(), file=sys.stderr)
