"""GPU bring-up check: CUDA path vs the CPU oracle port on random pairs (run under gpurun)."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poy5_b200 as pb
from poy5_b200 import synth
from poy5_b200.cost_matrix import Two_D
from poy5_b200.sequence import Align
from oracle.port import Port
from oracle import cost_matrix_oracle as cmo

P = Port()
ctx = pb.Context(0)
out = {}
for kind, name in enumerate(["iadd", "imin", "viaddmin", "vimin3", "cellmix"]):
    ops, mhz = ctx.microbench(kind)
    out[name] = dict(ops_per_s=ops, clock_mhz=mhz)
    print("microbench %-9s %.3e thread-ops/s  (%.0f MHz)" % (name, ops, mhz), flush=True)

def run(regime, seqs, ia, ib, tag, check_align=True):
    s_, g_, go = regime
    t2d = Two_D.of_transformations_and_gaps(s_, g_, go)
    f, o = cmo.dna_matrices(s_, g_, go)
    pc = P.cm(f)
    cm = pb.CostModel(ctx, t2d.full)
    pool = pb.Pool(ctx, seqs)
    t0 = time.time()
    cost = Align.cost_2(ctx, cm, pool, ia, ib)
    t1 = time.time()
    bad = 0
    for p in range(len(ia)):
        c = P.cost_affine(pc, seqs[ia[p]], seqs[ib[p]])
        if c != cost[p]:
            bad += 1
            if bad <= 3:
                print("  COST MISMATCH", tag, regime, "pair", p, "lens", len(seqs[ia[p]]), len(seqs[ib[p]]), "gpu", cost[p], "oracle", c)
                if len(seqs[ia[p]]) < 40:
                    print("   a=", list(seqs[ia[p]]), "b=", list(seqs[ib[p]]))
    print("%s %s cost_2: %d pairs, %d mismatches, gpu %.3fs" % (tag, regime, len(ia), bad, t1 - t0), flush=True)
    abad = 0
    if check_align:
        t0 = time.time()
        r = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
        t1 = time.time()
        for p in range(len(ia)):
            a, b = seqs[ia[p]], seqs[ib[p]]
            sw = int(len(a) > len(b))
            si, sj = (b, a) if sw else (a, b)
            oc, om, ow, ori, orj, st = P.align_affine(pc, si, sj, sw, with_stats=True)
            ra, rb = (orj, ori) if sw else (ori, orj)
            ok = (oc == r["cost"][p] and np.array_equal(om, r["median"][p]) and np.array_equal(ow, r["medianwg"][p])
                  and np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]))
            if not ok:
                abad += 1
                if abad <= 3:
                    print("  ALIGN MISMATCH", tag, regime, "pair", p, "lens", len(a), len(b), "sw", sw, "gpu cost", r["cost"][p], "oracle", oc,
                          "iters gpu/oracle", r["stats"][p][0], st.iterations, "k", r["stats"][p][2], st.final_k)
                    if len(a) < 40:
                        print("   a=", list(a), "b=", list(b))
                        print("   gpu med", list(r["median"][p]), "oracle", list(om))
                        print("   gpu ra", list(r["res_a"][p]), "oracle", list(ra))
                        print("   gpu rb", list(r["res_b"][p]), "oracle", list(rb))
        print("%s %s align_affine_3: %d pairs, %d mismatches, gpu %.3fs" % (tag, regime, len(ia), abad, t1 - t0), flush=True)
    cm.close(); pool.close()
    return bad, abad

tot = [0, 0]
rng = np.random.default_rng(7)
for rname, regime in synth.REGIMES.items():
    # tiny / ragged / empty
    seqs = []
    for p in range(600):
        L = int(rng.integers(0, 40))
        anc = synth.random_seq(rng, L)
        a = synth.evolve(rng, anc, 0.15, 0.06); b = synth.evolve(rng, anc, 0.15, 0.06)
        if p % 3 == 0:
            a, b = synth.decorate(rng, a, 0.1, 0.1), synth.decorate(rng, b, 0.1, 0.1)
        if p % 7 == 0:
            b = synth.random_seq(rng, int(rng.integers(0, 50)))
        seqs += [synth.with_gap(a), synth.with_gap(b)]
    idx = np.arange(600, dtype=np.int32)
    b_, a_ = run(regime, seqs, 2 * idx, 2 * idx + 1, "tiny")
    tot[0] += b_; tot[1] += a_
    for L, n, dec in ((150, 200, 0.3), (600, 100, 0.3), (1300, 40, 0.5), (2500, 12, 0.0)):
        seqs, ia, ib = synth.pair_batch(1000 + L, n, L, frac_decorated=dec, jitter=0.2)
        b_, a_ = run(regime, seqs, ia, ib, "L%d" % L)
        tot[0] += b_; tot[1] += a_
# unrelated sequences (wide bands -> generic kernel)
seqs = []
for p in range(6):
    seqs += [synth.with_gap(synth.random_seq(rng, 300 + 40 * p)), synth.with_gap(synth.random_seq(rng, 700))]
idx = np.arange(6, dtype=np.int32)
b_, a_ = run(synth.REGIMES["R1"], seqs, 2 * idx, 2 * idx + 1, "unrelated")
tot[0] += b_; tot[1] += a_
print("TOTAL cost mismatches %d, align mismatches %d, launches %d" % (tot[0], tot[1], ctx.launches))
out["mismatch"] = tot
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
