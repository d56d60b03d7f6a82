"""Generates tests/golden/linear_golden.npz from the UNMODIFIED reference C (oracle/_ref): linear-gap
algn_CAML_align_2d (cost + backtrace_2d) and the column-wise helpers (median_2, union, worst_2, verify_2,
ancestor_2) on the aligned rows.  Run in the build container:  python tests/golden/make_golden_linear.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cost_matrix_oracle as cmo
from oracle.refbind import RefLib
from tests.helpers import edge_pairs
from poy5_b200 import synth

CASES = {"L11": (1, 1, None), "L21": (2, 1, None), "L12": (1, 2, None), "A213": (2, 1, 3)}


def main():
    R = RefLib(False)
    out = {}
    rng = np.random.default_rng(99)
    for name, (s_, g_, go) in CASES.items():
        full, _ = cmo.dna_matrices(s_, g_, go)
        rc = R.cm(full)
        seqs, ia, ib = edge_pairs(700 + s_ * 10 + g_, n=90, maxlen=90)
        more, ja, jb = synth.pair_batch(17 + s_, 8, 400, frac_decorated=0.5, jitter=0.3)
        base = len(seqs); seqs = seqs + more
        ia = np.concatenate([ia, ja + base]); ib = np.concatenate([ib, jb + base])
        lens = np.array([len(s) for s in seqs], np.int64)
        off = np.zeros(len(seqs) + 1, np.int64); np.cumsum(lens, out=off[1:])
        cost, dws, blobs, blens, ints = [], [], [], [], []
        for p in range(len(ia)):
            a, b = seqs[ia[p]], seqs[ib[p]]
            sw = int(len(a) > len(b))
            s1, s2 = (b, a) if sw else (a, b)
            if go is None:
                dw = int(rng.integers(0, 25))
                c, r1, r2 = R.align_linear(rc, s1, s2, dw, sw)
            else:
                dw = 0
                c, _, _, r1, r2 = R.align_affine(rc, s1, s2, sw)
            cost.append(c); dws.append(dw)
            parts = [r1, r2, R.median_2(rc, r1, r2, 0), R.median_2(rc, r1, r2, 1), R.union(r1, r2), R.ancestor_2(rc, r1, r2)]
            for x in parts:
                blobs.append(x); blens.append(len(x))
            ints += [R.worst_2(rc, r1, r2), R.verify_2(rc, r1, r2)]
        out[name + "_data"] = np.concatenate(seqs).astype(np.uint8); out[name + "_off"] = off
        out[name + "_ia"] = ia.astype(np.int32); out[name + "_ib"] = ib.astype(np.int32)
        out[name + "_cost"] = np.array(cost, np.int32); out[name + "_dw"] = np.array(dws, np.int32)
        out[name + "_blob"] = np.concatenate(blobs).astype(np.uint8); out[name + "_blens"] = np.array(blens, np.int32)
        out[name + "_ints"] = np.array(ints, np.int32)
        out[name + "_regime"] = np.array([s_, g_, -1 if go is None else go], np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "linear_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
