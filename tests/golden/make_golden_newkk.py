"""Generates tests/golden/newkk_golden.npz from the UNMODIFIED reference src/newkkonen.c compiled into
oracle/_ref/libpoyref_long.so (oracle/ref_newkk_driver.c): Sequence.NewkkAlign, affine entry point
(newkkonen_CAML_algn_affine + newkkonen_CAML_backtrace_affine), for four cost regimes -- similar, decorated
(ambiguity / gap-bit symbols), unrelated, ragged / tiny pairs and the trivial path (len1 * 100 < len2).
Inputs are stored; outputs as cost + final k + row lengths + SHA-1 of the two aligned rows.
Run in the build container:  python tests/golden/make_golden_newkk.py"""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cost_matrix_oracle as cmo
from oracle.refbind import RefLib
from poy5_b200 import synth

REGIMES = dict(R1=(1, 1, 3), R2=(2, 1, 5), R3=(1, 2, 0), R5=(3, 2, 4))


def digest(arrs):
    h = hashlib.sha1()
    for x in arrs:
        h.update(np.ascontiguousarray(x, np.uint8).tobytes()); h.update(b"|")
    return np.frombuffer(h.digest(), np.uint8)


def pairs_for(seed):
    rng = np.random.default_rng(seed)
    out = []
    for L in (30, 150, 600, 1500):
        for _ in range(4 if L <= 600 else 2):
            anc = synth.random_seq(rng, L)
            out.append((synth.evolve(rng, anc, 0.10, 0.01), synth.evolve(rng, anc, 0.10, 0.02)))
            a = synth.decorate(rng, synth.evolve(rng, anc, 0.10, 0.01), 0.05, 0.05)
            b = synth.decorate(rng, synth.evolve(rng, anc, 0.25, 0.03), 0.05, 0.05)
            out.append((a, b))
        out.append((synth.random_seq(rng, L), synth.decorate(rng, synth.random_seq(rng, int(L * 0.7)), 0.05, 0.05)))
    for _ in range(12):      # ragged / tiny / empty
        out.append((synth.random_seq(rng, int(rng.integers(0, 6))), synth.decorate(rng, synth.random_seq(rng, int(rng.integers(0, 40))), 0.2, 0.2)))
    out.append((synth.random_seq(rng, 1), synth.random_seq(rng, 400)))        # trivial path
    out.append((np.zeros(0, np.uint8), synth.random_seq(rng, 150)))           # trivial path, empty shorter sequence
    return [(synth.with_gap(a), synth.with_gap(b)) for a, b in out]


def main():
    R = RefLib(True)
    out = {}
    for rname, (s_, g_, go) in REGIMES.items():
        full, _ = cmo.dna_matrices(s_, g_, go)
        rc = R.cm(full)
        seqs, cost, kk, lens2, sha = [], [], [], [], []
        for a, b in pairs_for(4100 + go):
            seqs += [a, b]
            sw = int(len(a) > len(b))
            s1, s2 = (b, a) if sw else (a, b)
            c, r1, r2, k = R.newkk_align(rc, s1, s2, 1, sw)
            triv = len(s1) * 100 < len(s2)
            cost.append(c); kk.append(-1 if triv else k); lens2.append([len(r1), len(r2)]); sha.append(digest([r1, r2]))
        lens = np.array([len(s) for s in seqs], np.int64)
        off = np.zeros(len(seqs) + 1, np.int64); np.cumsum(lens, out=off[1:])
        out[rname + "_data"] = np.concatenate(seqs).astype(np.uint8); out[rname + "_off"] = off
        out[rname + "_cost"] = np.array(cost, np.int32); out[rname + "_k"] = np.array(kk, np.int32)
        out[rname + "_lens2"] = np.array(lens2, np.int32); out[rname + "_sha"] = np.stack(sha)
        out[rname + "_regime"] = np.array([s_, g_, go], np.int32)
        print(rname, len(cost), "pairs", flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "newkk_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
