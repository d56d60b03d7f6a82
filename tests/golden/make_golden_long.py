"""Generates tests/golden/affine_long_golden.npz from the UNMODIFIED reference C compiled with
-DUSE_LONG_SEQUENCES (oracle/_ref/libpoyref_long.so): pins at the BASELINE lengths 2 kb and 10 kb for the three
cost regimes -- similar pairs, pairs with ambiguity / gap-bit symbols (interior-node-like), unrelated pairs (wide
bands: classes >= 1280, multi-block cost-only path) and len1 + len2 > 16382.  Inputs are stored (compressed);
outputs are stored as costs + lengths + SHA-1 of the four result sequences to keep the fixture small.
Run in the build container:  python tests/golden/make_golden_long.py"""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cost_matrix_oracle as cmo
from oracle.refbind import RefLib
from poy5_b200 import synth


def digest(arrs):
    h = hashlib.sha1()
    for x in arrs:
        h.update(np.ascontiguousarray(x, np.uint8).tobytes()); h.update(b"|")
    return np.frombuffer(h.digest(), np.uint8)


def pairs_for(seed, L):
    rng = np.random.default_rng(seed)
    out = []
    n_sim, n_dec, n_far = (2, 2, 2) if L <= 2000 else (1, 2, 1)
    for _ in range(n_sim):
        anc = synth.random_seq(rng, L)
        out.append((synth.evolve(rng, anc, 0.10, 0.01), synth.evolve(rng, anc, 0.10, 0.01)))
    for _ in range(n_dec):
        anc = synth.random_seq(rng, L)
        a = synth.decorate(rng, synth.evolve(rng, anc, 0.10, 0.01), 0.05, 0.03)
        b = synth.decorate(rng, synth.evolve(rng, anc, 0.25, 0.02), 0.05, 0.03)
        out.append((a, b))
    for k in range(n_far):
        a = synth.random_seq(rng, L)
        b = synth.random_seq(rng, int(L * (0.8 if k else 1.0)))
        if k:
            a, b = synth.decorate(rng, a, 0.03, 0.03), synth.decorate(rng, b, 0.03, 0.03)
        out.append((a, b))
    return [(synth.with_gap(a), synth.with_gap(b)) for a, b in out]


def main():
    R = RefLib(True)
    out = {}
    for rname, (s_, g_, go) in synth.REGIMES.items():
        full, _ = cmo.dna_matrices(s_, g_, go)
        rc = R.cm(full)
        seqs, cost, acost, lens4, sha = [], [], [], [], []
        for L in (2000, 10000):
            for a, b in pairs_for(7000 + L + go, L):
                seqs += [a, b]
                cost.append(R.cost_affine(rc, a, b))
                sw = int(len(a) > len(b))
                si, sj = (b, a) if sw else (a, b)
                r = R.align_affine(rc, si, sj, sw)
                acost.append(r[0]); lens4.append([len(x) for x in r[1:]]); sha.append(digest(r[1:]))
                print(rname, L, len(a), len(b), cost[-1], acost[-1], flush=True)
        lens = np.array([len(s) for s in seqs], np.int64)
        off = np.zeros(len(seqs) + 1, np.int64); np.cumsum(lens, out=off[1:])
        out[rname + "_data"] = np.concatenate(seqs).astype(np.uint8); out[rname + "_off"] = off
        out[rname + "_cost"] = np.array(cost, np.int32); out[rname + "_acost"] = np.array(acost, np.int32)
        out[rname + "_lens4"] = np.array(lens4, np.int32); out[rname + "_sha"] = np.stack(sha)
        out[rname + "_regime"] = np.array([s_, g_, go], np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "affine_long_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
