"""Generates tests/golden/affine_golden.npz from the UNMODIFIED reference C (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixture then travels with the repo; nothing at test time reads /root/reference."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cost_matrix_oracle as cmo
from oracle.refbind import RefLib
from tests.helpers import REGIMES, edge_pairs, witness_pairs
from poy5_b200 import synth


def main():
    R = RefLib(False)
    out = {}
    for rname, (s_, g_, go) in REGIMES.items():
        full, _ = cmo.dna_matrices(s_, g_, go)
        rc = R.cm(full)
        seqs, ia, ib = edge_pairs(4242 + go, n=120, maxlen=48)
        more, ja, jb = synth.pair_batch(99 + go, 10, 260, frac_decorated=0.5, jitter=0.3)
        base = len(seqs)
        seqs = seqs + more
        ia = np.concatenate([ia, ja + base]); ib = np.concatenate([ib, jb + base])
        if rname in witness_pairs():
            a, b = witness_pairs()[rname]
            seqs += [np.array(a, np.uint8), np.array(b, np.uint8)]
            ia = np.append(ia, len(seqs) - 2); ib = np.append(ib, len(seqs) - 1)
        lens = np.array([len(s) for s in seqs], np.int64)
        off = np.zeros(len(seqs) + 1, np.int64); np.cumsum(lens, out=off[1:])
        data = np.concatenate(seqs).astype(np.uint8)
        cost, acost, blobs, blens = [], [], [], []
        for p in range(len(ia)):
            a, b = seqs[ia[p]], seqs[ib[p]]
            cost.append(R.cost_affine(rc, a, b))
            sw = int(len(a) > len(b))
            si, sj = (b, a) if sw else (a, b)
            r = R.align_affine(rc, si, sj, sw)
            acost.append(r[0])
            for x in r[1:]:
                blobs.append(x); blens.append(len(x))
        out[rname + "_data"] = data; out[rname + "_off"] = off
        out[rname + "_ia"] = ia.astype(np.int32); out[rname + "_ib"] = ib.astype(np.int32)
        out[rname + "_cost"] = np.array(cost, np.int32); out[rname + "_acost"] = np.array(acost, np.int32)
        out[rname + "_blob"] = np.concatenate(blobs).astype(np.uint8); out[rname + "_blens"] = np.array(blens, np.int32)
        out[rname + "_regime"] = np.array([s_, g_, go], np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "affine_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
