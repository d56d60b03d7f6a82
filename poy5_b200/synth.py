"""Seeded synthetic DNA workloads of the shapes BASELINE.json names.

Used by bench.py and the tests only; nothing here is on the alignment path."""
import numpy as np

BASES = np.array([1, 2, 4, 8], np.uint8)
GAP = 16
REGIMES = {"R1": (1, 1, 3), "R2": (2, 1, 5), "R3": (1, 2, 0)}


def random_seq(rng, length):
    return BASES[rng.integers(0, 4, size=length)]


def evolve(rng, anc, subst=0.10, indel=0.01, mean_indel=3.0):
    """child = ancestor with `subst` substitutions per site and indel events at rate `indel` per
    site, geometric length with mean `mean_indel` (vectorised)."""
    n = len(anc)
    child = anc.copy()
    m = rng.random(n) < subst
    child[m] = BASES[rng.integers(0, 4, size=int(m.sum()))]
    ev = rng.random(n) < indel
    if not ev.any():
        return child
    pos = np.flatnonzero(ev)
    lens = rng.geometric(1.0 / mean_indel, size=len(pos))
    is_del = rng.random(len(pos)) < 0.5
    keep = np.ones(n, bool)
    pieces, last = [], 0
    for p, l, d in zip(pos, lens, is_del):
        if p < last:
            continue
        pieces.append(child[last:p])
        if d:
            last = min(n, p + l)
        else:
            pieces.append(random_seq(rng, l))
            last = p
    pieces.append(child[last:])
    return np.concatenate(pieces).astype(np.uint8)


def decorate(rng, s, p_amb=0.05, p_gap=0.03):
    """internal-node-like symbols: ambiguity codes (OR of two bases) and gap-bit codes (x|16)."""
    s = s.copy()
    m = rng.random(len(s)) < p_amb
    s[m] |= BASES[rng.integers(0, 4, size=int(m.sum()))]
    m = rng.random(len(s)) < p_gap
    s[m] |= GAP
    return s


def with_gap(s):
    return np.concatenate([np.array([GAP], np.uint8), np.asarray(s, np.uint8)])


def pair_batch(seed, n, length, subst=0.10, indel=0.01, frac_decorated=0.0, jitter=0.0):
    """n (ancestor-derived) pairs of length ~`length`.  Returns (list of sequences, idx_a, idx_b):
    pair p is sequences 2p and 2p+1."""
    rng = np.random.default_rng(seed)
    seqs = []
    for p in range(n):
        L = length if jitter == 0 else max(1, int(length * (1 + jitter * (rng.random() * 2 - 1))))
        anc = random_seq(rng, L)
        a, b = evolve(rng, anc, subst, indel), evolve(rng, anc, subst, indel)
        if rng.random() < frac_decorated:
            a, b = decorate(rng, a), decorate(rng, b)
        seqs.append(with_gap(a)); seqs.append(with_gap(b))
    idx = np.arange(n, dtype=np.int32)
    return seqs, 2 * idx, 2 * idx + 1


def cells(pool_lens, a, b):
    """full-matrix cells (len_i-1)*(len_j-1) per pair."""
    return (pool_lens[a] - 1).astype(np.int64) * (pool_lens[b] - 1).astype(np.int64)


# ---- fast C generator (poy5_b200/csrc/synth.c -> libpoysynth.so) -------------------------------
_SYNTH = None


def _synth_lib():
    global _SYNTH
    if _SYNTH is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpoysynth.so")
        if not os.path.exists(path):
            raise ImportError("libpoysynth.so is not built; run __graft_entry__.build()")
        L = C.CDLL(path)
        L.synth_pairs_capacity.restype = C.c_int64
        L.synth_pairs_capacity.argtypes = [C.c_int64, C.c_int, C.c_double]
        L.synth_pairs.restype = C.c_int64
        L.synth_pairs.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        _SYNTH = L
    return _SYNTH


def pair_pool(seed, first_pair, n, length, jitter=0.0, subst=0.10, indel=0.01, decorated=0.10, out=None, nthreads=8):
    """n pairs (pair p = sequences 2p, 2p+1) as (data uint8[total], offsets int64[2n+1]).
    `out` may be a preallocated uint8 array of at least pair_pool_capacity() bytes (e.g. pinned)."""
    L = _synth_lib()
    cap = L.synth_pairs_capacity(n, length, jitter)
    data = out if out is not None else np.empty(cap, np.uint8)
    assert data.nbytes >= cap
    off = np.empty(2 * n + 1, np.int64)
    tot = L.synth_pairs(seed, first_pair, n, length, jitter, subst, indel, decorated, data.ctypes.data, off.ctypes.data,
                        nthreads)
    return data[:tot], off


def pair_pool_capacity(n, length, jitter=0.0):
    return _synth_lib().synth_pairs_capacity(n, length, jitter)
