"""CUDA path vs the CPU checker on sequences over the WHOLE bitset alphabet 1..31 (any ambiguity code, gap bits anywhere,
runs of pure gaps) -- the inputs interior nodes produce after several medians; tests/test_oracle_vs_ref.py pins the
checker itself on the same kind of input against the compiled reference."""
import numpy as np
import pytest
from tests.test_gpu_parity import check_batch
from tests.helpers import REGIMES

pytestmark = pytest.mark.gpu


def _pairs(seed, n, maxlen):
    rng = np.random.default_rng(seed)
    seqs = []
    for p in range(n):
        la, lb = int(rng.integers(0, maxlen)), int(rng.integers(0, maxlen))
        a = rng.integers(1, 32, size=la).astype(np.uint8)
        b = rng.integers(1, 32, size=lb).astype(np.uint8)
        if p % 4 == 0 and la:                 # related pair: b is a noisy copy of a
            b = a.copy()
            m = rng.random(la) < 0.2
            b[m] = rng.integers(1, 32, size=int(m.sum()))
        seqs += [np.concatenate([[16], a]).astype(np.uint8), np.concatenate([[16], b]).astype(np.uint8)]
    idx = np.arange(n, dtype=np.int32)
    return seqs, 2 * idx, 2 * idx + 1


@pytest.mark.parametrize("rname", ["R1", "R2", "R3"])
def test_full_alphabet_small(ctx, port, rname):
    seqs, ia, ib = _pairs(5 + len(rname), 400, 48)
    check_batch(ctx, port, REGIMES[rname], seqs, ia, ib)


def test_full_alphabet_medium(ctx, port):
    seqs, ia, ib = _pairs(99, 40, 420)
    check_batch(ctx, port, REGIMES["R1"], seqs, ia, ib)
