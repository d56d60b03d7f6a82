"""Parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs,
plus size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest
from oracle import cost_matrix_oracle as cmo
from tests.helpers import REGIMES, edge_pairs, oracle_align
from poy5_b200 import synth

pytestmark = pytest.mark.gpu


def setup(ctx, regime):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    s_, g_, go = regime
    return pb.CostModel(ctx, Two_D.of_transformations_and_gaps(s_, g_, go).full), cmo.dna_matrices(s_, g_, go)[0]


def check_batch(ctx, port, regime, seqs, ia, ib, align=True):
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, regime)
    pc = port.cm(full)
    pool = pb.Pool(ctx, seqs)
    cost = Align.cost_2(ctx, cm, pool, ia, ib)
    for p in range(len(ia)):
        assert cost[p] == port.cost_affine(pc, seqs[ia[p]], seqs[ib[p]]), ("cost", p)
    if align:
        r = Align.align_affine_3(ctx, cm, pool, ia, ib)
        for p in range(len(ia)):
            oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
            assert oc == r["cost"][p], ("align cost", p)
            assert np.array_equal(om, r["median"][p]) and np.array_equal(ow, r["medianwg"][p]), ("median", p)
            assert np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), ("rows", p)
    cm.close(); pool.close()


@pytest.mark.parametrize("rname", list(REGIMES))
def test_edge_cases(ctx, port, rname):
    """empty, ragged, tiny (exact-emulation path), pure-gap runs, ambiguity and gap-bit codes"""
    seqs, ia, ib = edge_pairs(31 + len(rname), n=500, maxlen=44)
    check_batch(ctx, port, REGIMES[rname], seqs, ia, ib)


@pytest.mark.parametrize("L,n,dec", [(150, 150, 0.3), (600, 60, 0.4), (1300, 24, 0.5), (2500, 8, 0.2)])
@pytest.mark.parametrize("rname", ["R1", "R2", "R3"])
def test_related_pairs(ctx, port, rname, L, n, dec):
    seqs, ia, ib = synth.pair_batch(1000 + L, n, L, frac_decorated=dec, jitter=0.25)
    check_batch(ctx, port, REGIMES[rname], seqs, ia, ib)


def test_unrelated_pairs_wide_bands(ctx, port):
    rng = np.random.default_rng(3)
    seqs = []
    for p in range(8):
        seqs += [synth.with_gap(synth.random_seq(rng, 200 + 60 * p)), synth.with_gap(synth.random_seq(rng, 900))]
    idx = np.arange(8, dtype=np.int32)
    check_batch(ctx, port, REGIMES["R1"], seqs, 2 * idx, 2 * idx + 1)
    check_batch(ctx, port, REGIMES["R2"], seqs, 2 * idx + 1, 2 * idx)   # longer first: swaped path


def test_wrong_order_is_reported(ctx):
    import ctypes
    import poy5_b200 as pb
    from poy5_b200.api import _ptr
    cm, _ = setup(ctx, REGIMES["R1"])
    pool = pb.Pool(ctx, [synth.with_gap([1, 2, 4]), synth.with_gap([1, 2])])
    si = np.array([0], np.int32); sj = np.array([1], np.int32); cost = np.zeros(1, np.int32)
    st = ctx.L.poy_batch_align_affine(ctx.h, cm.h, pool.h, 1, _ptr(si), _ptr(sj), None, None, _ptr(cost),
                                      None, None, None, None, None, None)
    assert st == -4 and b"shorter" in ctx.L.poy_last_error(ctx.h)   # "pass the shorter one as first"
    cm.close(); pool.close()


def test_model_mismatch_is_reported(ctx):
    """the affine entry points refuse a linear cost model and vice versa (POY_ERR_MODEL)"""
    import poy5_b200 as pb
    from poy5_b200.api import _ptr
    from poy5_b200.cost_matrix import Two_D
    lin = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, None).full)
    aff = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, 3).full)
    pool = pb.Pool(ctx, [synth.with_gap([1, 2]), synth.with_gap([1, 2, 4])])
    a = np.array([0], np.int32); b = np.array([1], np.int32); cost = np.zeros(1, np.int32); dw = np.array([3], np.int32)
    assert ctx.L.poy_batch_cost_affine(ctx.h, lin.h, pool.h, 1, _ptr(a), _ptr(b), _ptr(cost)) == -6
    assert ctx.L.poy_batch_cost_linear(ctx.h, aff.h, pool.h, 1, _ptr(a), _ptr(b), _ptr(dw), _ptr(cost)) == -6
    assert ctx.L.poy_batch_cost_linear(ctx.h, lin.h, pool.h, 1, _ptr(b), _ptr(a), _ptr(dw), _ptr(cost)) == -4   # longer first
    lin.close(); aff.close(); pool.close()


# ---- BASELINE sizes: size-independent properties + sampled oracle checks ------------------------------
def _strip(row):
    return row[row != 16]


@pytest.mark.parametrize("L,n,sample", [(500, 20000, 40), (2000, 4000, 12), (10000, 400, 3)])
def test_full_size_properties(ctx, port, L, n, sample):
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES["R1"])
    pc = port.cm(full)
    data, off = synth.pair_pool(0x504F5935 + L, 0, n, L)
    pool = pb.Pool(ctx, data=data, offsets=off)
    ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
    c_ab = Align.cost_2(ctx, cm, pool, ia, ib)
    c_ba = Align.cost_2(ctx, cm, pool, ib, ia)
    # the stub accepts either order and swaps so that the shorter is the row sequence; only for
    # equal lengths is the result order-dependent (rows/columns differ)
    neq = pool.lens[ia] != pool.lens[ib]
    assert np.array_equal(c_ab[neq], c_ba[neq])
    for p in np.flatnonzero(~neq)[:6]:
        assert c_ba[p] == port.cost_affine(pc, pool.seq(ib[p]), pool.seq(ia[p]))
    sub = np.arange(0, n, max(1, n // 2000))
    r = Align.align_affine_3(ctx, cm, pool, ia[sub], ib[sub])
    for q, p in enumerate(sub):
        a, b = pool.seq(ia[p]), pool.seq(ib[p])
        ra, rb = r["res_a"][q], r["res_b"][q]
        assert len(ra) == len(rb) == len(r["medianwg"][q])
        # implied alignment reproduces the inputs once the inserted pure gaps are removed
        assert np.array_equal(np.concatenate([[16], _strip(ra[1:])]), np.concatenate([[16], _strip(a[1:])]))
        assert np.array_equal(np.concatenate([[16], _strip(rb[1:])]), np.concatenate([[16], _strip(b[1:])]))
        assert not np.any((ra[1:] == 16) & (rb[1:] == 16))  # no all-gap column
        assert np.array_equal(r["median"][q], np.concatenate([[16], _strip(r["medianwg"][q][1:])]))
        assert r["cost"][q] >= 0
    # sampled pairs against the oracle at full length
    rng = np.random.default_rng(L)
    for q in rng.choice(len(sub), size=sample, replace=False):
        p = sub[q]
        a, b = pool.seq(ia[p]), pool.seq(ib[p])
        assert c_ab[p] == port.cost_affine(pc, a, b)
        oc, om, ow, ra, rb = oracle_align(port, pc, a, b)
        assert oc == r["cost"][q] and np.array_equal(om, r["median"][q]) and np.array_equal(ow, r["medianwg"][q])
        assert np.array_equal(ra, r["res_a"][q]) and np.array_equal(rb, r["res_b"][q])
    cm.close(); pool.close()


def test_arena_waves_do_not_change_results(ctx, port):
    """a tiny direction arena forces many waves; results must be identical"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES["R1"])
    seqs, ia, ib = synth.pair_batch(5, 300, 400, frac_decorated=0.3, jitter=0.2)
    pool = pb.Pool(ctx, seqs)
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib)
    ctx.set_arena_limit(2 << 20)
    r2 = Align.align_affine_3(ctx, cm, pool, ia, ib)
    ctx.set_arena_limit(8 << 30)
    assert np.array_equal(r1["cost"], r2["cost"])
    for p in range(len(ia)):
        assert np.array_equal(r1["median"][p], r2["median"][p]) and np.array_equal(r1["res_a"][p], r2["res_a"][p])
    cm.close(); pool.close()


@pytest.mark.parametrize("rname", ["R1", "R2", "R3"])
def test_probe_fills_do_not_change_results(ctx, port, monkeypatch, rname):
    """fills without direction bytes until the stop rule fires (then the threshold is repeated with directions, from
    the snapshot of the stale EB row for pairs with gap-bit symbols): forced on for a small batch, compared with
    probes off and with the oracle"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES[rname])
    pc = port.cm(full)
    seqs, ia, ib = synth.pair_batch(21, 200, 260, frac_decorated=0.5, jitter=0.3)
    e_seqs, e_ia, e_ib = edge_pairs(31, n=120, maxlen=50)
    base = len(seqs)
    seqs = seqs + e_seqs
    keep = [p for p in range(len(e_ia)) if len(e_seqs[e_ia[p]]) <= len(e_seqs[e_ib[p]])]
    ia = np.concatenate([ia, e_ia[keep] + base]).astype(np.int32); ib = np.concatenate([ib, e_ib[keep] + base]).astype(np.int32)
    pool = pb.Pool(ctx, seqs)
    monkeypatch.setenv("POY_PROBE", "0")
    r0 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    monkeypatch.setenv("POY_PROBE", "2")
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    c1 = Align.align_affine_3(ctx, cm, pool, ia, ib, want=())["cost"]          # no traceback: every fill is a probe
    assert np.array_equal(r0["cost"], r1["cost"]) and np.array_equal(r0["cost"], c1)
    assert np.array_equal(r0["stats"][:, :3], r1["stats"][:, :3])                 # same iterations, T and k
    for p in range(len(ia)):
        for k in ("median", "medianwg", "res_a", "res_b"):
            assert np.array_equal(r0[k][p], r1[k][p]), (k, p)
    for p in range(0, len(ia), 3):
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r1["cost"][p] and np.array_equal(om, r1["median"][p]) and np.array_equal(ra, r1["res_a"][p]), p
    cm.close(); pool.close()


@pytest.mark.parametrize("rname", ["R1", "R2", "R3"])
def test_low_latency_shapes_do_not_change_results(ctx, port, monkeypatch, rname):
    """rounds with few pairs spread a band over more warps (2 / 4 diagonals per thread); forced on and off over
    band classes 64 ... 2048, gap-free and gap-bit pairs, compared with each other and with the oracle"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES[rname])
    pc = port.cm(full)
    rng = np.random.default_rng(77)
    seqs, ia, ib = synth.pair_batch(5, 60, 420, frac_decorated=0.5, jitter=0.3)
    seqs = list(seqs)
    extra_a, extra_b = [], []
    for q in range(24):                      # unrelated pairs: the band grows until it spans the matrix
        la, lb = 90 + 45 * q, 90 + 45 * q + int(rng.integers(0, 200))
        a, b = synth.random_seq(rng, la), synth.random_seq(rng, lb)
        if q % 2:
            a = synth.decorate(rng, a)
        extra_a.append(len(seqs)); seqs.append(synth.with_gap(a))
        extra_b.append(len(seqs)); seqs.append(synth.with_gap(b))
    # the band of the third fill is 4 * delta + 5 diagonals wide (k = 1.5 delta + 2): length differences that land in
    # the 1280 ... 4096 classes (16-warp shapes, mbarrier handshakes)
    for q, delta in enumerate((300, 400, 500, 600, 800, 1000)):
        la = int(1.6 * delta) + 200
        a = synth.random_seq(rng, la)
        b = np.concatenate([synth.evolve(rng, a, 0.1, 0.0), synth.random_seq(rng, delta)])
        if q % 2:
            b = synth.decorate(rng, b)
        extra_a.append(len(seqs)); seqs.append(synth.with_gap(a))
        extra_b.append(len(seqs)); seqs.append(synth.with_gap(b))
    ia = np.concatenate([ia, extra_a]).astype(np.int32); ib = np.concatenate([ib, extra_b]).astype(np.int32)
    pool = pb.Pool(ctx, seqs)
    monkeypatch.setenv("POY_LOWLAT", "0")
    r0 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    monkeypatch.setenv("POY_LOWLAT", "2")
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    assert np.array_equal(r0["cost"], r1["cost"]) and np.array_equal(r0["stats"], r1["stats"])
    assert r0["stats"][:, 2].max() > 1200            # the wide classes were exercised
    for p in range(len(ia)):
        for k in ("median", "medianwg", "res_a", "res_b"):
            assert np.array_equal(r0[k][p], r1[k][p]), (k, p)
    for p in list(range(0, 60, 4)) + list(range(60, len(ia))):
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r1["cost"][p] and np.array_equal(om, r1["median"][p]) and np.array_equal(ra, r1["res_a"][p]), p
    cm.close(); pool.close()


@pytest.mark.parametrize("rname", ["R1", "R3"])
def test_speculative_doublings_do_not_change_results(ctx, port, monkeypatch, rname):
    """latency-bound rounds fill the next threshold doublings of a pair in the same round (host.cu, align_impl): with the
    speculation off, on, on with a small CTA budget (shorter chains, more rounds) and with every third speculative fill
    forced to give up on its predecessor (the host repeats it), costs, statistics and every traceback output agree;
    sampled against the oracle"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES[rname])
    pc = port.cm(full)
    rng = np.random.default_rng(4321)
    seqs, ia, ib = synth.pair_batch(11, 28, 700, frac_decorated=0.8, jitter=0.3)
    seqs = list(seqs)
    ia, ib = list(ia), list(ib)
    for q in range(10):                      # unrelated and length-skewed pairs: long chains, bands up to the widest classes
        la = 150 + 60 * q
        a = synth.random_seq(rng, la)
        b = np.concatenate([synth.evolve(rng, a, 0.15, 0.02), synth.random_seq(rng, 80 * q)])
        if q % 3:
            a = synth.decorate(rng, a); b = synth.decorate(rng, b)
        if len(a) > len(b):
            a, b = b, a
        ia.append(len(seqs)); seqs.append(synth.with_gap(a))
        ib.append(len(seqs)); seqs.append(synth.with_gap(b))
    ia = np.asarray(ia, np.int32); ib = np.asarray(ib, np.int32)
    pool = pb.Pool(ctx, seqs)
    monkeypatch.setenv("POY_SPEC", "0")
    r0 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    rounds0 = ctx.stats()["rounds"]
    monkeypatch.delenv("POY_SPEC")
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    rounds1 = ctx.stats()["rounds"] - rounds0
    assert rounds1 < rounds0 or rounds0 <= 2          # the speculation did run
    monkeypatch.setenv("POY_SPEC_CTAS", "50")
    r2 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    # a budget so small that the first rounds run the plain schedule and only the last stragglers speculate: the state
    # the plain rounds left on the device (T, EH[0][0]) must be what the speculative fills start from, and vice versa
    monkeypatch.setenv("POY_SPEC_CTAS", "2")
    r4 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    monkeypatch.delenv("POY_SPEC_CTAS")
    rep0 = ctx.stats()["repeated"]
    monkeypatch.setenv("POY_SPEC_TEST", "1")
    r3 = Align.align_affine_3(ctx, cm, pool, ia, ib, stats=True)
    monkeypatch.delenv("POY_SPEC_TEST")
    assert ctx.stats()["repeated"] > rep0               # fills did give up and were repeated
    for r in (r1, r2, r3, r4):
        assert np.array_equal(r0["cost"], r["cost"]) and np.array_equal(r0["stats"], r["stats"])
        for p in range(len(ia)):
            for k in ("median", "medianwg", "res_a", "res_b"):
                assert np.array_equal(r0[k][p], r[k][p]), (k, p)
    for p in list(range(0, 28, 3)) + list(range(28, len(ia))):
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r1["cost"][p] and np.array_equal(om, r1["median"][p]) and np.array_equal(ow, r1["medianwg"][p])
        assert np.array_equal(ra, r1["res_a"][p]) and np.array_equal(rb, r1["res_b"][p]), p
    cm.close(); pool.close()


def test_many_wide_pairs_per_cta(ctx, port, monkeypatch):
    """more 4096-class pairs than resident CTAs: every CTA of the 16-warp shape processes several pairs in a row, with
    and without probe fills; sampled against the oracle"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES["R1"])
    pc = port.cm(full)
    rng = np.random.default_rng(123)
    seqs, ia, ib = [], [], []
    for q in range(330):
        delta = 600 + 40 * (q % 10)
        a = synth.random_seq(rng, int(1.6 * delta) + 150)
        b = np.concatenate([synth.evolve(rng, a, 0.1, 0.0), synth.random_seq(rng, delta)])
        if q % 3 == 0:
            a = synth.decorate(rng, a)
        ia.append(len(seqs)); seqs.append(synth.with_gap(a))
        ib.append(len(seqs)); seqs.append(synth.with_gap(b))
    ia = np.asarray(ia, np.int32); ib = np.asarray(ib, np.int32)
    pool = pb.Pool(ctx, seqs)
    monkeypatch.setenv("POY_PROBE", "0")
    r0 = Align.align_affine_3(ctx, cm, pool, ia, ib, want=("median",), stats=True)
    monkeypatch.setenv("POY_PROBE", "2")
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib, want=("median",), stats=True)
    assert np.array_equal(r0["cost"], r1["cost"]) and np.array_equal(r0["stats"][:, :3], r1["stats"][:, :3])
    band = (pool.lens[ib] - pool.lens[ia]) + 2 * r0["stats"][:, 2] + 1
    assert ((band > 2048) & (band <= 4096)).sum() > 300
    for p in range(len(ia)):
        assert np.array_equal(r0["median"][p], r1["median"][p]), p
    for p in range(0, len(ia), 47):
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r1["cost"][p] and np.array_equal(om, r1["median"][p]), p
    cm.close(); pool.close()


def test_two_lane_split_does_not_change_results(ctx, port, monkeypatch):
    """large batches are cut in two halves that run on two stream sets / host threads; same results, also against
    the oracle, for the affine and the linear entry point"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES["R1"])
    pc = port.cm(full)
    seqs, ia, ib = synth.pair_batch(9, 257, 300, frac_decorated=0.3, jitter=0.2)
    pool = pb.Pool(ctx, seqs)
    monkeypatch.setenv("POY_SPLIT", "0")
    r1 = Align.align_affine_3(ctx, cm, pool, ia, ib)
    monkeypatch.setenv("POY_SPLIT", "1"); monkeypatch.setenv("POY_SPLIT_MIN", "16")
    r2 = Align.align_affine_3(ctx, cm, pool, ia, ib)
    assert np.array_equal(r1["cost"], r2["cost"])
    for p in range(len(ia)):
        for k in ("median", "medianwg", "res_a", "res_b"):
            assert np.array_equal(r1[k][p], r2[k][p]), (k, p)
    for p in range(0, len(ia), 7):
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r2["cost"][p] and np.array_equal(om, r2["median"][p]) and np.array_equal(rb, r2["res_b"][p])
    lin = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 2, None).full)
    pl = port.cm(cmo.dna_matrices(1, 2, None)[0])
    r3 = Align.align_2(ctx, lin, pool, ia, ib)
    for p in range(0, len(ia), 5):
        oc, ra, rb = _oracle_linear(port, pl, seqs[ia[p]], seqs[ib[p]])
        assert oc == r3["cost"][p] and np.array_equal(ra, r3["res_a"][p]) and np.array_equal(rb, r3["res_b"][p]), p
    cm.close(); lin.close(); pool.close()


# ---- linear-gap path (algn_CAML_simple_2 / backtrace_2d) -------------------------------------------------
def _oracle_linear(port, pc, pool_seq_a, pool_seq_b, deltaw=None):
    """Sequence.Align.cost_2 / align_2 semantics (linear) on top of the oracle."""
    a, b = pool_seq_a, pool_seq_b
    sw = int(len(a) > len(b))
    s1, s2 = (b, a) if sw else (a, b)
    gaps = max(int((a & 16 != 0).sum()), int((b & 16 != 0).sum()))
    lower = int(len(s1) * 0.10)
    dif = len(s1) - len(s2)
    dcalc = (lower // 2 if dif < lower else 2) if deltaw is None else (lower if dif < lower else deltaw)
    c, r1, r2 = port.align_linear(pc, s1, s2, gaps + dcalc, sw)
    return (c, r2, r1) if sw else (c, r1, r2)


@pytest.mark.parametrize("tcm", [(1, 1), (2, 1), (1, 2), (3, 2)])
def test_linear_edge_and_related(ctx, port, tcm):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    t2d = Two_D.of_transformations_and_gaps(tcm[0], tcm[1], None)
    full, orig = cmo.dna_matrices(tcm[0], tcm[1], None)
    for host, om in ((t2d.full, full), (t2d.original, orig)):
        cm = pb.CostModel(ctx, host)
        pc = port.cm(om)
        seqs, ia, ib = edge_pairs(77 + tcm[0], n=300, maxlen=120)
        more, ja, jb = synth.pair_batch(9 + tcm[1], 60, 500, frac_decorated=0.4, jitter=0.3)
        base = len(seqs); seqs = seqs + more
        ia = np.concatenate([ia, ja + base]); ib = np.concatenate([ib, jb + base])
        pool = pb.Pool(ctx, seqs)
        for dw in (None, 7):
            cost = Align.cost_2(ctx, cm, pool, ia, ib, deltaw=dw)
            for p in range(len(ia)):
                assert cost[p] == _oracle_linear(port, pc, seqs[ia[p]], seqs[ib[p]], dw)[0], ("linear cost", p, dw)
        r = Align.align_2(ctx, cm, pool, ia, ib)
        for p in range(len(ia)):
            oc, ra, rb = _oracle_linear(port, pc, seqs[ia[p]], seqs[ib[p]])
            assert oc == r["cost"][p], ("align_2 cost", p)
            assert np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), ("align_2 rows", p)
        cm.close(); pool.close()


def test_linear_long_and_unrelated(ctx, port):
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 1, None).full)
    pc = port.cm(cmo.dna_matrices(1, 1, None)[0])
    seqs, ia, ib = synth.pair_batch(321, 10, 2500, frac_decorated=0.3, jitter=0.1)
    rng = np.random.default_rng(8)
    for p in range(4):
        seqs += [synth.with_gap(synth.random_seq(rng, 300 + 50 * p)), synth.with_gap(synth.random_seq(rng, 1500))]
        ia = np.append(ia, len(seqs) - 2); ib = np.append(ib, len(seqs) - 1)
    pool = pb.Pool(ctx, seqs)
    r = Align.align_2(ctx, cm, pool, ia, ib)
    for p in range(len(ia)):
        oc, ra, rb = _oracle_linear(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r["cost"][p] and np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), p
    cm.close(); pool.close()


# ---- column-wise helpers and the SeqCS.DOS mirror -------------------------------------------------------
@pytest.mark.parametrize("go", [None, 3])
def test_columnwise_helpers(ctx, port, go):
    import poy5_b200 as pb
    from poy5_b200 import sequence
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(2, 1, go).full)
    pc = port.cm(cmo.dna_matrices(2, 1, go)[0])
    seqs, ia, ib = edge_pairs(5, n=200, maxlen=70)
    pool = pb.Pool(ctx, seqs)
    r = Align.align_2(ctx, cm, pool, ia, ib)
    ra, rb = r["res_a"], r["res_b"]
    for wg in (False, True):
        got = sequence.median_2(ctx, cm, ra, rb, wg)
        for p in range(len(ia)):
            assert np.array_equal(got[p], port.median_2(pc, ra[p], rb[p], wg)), ("median_2", wg, p)
    u = sequence.union(ctx, ra, rb)
    anc = sequence.ancestor_2(ctx, cm, ra, rb)
    worst = sequence.aligned_cost(ctx, cm, ra, rb, True)
    ver = sequence.aligned_cost(ctx, cm, ra, rb, False)
    for p in range(len(ia)):
        assert np.array_equal(u[p], port.union(ra[p], rb[p]))
        assert np.array_equal(anc[p], port.ancestor_2(pc, ra[p], rb[p])), ("ancestor_2", p)
        assert worst[p] == port.worst_2(pc, ra[p], rb[p]) and ver[p] == port.verify_2(pc, ra[p], rb[p])
    cm.close(); pool.close()


@pytest.mark.parametrize("go", [None, 3])
def test_dos_median_and_distance(ctx, port, go):
    """SeqCS.DOS.median / distance semantics incl. the empty-sequence rules (src/seqCS.ml:705-709,991-1039)"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import DOS, Heuristic
    t2d = Two_D.of_transformations_and_gaps(1, 2, go)
    full, orig = cmo.dna_matrices(1, 2, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    pf, po = port.cm(full), port.cm(orig)
    seqs, ia, ib = edge_pairs(23, n=150, maxlen=60)
    pool = pb.Pool(ctx, seqs)
    d = DOS.distance(ctx, h, pool, ia, ib, missing_distance=0)
    m = DOS.median(ctx, h, pool, ia, ib)
    for p in range(len(ia)):
        a, b = seqs[ia[p]], seqs[ib[p]]
        ea, eb = bool((a == 16).all()), bool((b == 16).all())
        if ea or eb:
            assert d[p] == 0
            assert np.array_equal(m["sequence"][p], b if ea else a) and m["cost2"][p] == 0
            continue
        if go is None:
            assert d[p] == _oracle_linear(port, po, a, b, max(abs(len(a) - len(b)), 8))[0]
            oc, ra, rb = _oracle_linear(port, pf, a, b)
            assert np.array_equal(m["sequence"][p], port.ancestor_2(pf, ra, rb))
            assert np.array_equal(m["median_wg"][p], port.median_2(pf, ra, rb, True))
        else:
            assert d[p] == port.cost_affine(po, a, b)
            oc, om, ow, ra, rb = oracle_align(port, pf, a, b)
            assert np.array_equal(m["sequence"][p], om) and np.array_equal(m["median_wg"][p], ow)
        assert m["cost2"][p] == oc and m["cost2_max"][p] == port.worst_2(pf, ra, rb)
        assert np.array_equal(m["aligned_a"][p], ra) and np.array_equal(m["aligned_b"][p], rb)
    pool.close()


@pytest.mark.parametrize("go", [None, 3])
def test_closest_and_to_single(ctx, port, go):
    """Sequence.Align.closest / DOS.to_single (src/sequence.ml:1180-1237, src/seqCS.ml:950-982): single assignment of
    an ambiguous node sequence given its parent; identical sequences, empty sequences and gap-bit symbols included"""
    import poy5_b200 as pb
    from poy5_b200 import sequence
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import Heuristic, to_single
    from tests.helpers import oracle_closest
    t2d = Two_D.of_transformations_and_gaps(1, 2, go)
    full, orig = cmo.dna_matrices(1, 2, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    pf = port.cm(full)
    rng = np.random.default_rng(77)
    seqs = []
    n = 120
    for t in range(n):
        anc = synth.random_seq(rng, int(rng.integers(1, 150)))
        par = synth.evolve(rng, anc, 0.08, 0.03)
        mine = synth.decorate(rng, synth.evolve(rng, anc, 0.1, 0.04), 0.25, 0.1)   # ambiguity codes + gap bits
        if t % 9 == 0:
            par = mine.copy()                                   # identical sequences: the `comp` branch
        if t % 10 == 3:
            mine = np.zeros(0, np.uint8)                        # empty `mine`
        if t % 10 == 6:
            par = np.zeros(0, np.uint8)                         # empty parent: to_single aligns mine with itself
        seqs += [synth.with_gap(par), synth.with_gap(mine)]
    pool = pb.Pool(ctx, seqs)
    ip = np.arange(0, 2 * n, 2, dtype=np.int32); im = ip + 1
    lin = lambda a, b: _oracle_linear(port, pf, a, b)
    got, cost = sequence.closest(ctx, h.c2_full, pool, ip, im)
    for t in range(n):
        es, ec = oracle_closest(port, cmo, pf, full, seqs[ip[t]], seqs[im[t]], lin)
        assert np.array_equal(got[t], es) and cost[t] == ec, ("closest", t)
    got, cost = to_single(ctx, h, pool, ip, im)
    for t in range(n):
        par = seqs[ip[t]] if not bool((seqs[ip[t]] == 16).all()) else seqs[im[t]]
        es, ec = oracle_closest(port, cmo, pf, full, par, seqs[im[t]], lin)
        assert np.array_equal(got[t], es) and cost[t] == ec, ("to_single", t)
    pool.close()


@pytest.mark.parametrize("go", [None, 3])
def test_dos_readjust(ctx, port, go):
    """SeqCS.DOS.readjust (src/seqCS.ml:820-947, `ApproxD): the five cases of (is_empty ch1, is_empty ch2, is_empty parent),
    the `changed` verdict against matching and non-matching previous costs, batched; one node at a time on the oracle"""
    import poy5_b200 as pb
    from poy5_b200 import seqcs
    from poy5_b200.cost_matrix import Two_D
    from tests.helpers import oracle_dos_readjust
    t2d = Two_D.of_transformations_and_gaps(1, 1, go)
    full, _ = cmo.dna_matrices(1, 1, go)
    cm = pb.CostModel(ctx, t2d.full)
    h = seqcs.Heuristic(cm, cm)
    pf = port.cm(full)
    rng = np.random.default_rng(17)
    gapseq = lambda: synth.with_gap(np.zeros(0, np.uint8))
    seqs, n = [], 42
    for t in range(n):
        anc = synth.random_seq(rng, int(rng.integers(5, 100)))
        quad = [synth.evolve(rng, anc, 0.1, 0.04) for _ in range(4)]           # ch1, ch2, parent, mine
        if t % 4 == 0:
            quad[0] = synth.decorate(rng, quad[0], 0.1, 0.1)
        quad = [synth.with_gap(x) for x in quad]
        # which of (ch1, ch2, parent) are empty: every pattern, half of the nodes complete
        empt = [(), (0,), (1,), (2,), (0, 1), (0, 2), (1, 2)][t % 7] if t % 14 < 7 else ((0, 1, 2) if t % 14 == 13 else ())
        for slot in empt:
            quad[slot] = gapseq()
        seqs += quad
    pool = pb.Pool(ctx, seqs)
    ia = np.arange(0, 4 * n, 4, dtype=np.int32)
    ch_sum = rng.integers(0, 50, n).astype(np.int64)
    lin = lambda x, y: _oracle_linear(port, pf, x, y)

    def dist(a, b):
        if bool((np.asarray(a) == 16).all()) or bool((np.asarray(b) == 16).all()):
            return 0
        if go is not None:
            return int(port.cost_affine(pf, a, b)) if len(a) <= len(b) else int(port.cost_affine(pf, b, a))
        return int(_oracle_linear(port, pf, a, b, deltaw=max(abs(len(a) - len(b)), 8))[0])
    # pass 1 with arbitrary previous costs, pass 2 with the costs / sequences pass 1 produced (nothing may change then)
    prev = np.stack([rng.integers(0, 30, n), rng.integers(0, 60, n), rng.integers(0, 90, n)], 1).astype(np.int64)
    got = seqcs.readjust(ctx, h, pool, ia, ia + 1, ia + 2, ia + 3, prev, ch_sum)
    for t in range(n):
        q = [seqs[ia[t] + k] for k in range(4)]
        ch, new, rec, c2, mx, c3, sm, amp = oracle_dos_readjust(port, cmo, pf, full, q[0], q[1], q[2], q[3], prev[t], int(ch_sum[t]), dist, lin)
        assert bool(got["changed"][t]) == bool(ch) and got["from_record"][t] == rec, ("verdict", t)
        assert np.array_equal(got["sequence"][t], new), ("sequence", t)
        assert (got["cost2"][t], got["cost2_max"][t], got["cost3"][t], got["sum_cost"][t]) == (c2, mx, c3, sm), ("costs", t)
        assert (amp is None and got["aligned"][t] is None) or np.array_equal(got["aligned"][t], amp), ("aligned", t)
    seqs2 = list(seqs)
    for t in range(n):
        seqs2[ia[t] + 3] = got["sequence"][t]
    pool2 = pb.Pool(ctx, seqs2)
    prev2 = np.stack([got["cost2"], got["cost3"], got["sum_cost"]], 1)
    again = seqcs.readjust(ctx, h, pool2, ia, ia + 1, ia + 2, ia + 3, prev2, ch_sum)
    stable = [t for t in range(n) if again["from_record"][t] < 0]
    # (a node that copies a record keeps cost2 = cost3 = 0 in the verdict, see the reference's match)
    for t in stable:
        if np.array_equal(again["sequence"][t], got["sequence"][t]) and again["cost2"][t] == got["cost2"][t] and again["cost3"][t] == got["cost3"][t]:
            assert not again["changed"][t], ("fixed point", t)
    pool.close(); pool2.close(); cm.close()


@pytest.mark.parametrize("go", [None, 3])
def test_readjust(ctx, port, go):
    """Sequence.readjust (src/sequence.ml:2097-2156): approximate three-way re-optimisation of an interior node,
    a batched composition of 9 alignments + closest per node"""
    import poy5_b200 as pb
    from poy5_b200 import sequence
    from poy5_b200.cost_matrix import Two_D
    from tests.helpers import oracle_readjust
    t2d = Two_D.of_transformations_and_gaps(1, 1, go)
    full, _ = cmo.dna_matrices(1, 1, go)
    cm = pb.CostModel(ctx, t2d.full)
    pf = port.cm(full)
    rng = np.random.default_rng(5)
    seqs = []
    n = 40
    for t in range(n):
        anc = synth.random_seq(rng, int(rng.integers(5, 120)))
        tri = [synth.evolve(rng, anc, 0.1, 0.04) for _ in range(3)]
        if t % 3 == 0:
            tri[2] = synth.decorate(rng, tri[2], 0.1, 0.1)
        seqs += [synth.with_gap(x) for x in tri]
    pool = pb.Pool(ctx, seqs)
    ia = np.arange(0, 3 * n, 3, dtype=np.int32)
    got = sequence.readjust(ctx, cm, pool, ia, ia + 1, ia + 2)
    lin = lambda x, y: _oracle_linear(port, pf, x, y)
    for t in range(n):
        c3, c2, new, amp = oracle_readjust(port, cmo, pf, full, seqs[ia[t]], seqs[ia[t] + 1], seqs[ia[t] + 2], lin)
        assert got["cost3"][t] == c3 and got["cost2"][t] == c2, ("readjust cost", t)
        assert np.array_equal(got["sequence"][t], new) and np.array_equal(got["aligned_mp"][t], amp), ("readjust seq", t)
    pool.close(); cm.close()


@pytest.mark.parametrize("go", [None, 3])
def test_median_3_union(ctx, port, go):
    """config #3's live path: parent x union(children) alignment + median_2"""
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.seqcs import DOS, Heuristic, median_3_union
    t2d = Two_D.of_transformations_and_gaps(1, 1, go)
    full, orig = cmo.dna_matrices(1, 1, go)
    h = Heuristic(pb.CostModel(ctx, t2d.full), pb.CostModel(ctx, t2d.original))
    pf = port.cm(full)
    rng = np.random.default_rng(12)
    seqs = []
    nt = 60
    for t in range(nt):
        anc = synth.random_seq(rng, int(rng.integers(20, 220)))
        par = synth.evolve(rng, anc, 0.08, 0.03)
        c1 = synth.evolve(rng, anc, 0.1, 0.04); c2 = synth.evolve(rng, anc, 0.1, 0.04)
        if t % 4 == 0:
            par = synth.decorate(rng, par, 0.1, 0.1)
        seqs += [synth.with_gap(par), synth.with_gap(c1), synth.with_gap(c2)]
    pool = pb.Pool(ctx, seqs)
    ip = np.arange(0, 3 * nt, 3, dtype=np.int32); i1 = ip + 1; i2 = ip + 2
    node = DOS.median(ctx, h, pool, i1, i2)
    got = median_3_union(ctx, h.c2_full, pool, ip, node["aligned_a"], node["aligned_b"])
    for t in range(nt):
        u = port.union(node["aligned_a"][t], node["aligned_b"][t])
        p = seqs[ip[t]]
        if go is None:
            oc, ra, rb = _oracle_linear(port, pf, p, u)
        else:
            oc, _, _, ra, rb = oracle_align(port, pf, p, u)
        assert got["cost"][t] == oc
        assert np.array_equal(got["sequence"][t], port.median_2(pf, ra, rb, False))
        assert got["cost_max"][t] == port.worst_2(pf, ra, rb)
    pool.close()


# ---- fallback kernels, custom prepend/tail tables ------------------------------------------------------------
def test_generic_fallback_kernels(ctx, port, monkeypatch):
    """the any-width fallback (k_band_generic / k_band_lin_generic) must agree with the register kernels"""
    monkeypatch.setenv("POY_FORCE_GENERIC", "1")
    seqs, ia, ib = edge_pairs(91, n=150, maxlen=60)
    more, ja, jb = synth.pair_batch(92, 12, 700, frac_decorated=0.4, jitter=0.3)
    base = len(seqs); seqs = seqs + more
    ia = np.concatenate([ia, ja + base]); ib = np.concatenate([ib, jb + base])
    check_batch(ctx, port, REGIMES["R2"], seqs, ia, ib)
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    cm = pb.CostModel(ctx, Two_D.of_transformations_and_gaps(1, 2, None).full)
    pc = port.cm(cmo.dna_matrices(1, 2, None)[0])
    pool = pb.Pool(ctx, seqs)
    r = Align.align_2(ctx, cm, pool, ia, ib)
    for p in range(len(ia)):
        oc, ra, rb = _oracle_linear(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r["cost"][p] and np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), p
    cm.close(); pool.close()


@pytest.mark.parametrize("go", [None, 4])
def test_custom_prepend_tail_and_nonmetric(ctx, port, go):
    """transform(prepend:..., tail:...) and a non-metric input matrix: the affine kernels take their column gap
    costs from prepend_cost, the linear ones use prepend / tail on row 0, column 0 and the last column"""
    import ctypes
    import poy5_b200 as pb
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.sequence import Align
    rows = [[0, 2, 1, 2, 3], [2, 0, 2, 1, 3], [1, 2, 0, 2, 2], [2, 1, 2, 0, 3], [3, 3, 2, 3, 0]]
    t2d = Two_D.of_list(rows, go)
    full, _ = cmo.of_list(rows)
    if go is not None:
        full = cmo.set_cost_model(full.clone(), 1, go)
    rng = np.random.default_rng(2)
    pre = rng.integers(0, 5, size=32).astype(np.int32); tail = rng.integers(0, 5, size=32).astype(np.int32)
    pre[0] = tail[0] = 0
    full.prepend[:] = pre; full.tail[:] = tail
    for a in range(32):
        t2d.full.prepend[a] = int(pre[a]); t2d.full.tail[a] = int(tail[a])
    cm = pb.CostModel(ctx, t2d.full)
    pc = port.cm(full)
    seqs, ia, ib = edge_pairs(61, n=250, maxlen=80)
    pool = pb.Pool(ctx, seqs)
    if go is None:
        r = Align.align_2(ctx, cm, pool, ia, ib)
        for p in range(len(ia)):
            oc, ra, rb = _oracle_linear(port, pc, seqs[ia[p]], seqs[ib[p]])
            assert oc == r["cost"][p] and np.array_equal(ra, r["res_a"][p]) and np.array_equal(rb, r["res_b"][p]), p
    else:
        cost = Align.cost_2(ctx, cm, pool, ia, ib)
        r = Align.align_affine_3(ctx, cm, pool, ia, ib)
        for p in range(len(ia)):
            assert cost[p] == port.cost_affine(pc, seqs[ia[p]], seqs[ib[p]]), p
            oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
            assert oc == r["cost"][p] and np.array_equal(om, r["median"][p]) and np.array_equal(ra, r["res_a"][p]), p
    cm.close(); pool.close()


def test_long_pair_over_16k(ctx, port):
    """len1 + len2 > 16382 needs --enable-long-sequences in the reference; no such limit here"""
    import poy5_b200 as pb
    from poy5_b200.sequence import Align
    cm, full = setup(ctx, REGIMES["R1"])
    pc = port.cm(full)
    seqs, ia, ib = synth.pair_batch(404, 2, 12000, jitter=0.02)
    pool = pb.Pool(ctx, seqs)
    cost = Align.cost_2(ctx, cm, pool, ia, ib)
    r = Align.align_affine_3(ctx, cm, pool, ia, ib)
    for p in range(2):
        assert cost[p] == port.cost_affine(pc, seqs[ia[p]], seqs[ib[p]])
        oc, om, ow, ra, rb = oracle_align(port, pc, seqs[ia[p]], seqs[ib[p]])
        assert oc == r["cost"][p] and np.array_equal(om, r["median"][p]) and np.array_equal(ra, r["res_a"][p])
    cm.close(); pool.close()
