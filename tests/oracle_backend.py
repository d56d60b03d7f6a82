"""TEST INFRASTRUCTURE: the treesearch backend interface implemented on the CPU oracle (SeqCS.DOS.median /
DOS.distance semantics, src/seqCS.ml:701-774, 985-1084), used to replay the GPU driver's call sequence."""
import numpy as np


class OracleBackend:
    def __init__(self, port, full, orig):
        self.P, self.full, self.orig = port, full, orig
        self.pf, self.po = port.cm(full), port.cm(orig)
        self.affine = full.cost_model_type == 1

    @staticmethod
    def _empty(s):
        return bool((np.asarray(s) == 16).all())

    def _lin(self, pc, a, b, deltaw):
        sw = int(len(a) > len(b))
        s1, s2 = (b, a) if sw else (a, b)
        gaps = max(int((np.asarray(a) & 16 != 0).sum()), int((np.asarray(b) & 16 != 0).sum()))
        lower = int(len(s1) * 0.10); dif = len(s1) - len(s2)
        dcalc = (lower // 2 if dif < lower else 2) if deltaw is None else (lower if dif < lower else deltaw)
        c, r1, r2 = self.P.align_linear(pc, s1, s2, gaps + dcalc, sw)
        return (c, r2, r1) if sw else (c, r1, r2)

    def median(self, pairs):
        out = []
        for a, b in pairs:
            if self._empty(a):
                out.append((np.array(b, np.uint8), 0)); continue
            if self._empty(b):
                out.append((np.array(a, np.uint8), 0)); continue
            if self.affine:
                sw = int(len(a) > len(b))
                si, sj = (b, a) if sw else (a, b)
                c, m, _, _, _ = self.P.align_affine(self.pf, si, sj, sw)
                out.append((m, int(c)))
            else:
                c, ra, rb = self._lin(self.pf, a, b, None)
                out.append((self.P.ancestor_2(self.pf, ra, rb), int(c)))
        return out

    def distance(self, pairs):
        out = []
        for a, b in pairs:
            if self._empty(a) or self._empty(b):
                out.append(0)
            elif self.affine:
                out.append(int(self.P.cost_affine(self.po, a, b)))
            else:
                out.append(int(self._lin(self.po, a, b, max(abs(len(a) - len(b)), 8))[0]))
        return out

    def single(self, pairs):
        """DOS.to_single (src/seqCS.ml:950-982): empty `mine` stays, an empty parent is replaced by `mine`, otherwise
        Sequence.Align.closest parent mine under c2_full"""
        from oracle import cost_matrix_oracle as cmo
        from tests.helpers import oracle_closest
        out = []
        for parent, mine in pairs:
            if self._empty(parent):
                parent = mine
            s, c = oracle_closest(self.P, cmo, self.pf, self.full, parent, mine,
                                  (lambda x, y: self._lin(self.pf, x, y, None)) if not self.affine else None)
            out.append((s, int(c)))
        return out


class OracleStoreBackend:
    """The id-array interface of treesearch.StoreBackend (put / median_ids / distance_ids / mark / release / fetch) on top
    of the CPU checker: drives the vectorised swap-round driver without a GPU."""

    def __init__(self, port, full, orig):
        from poy5_b200.treesearch import Node
        self.Node = Node
        self.ob = OracleBackend(port, full, orig)
        self.seqs = []

    def put(self, seqs):
        first = len(self.seqs)
        self.seqs += [np.asarray(s, np.uint8) for s in seqs]
        return [self.Node(first + i, len(s)) for i, s in enumerate(seqs)]

    def fetch(self, nodes):
        return [self.seqs[x.id] for x in nodes]

    def mark(self):
        return len(self.seqs)

    def release(self, mark):
        del self.seqs[mark:]

    def median(self, pairs):
        res = self.ob.median([(self.seqs[a.id], self.seqs[b.id]) for a, b in pairs])
        return [(self.put([m])[0], c) for m, c in res]

    def distance(self, pairs):
        return self.ob.distance([(self.seqs[a.id], self.seqs[b.id]) for a, b in pairs])

    def median_ids(self, a, b):
        res = self.ob.median([(self.seqs[int(x)], self.seqs[int(y)]) for x, y in zip(a, b)])
        nodes = self.put([m for m, _ in res])
        return (np.array([x.id for x in nodes], np.int32), np.array([x.n for x in nodes], np.int32),
                np.array([c for _, c in res], np.int32))

    def distance_ids(self, a, b, la, lb):
        return np.array(self.ob.distance([(self.seqs[int(x)], self.seqs[int(y)]) for x, y in zip(a, b)]), np.int32)


def replay_sample(med, dis, regime):
    """CPU checker replay of a recorded sample of the swap-evaluation workload (poy5_b200.swap_eval.SampleRecorder):
    med = [(a, b, median, cost2)], dis = [(a, b, cost)] as host arrays.  -> parity dict"""
    from oracle import cost_matrix_oracle as cmo
    from oracle.port import Port
    full, orig = cmo.dna_matrices(*regime)
    ob = OracleBackend(Port(), full, orig)
    bad = 0
    for (a, b, m, c), (om, oc) in zip(med, ob.median([(a, b) for a, b, _, _ in med])):
        bad += not (np.array_equal(m, om) and c == oc)
    for (a, b, c), oc in zip(dis, ob.distance([(a, b) for a, b, _ in dis])):
        bad += int(c != oc)
    return dict(checked=len(med) + len(dis), checked_medians=len(med), checked_distances=len(dis), mismatches=int(bad),
                against="oracle port (plain-C restatement of algn.c, pinned to the compiled reference)")


def replay_triplets(sample, regime):
    """CPU checker replay of recorded (parent, child 1, child 2, cost, median) triplets of the median_3_union path
    (poy5_b200.workloads.config3): union of the aligned children, parent x union alignment, median_2"""
    from oracle import cost_matrix_oracle as cmo
    from oracle.port import Port
    from tests.helpers import oracle_align
    full, _ = cmo.dna_matrices(*regime)
    P = Port(); pf = P.cm(full)
    bad = 0
    for p, c1, c2, cost, med in sample:
        _, _, _, ra, rb = oracle_align(P, pf, c1, c2)
        u = P.union(ra, rb)
        oc, _, _, xa, xb = oracle_align(P, pf, p, u)
        bad += not (oc == cost and np.array_equal(P.median_2(pf, xa, xb, False), med))
    return dict(checked=len(sample), mismatches=int(bad), against="oracle port (plain-C restatement of algn.c, pinned to the compiled reference)")


def replay_newkk(sample, regime):
    """CPU checker replay of recorded Sequence.NewkkAlign.align_2 results (poy5_b200.workloads.newkk): (a, b, cost, row a, row b)"""
    from oracle import cost_matrix_oracle as cmo
    from oracle.port import Port
    full, _ = cmo.dna_matrices(*regime)
    P = Port(); pf = P.cm(full)
    bad = 0
    for a, b, cost, ra, rb in sample:
        sw = int(len(a) > len(b))
        s1, s2 = (b, a) if sw else (a, b)
        oc, o1, o2 = P.newkk_align(pf, s1, s2, sw)
        xa, xb = (o2, o1) if sw else (o1, o2)
        bad += not (oc == cost and np.array_equal(xa, ra) and np.array_equal(xb, rb))
    return dict(checked=len(sample), mismatches=int(bad), against="oracle/newkk_oracle.c (restatement pinned to the compiled src/newkkonen.c)")
