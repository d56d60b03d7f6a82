"""Tree-level driver for the candidate-list seam: Wagner build and one TBR round over
dynamic-homology sequence characters, written so that EVERY alignment is issued in large batches:

* all-direction medians (the three directional medians per node of ``AllDirNode``) are computed level by
  level, one ``DOS.median`` batch per level (the downpass / join path, src/allDirChar.ml:2033-2125);
* candidate evaluation is the Parmap seam of src/ptree.ml:1226-1268 (Wagner) and :1356-1453 (SPR/TBR):
  collect every (clade median, edge median) pair of every break / reroot / join edge, ONE
  ``DOS.distance`` batch (cost-only alignments under c2_original, src/allDirChar.ml:2132-2177), then a MIN
  reduction (over ranks too when the batch is sharded, see shard.py).

This is a workload driver, not a re-implementation of POY's search managers: acceptance rules, tabu lists and
lazy edges are out of scope.  The algorithms only talk to a *backend* with two methods

    median(pairs)   -> [(median_sequence, cost2), ...]      (SeqCS.DOS.median semantics)
    distance(pairs) -> [cost, ...]                           (SeqCS.DOS.distance semantics)

``GpuBackend`` below is the product path; tests replay the identical call sequence through a CPU checker
implementation of the same interface and compare tree costs bit for bit."""
import numpy as np


class GpuBackend:
    """Batches go through libpoy5b200.so (seqcs.DOS.median / DOS.distance)."""

    def __init__(self, ctx, h, lean=True):
        """lean: medians read back only what a tree pass consumes (sequence + cost2, DOS.median_cost) instead of
        the full DOS.median record (aligned children, median_wg, cost2_max); affine models only"""
        self.ctx, self.h = ctx, h
        self.lean = lean and h.c2_full.host.cost_model_type == 1
        self.n_median = self.n_distance = 0
        self.cells_distance = 0

    def _pool(self, pairs):
        from .api import Pool
        ids, seqs, ia, ib = {}, [], [], []
        for a, b in pairs:
            for s, out in ((a, ia), (b, ib)):
                k = id(s)
                if k not in ids:
                    ids[k] = len(seqs); seqs.append(s)
                out.append(ids[k])
        return Pool(self.ctx, seqs), np.asarray(ia, np.int32), np.asarray(ib, np.int32)

    def median(self, pairs):
        from .seqcs import DOS
        if not pairs:
            return []
        pool, ia, ib = self._pool(pairs)
        self.n_median += len(pairs)
        if self.lean:
            seqs, cost = DOS.median_cost(self.ctx, self.h, pool, ia, ib)
            pool.close()
            return list(zip(seqs, cost.tolist()))
        r = DOS.median(self.ctx, self.h, pool, ia, ib)
        pool.close()
        return [(np.array(r["sequence"][p], np.uint8), int(r["cost2"][p])) for p in range(len(pairs))]

    def single(self, pairs):
        """[(parent, mine)] -> [(single-assignment sequence, cost)] (seqcs.to_single)"""
        from .seqcs import to_single
        if not pairs:
            return []
        pool, ia, ib = self._pool(pairs)
        seqs, cost = to_single(self.ctx, self.h, pool, ia, ib)
        pool.close()
        return [(np.array(s, np.uint8), int(c)) for s, c in zip(seqs, cost)]

    def distance(self, pairs):
        from .seqcs import DOS
        if not pairs:
            return []
        pool, ia, ib = self._pool(pairs)
        d = DOS.distance(self.ctx, self.h, pool, ia, ib, missing_distance=0)
        self.cells_distance += int(((pool.lens[ia] - 1) * (pool.lens[ib] - 1)).sum())
        pool.close()
        self.n_distance += len(pairs)
        return [int(x) for x in d]


class Node:
    """Handle of a sequence that lives in the device-resident node store: id + length (what the drivers need on the
    host: `len(x)` for work estimates, identity for bookkeeping)."""
    __slots__ = ("id", "n")

    def __init__(self, id_, n):
        self.id, self.n = int(id_), int(n)

    def __len__(self):
        return self.n

    def __repr__(self):
        return "Node(%d, len %d)" % (self.id, self.n)


class StoreBackend:
    """The product path of a tree pass: every sequence is a `Node` handle into ONE device-resident store
    (poy_store, csrc/store.cu); `median` appends its results in HBM and returns handles, `distance` runs the cost-only
    batch on ids.  Nothing but ids, lengths and costs crosses PCIe.  `mark()` / `release(mark)` give the drivers
    stack discipline for the temporaries of a chunk of candidates."""

    def __init__(self, ctx, h, cap_bytes=1 << 26, cap_seqs=1 << 16):
        from .api import Store
        self.ctx, self.h = ctx, h
        self.store = Store(ctx, cap_bytes, cap_seqs)
        self.n_median = self.n_distance = 0
        self.cells_distance = 0

    def put(self, seqs):
        ids = self.store.append(seqs)
        return [Node(i, len(s)) for i, s in zip(ids, seqs)]

    def fetch(self, nodes):
        return self.store.read(np.fromiter((x.id for x in nodes), np.int32, len(nodes)))

    def mark(self):
        return len(self.store)

    def release(self, mark):
        self.store.truncate(mark)

    def median(self, pairs):
        if not pairs:
            return []
        n = len(pairs)
        a = np.fromiter((p[0].id for p in pairs), np.int32, n); b = np.fromiter((p[1].id for p in pairs), np.int32, n)
        ids, ln, c2 = self.store.median(self.h, a, b)
        self.n_median += n
        return [(Node(i, l), c) for i, l, c in zip(ids.tolist(), ln.tolist(), c2.tolist())]

    def distance(self, pairs):
        if not pairs:
            return []
        n = len(pairs)
        a = np.fromiter((p[0].id for p in pairs), np.int32, n); b = np.fromiter((p[1].id for p in pairs), np.int32, n)
        la = np.fromiter((p[0].n for p in pairs), np.int64, n); lb = np.fromiter((p[1].n for p in pairs), np.int64, n)
        self.cells_distance += int(((la - 1) * (lb - 1)).sum())
        self.n_distance += n
        return self.store.distance(self.h, a, b).tolist()

    # the same two calls on id arrays: what the vectorised swap-round driver (spr_round on a node store) uses, so that
    # a neighbourhood of 10^5 candidates costs numpy index arithmetic on the host instead of one Python object per sequence
    def median_ids(self, a, b):
        """int32 id arrays -> (ids, lengths, cost2) arrays of the medians, appended to the store"""
        if len(a) == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)
        self.n_median += len(a)
        return self.store.median(self.h, np.ascontiguousarray(a, np.int32), np.ascontiguousarray(b, np.int32))

    def distance_ids(self, a, b, la, lb):
        if len(a) == 0:
            return np.zeros(0, np.int32)
        self.cells_distance += int(((np.asarray(la, np.int64) - 1) * (np.asarray(lb, np.int64) - 1)).sum())
        self.n_distance += len(a)
        return self.store.distance(self.h, np.ascontiguousarray(a, np.int32), np.ascontiguousarray(b, np.int32))

    def single(self, pairs):
        """to_single is an O(n) pass per tree: sequences go through the host-array path (seqcs.to_single)"""
        from .api import Pool
        from .seqcs import to_single
        if not pairs:
            return []
        flat = self.fetch([x for p in pairs for x in p])
        pool = Pool(self.ctx, flat)
        n = len(pairs)
        seqs, cost = to_single(self.ctx, self.h, pool, np.arange(0, 2 * n, 2, dtype=np.int32), np.arange(1, 2 * n, 2, dtype=np.int32))
        pool.close()
        return [(x, int(c)) for x, c in zip(self.put([np.array(s_, np.uint8) for s_ in seqs]), cost)]

    def close(self):
        self.store.close()


class Tree:
    """Unrooted binary tree: leaves 0..n-1 hold observed sequences, internal nodes are created on insertion."""

    def __init__(self):
        self.adj = {}

    def copy(self):
        t = Tree()
        t.adj = {u: list(vs) for u, vs in self.adj.items()}
        return t

    def add_edge(self, u, v):
        self.adj.setdefault(u, []).append(v)
        self.adj.setdefault(v, []).append(u)

    def remove_edge(self, u, v):
        self.adj[u].remove(v); self.adj[v].remove(u)

    def edges(self):
        return sorted((u, v) for u in self.adj for v in self.adj[u] if u < v)

    def new_node(self, n_leaves=0):
        """id for a new internal node: above every node in the tree AND above the leaf ids 0..n_leaves-1, which may
        not have been inserted yet"""
        return max(max(self.adj) + 1, int(n_leaves))

    def insert_leaf(self, leaf, edge, n_leaves=0):
        """split `edge` with a new internal node and hang `leaf` on it"""
        u, v = edge
        w = self.new_node(max(n_leaves, leaf + 1))
        self.remove_edge(u, v); self.add_edge(u, w); self.add_edge(w, v); self.add_edge(w, leaf)
        return w

    def component(self, start, banned):
        seen, stack = {start}, [start]
        while stack:
            x = stack.pop()
            for y in self.adj[x]:
                if y != banned and y not in seen:
                    seen.add(y); stack.append(y)
        return seen


def directional_medians_multi(problems, backend):
    """dm[(u, v)] = (median, accumulated cost) of the subtree that contains u once edge (u, v) is cut, seen from
    u -- for SEVERAL (tree, leaf_seq, nodes) problems in lockstep.  Level-synchronous: every round issues ONE
    median batch with all directed edges, of all problems, whose two inputs are ready (AllDirNode's three lazy
    medians per node, evaluated eagerly and in bulk)."""
    dms, pend = [], []
    for tree, leaf_seq, nodes in problems:
        nodes = set(tree.adj) if nodes is None else nodes
        dm, pending = {}, []
        for u in nodes:
            for v in tree.adj[u]:
                if v not in nodes:
                    continue
                others = [w for w in tree.adj[u] if w != v and w in nodes]
                if len(others) == 0:
                    dm[(u, v)] = (leaf_seq[u], 0)
                else:                       # 1 other: a degree-2 node left behind by a break passes its clade through
                    pending.append((u, v, others))
        dms.append(dm); pend.append(pending)
    while any(pend):
        batch, owners = [], []
        progressed = False
        for k, pending in enumerate(pend):
            dm = dms[k]
            ready = [x for x in pending if all((w, x[0]) in dm for w in x[2])]
            if not ready:
                continue
            progressed = True
            pend[k] = [x for x in pending if not all((w, x[0]) in dm for w in x[2])]
            for u, v, o in ready:
                if len(o) == 2:
                    batch.append((dm[(o[0], u)][0], dm[(o[1], u)][0])); owners.append((k, u, v, o))
            for u, v, o in ready:            # pass-throughs only depend on entries that were already there
                if len(o) == 1:
                    dm[(u, v)] = dm[(o[0], u)]
        if not progressed:
            raise RuntimeError("cyclic dependency in directional medians")
        for (k, u, v, o), (seq, c2) in zip(owners, backend.median(batch)):
            dms[k][(u, v)] = (seq, c2 + dms[k][(o[0], u)][1] + dms[k][(o[1], u)][1])
    return dms


def directional_medians(tree, leaf_seq, backend, nodes=None):
    return directional_medians_multi([(tree, leaf_seq, nodes)], backend)[0]


def edge_medians_multi(problems, dms, backend):
    """median + total cost for every edge taken as the root (the data `cost_fn` compares a clade against), for
    several (tree, edges) problems in ONE batch"""
    batch, owners = [], []
    for k, (tree, edges) in enumerate(problems):
        for (u, v) in edges:
            batch.append((dms[k][(u, v)][0], dms[k][(v, u)][0])); owners.append((k, (u, v)))
    out = [dict() for _ in problems]
    for (k, e), (seq, c2) in zip(owners, backend.median(batch)):
        out[k][e] = (seq, c2 + dms[k][(e[0], e[1])][1] + dms[k][(e[1], e[0])][1])
    return out


def edge_medians(tree, dm, backend, edges=None):
    edges = tree.edges() if edges is None else edges
    return edge_medians_multi([(tree, edges)], [dm], backend)[0]


def tree_cost(tree, leaf_seq, backend):
    """cost of the tree rooted on its first edge (sum of cost2 over the downpass + the root median)"""
    dm = directional_medians(tree, leaf_seq, backend)
    e = tree.edges()[0]
    return edge_medians(tree, dm, backend, [e])[e][1]


def wagner_build(leaf_seq, backend, order=None):
    """sequential-addition build (src/ptree.ml:1144-1270): for every new taxon ONE distance batch over all edges"""
    n = len(leaf_seq)
    order = list(range(n)) if order is None else list(order)
    tree = Tree()
    tree.add_edge(order[0], order[1])
    seqs = dict(enumerate(leaf_seq))
    for t in order[2:]:
        dm = directional_medians(tree, seqs, backend)
        em = edge_medians(tree, dm, backend)
        edges = tree.edges()
        d = backend.distance([(leaf_seq[t], em[e][0]) for e in edges])
        best = int(np.argmin(d))                    # ties: first edge in sorted order
        w = max(max(tree.adj) + 1, n)
        u, v = edges[best]
        tree.remove_edge(u, v)
        tree.add_edge(u, w); tree.add_edge(w, v); tree.add_edge(w, t)
    return tree


def tbr_round(tree, leaf_seq, backend, reduce_best=None):
    """One TBR neighbourhood, evaluated speculatively for ALL break edges at once (multi-break
    batching).  For each break: all-direction medians inside the two components, then every
    (reroot edge of the pruned clade) x (join edge of the rest) pair goes into one global distance batch.
    Returns (best_estimate, move, n_candidates); move = (break_edge, clade_edge, join_edge) or None."""
    seqs = dict(enumerate(leaf_seq))
    breaks = []
    for (u, v) in tree.edges():
        t = tree.copy()
        t.remove_edge(u, v)
        A, B = t.component(u, None), t.component(v, None)
        if len([x for x in A if x < len(leaf_seq)]) < 1 or len([x for x in B if x < len(leaf_seq)]) < 1:
            continue
        breaks.append(((u, v), t, A, B))
    # medians of both components of every break: all (break, side) problems advance in lockstep, so each tree
    # level costs ONE median batch for the whole neighbourhood
    probs, where = [], []
    for bi, (brk, t, A, B) in enumerate(breaks):
        for side, comp in enumerate((A, B)):
            if len(comp) > 1:
                probs.append((t, seqs, comp)); where.append((bi, side))
    dms = directional_medians_multi(probs, backend)
    eprobs = [(t, [(a, b) for (a, b) in t.edges() if a in comp and b in comp]) for (t, _, comp) in probs]
    ems = edge_medians_multi(eprobs, dms, backend)
    sides_of = {}
    for (bi, side), em in zip(where, ems):
        sides_of[(bi, side)] = em
    cand, meta = [], []
    for bi, (brk, t, A, B) in enumerate(breaks):
        sides = []
        for side, comp in enumerate((A, B)):
            if len(comp) == 1:
                sides.append({None: (seqs[next(iter(comp))], 0)})
            else:
                sides.append(sides_of[(bi, side)])
        for ea, (sa, ca) in sides[0].items():
            for eb, (sb, cb) in sides[1].items():
                cand.append((sa, sb)); meta.append((brk, ea, eb, ca + cb))
    d = backend.distance(cand)
    est = np.array([dd + m[3] for dd, m in zip(d, meta)], np.int64)
    if len(est) == 0:
        return None, None, 0
    k = int(np.argmin(est))
    best = (int(est[k]), k)
    if reduce_best is not None:
        best = reduce_best(best)
    return best[0], meta[best[1]][:3], len(cand)


def apply_tbr(tree, move):
    """reconnect: remove the break edge, suppress the two degree-2 nodes, join the midpoints of the two edges"""
    (u, v), ea, eb = move
    t = tree.copy()
    t.remove_edge(u, v)

    def suppress(x):
        if x in t.adj and len(t.adj[x]) == 2:
            a, b = t.adj[x]
            t.remove_edge(x, a); t.remove_edge(x, b); del t.adj[x]
            t.add_edge(a, b)
            return (x, a, b)
        return None
    su, sv = suppress(u), suppress(v)

    def fix(e, s):
        # an edge of the component that touched the suppressed node now runs between its two neighbours
        if e is None or s is None:
            return e
        x, a, b = s
        if x in e:
            return tuple(sorted((a, b)))
        return e
    ea, eb = fix(ea, su), fix(eb, sv)
    nxt = max(t.adj) + 1

    def midpoint(e, lone):
        nonlocal nxt
        if e is None:
            return lone
        a, b = e
        w = nxt; nxt += 1
        t.remove_edge(a, b); t.add_edge(a, w); t.add_edge(w, b)
        return w
    lone_u = u if u in t.adj else None
    lone_v = v if v in t.adj else None
    ma = midpoint(ea, lone_u)
    mb = midpoint(eb, lone_v)
    t.add_edge(ma, mb)
    return t


class ShardedBackend:
    """N>1: candidates are independent, so `distance` batches are dealt to the ranks by estimated cells (LPT),
    every rank evaluates its shard and the per-candidate costs are combined with ONE all-reduce (each rank
    contributes its own entries, zeros elsewhere).  Medians (the downpass) are replicated by default -- every rank
    needs every node sequence for the next level anyway -- or, for wide levels, sharded and all-gathered
    (`min_shard_medians`)."""

    def __init__(self, backend, device=None, min_shard_medians=1 << 30):
        import torch.distributed as dist
        self.b, self.device, self.min_shard_medians = backend, device, min_shard_medians
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def median(self, pairs):
        """Batches of `min_shard_medians` pairs or more are dealt to the ranks like the candidates; the new medians
        (sequence + cost2, padded rows) are exchanged with ONE all-gather per tree level."""
        import torch
        import torch.distributed as dist
        from . import shard
        if self.world == 1 or len(pairs) < self.min_shard_medians:
            return self.b.median(pairs)
        work = [(len(a) - 1) * (len(b) - 1) for a, b in pairs]
        parts = shard.lpt_partition(work, self.world)
        mine = parts[self.rank]
        res = self.b.median([pairs[i] for i in mine])
        cap = 16 + max(len(a) + len(b) for a, b in pairs)
        rows = max(len(p) for p in parts)
        buf = np.zeros((rows, cap), np.uint8)
        for q, (seq, c2) in enumerate(res):
            buf[q, :16].view(np.int64)[:] = (len(seq), c2)
            buf[q, 16:16 + len(seq)] = seq
        t = torch.from_numpy(buf).to(self.device) if self.device is not None else torch.from_numpy(buf)
        got = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(got, t)
        out = [None] * len(pairs)
        for r, g in enumerate(got):
            g = g.cpu().numpy()
            for q, i in enumerate(parts[r]):
                ln, c2 = g[q, :16].view(np.int64)
                out[int(i)] = (g[q, 16:16 + int(ln)].copy(), int(c2))
        return out

    def single(self, pairs):
        return self.b.single(pairs)       # O(n) pairs per pass: replicated

    def distance(self, pairs):
        import torch
        import torch.distributed as dist
        from . import shard
        if self.world == 1 or not pairs:
            return self.b.distance(pairs)
        work = [(len(a) - 1) * (len(b) - 1) for a, b in pairs]
        mine = shard.lpt_partition(work, self.world)[self.rank]
        out = torch.zeros(len(pairs), dtype=torch.int64, device=self.device)
        if len(mine):
            vals = self.b.distance([pairs[i] for i in mine])
            out[torch.as_tensor(mine, device=self.device)] = torch.as_tensor(vals, dtype=torch.int64, device=self.device)
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return [int(x) for x in out.cpu().tolist()]


# ------------------------------------------------------------------------------------------------------------
# SPR round over several loci with incremental medians (BASELINE configs #4 / #5)
# ------------------------------------------------------------------------------------------------------------

def all_directions(tree, loci, backend):
    """All-direction medians of one tree for every locus (`loci[l][taxon]`), all loci in lockstep: one median
    batch per tree level for the whole data set (the Array_ops.map over loci of src/seqCS.ml:2232-2345 turned
    inside out)."""
    return directional_medians_multi([(tree, dict(enumerate(ls)), None) for ls in loci], backend)


def downpass(tree, loci, backend, root=None):
    """Post-order pass towards the root edge only (n-2 medians + the root median per locus, level-synchronous;
    src/allDirChar.ml:2033-2125).  Returns (total cost over loci, per-locus dict of towards-root medians)."""
    root = tree.edges()[0] if root is None else root
    parent, order = {root[0]: root[1], root[1]: root[0]}, [root[0], root[1]]
    for x in order:
        for y in tree.adj[x]:
            if y not in parent:
                parent[y] = x; order.append(y)
    kids = {x: [y for y in tree.adj[x] if parent.get(y) == x and parent[x] != y] for x in order}
    height = {}
    for x in reversed(order):
        height[x] = 1 + max((height[k] for k in kids[x]), default=-1)
    dms = [{(x, parent[x]): (ls[x], 0) for x in order if not kids[x]} for ls in loci]
    for h in range(1, max(height.values()) + 1):
        lvl = [x for x in order if height[x] == h]
        res = backend.median([(dm[(kids[x][0], x)][0], dm[(kids[x][1], x)][0]) for dm in dms for x in lvl])
        for k, dm in enumerate(dms):
            for q, x in enumerate(lvl):
                seq, c2 = res[k * len(lvl) + q]
                dm[(x, parent[x])] = (seq, c2 + dm[(kids[x][0], x)][1] + dm[(kids[x][1], x)][1])
    a, b = root
    res = backend.median([(dm[(a, b)][0], dm[(b, a)][0]) for dm in dms])
    total = sum(c2 + dm[(a, b)][1] + dm[(b, a)][1] for (seq, c2), dm in zip(res, dms))
    return int(total), dms


def spr_prunings(tree, n_leaves):
    """Every (u, v): cut edge u-v, prune the clade on v's side, rejoin it inside the rest (u's side).  Prunings whose
    rest has fewer than three leaves have no alternative position and are left out."""
    out = []
    for (a, b) in tree.edges():
        for (u, v) in ((a, b), (b, a)):
            rest = tree.component(u, v)
            if sum(1 for x in rest if x < n_leaves) >= 3:
                out.append((u, v))
    return out


def _even_chunk(n, chunk):
    """`chunk` is an upper bound (memory); the prunings are cut into equal chunks, because every chunk pays the same
    latency-bound chain of median levels however few prunings share it"""
    if n <= 0:
        return max(1, chunk)
    k = (n + chunk - 1) // chunk
    return (n + k - 1) // k


def spr_round(tree, loci, backend, dms=None, prunings=None, chunk=64, where=None, merge_edges=False):
    """One SPR neighbourhood over several loci.  For the pruning (u, v) the rest tree is u's side with u suppressed
    (its neighbours x1, x2 joined); only the medians that *see* the cut are recomputed, top-down from the cut:

        up[x1] = dm[(x2, u)], up[x2] = dm[(x1, u)]                  (reused from the unbroken tree)
        up[c]  = median(up[a], dm[(s, a)])    for the children c, s of a (away from the cut), level by level
        join edge a-c:  em = median(up[c], dm[(c, a)]);  estimate = distance(dm[(v, u)], em) + both accumulated costs

    which is what AllDirNode's lazy directional medians recompute after a break (src/allDirChar.ml:2132-2204) and
    the Parmap candidate seam of src/ptree.ml:1356-1453 evaluates.  All prunings of a chunk and all loci advance
    in lockstep: one median batch per level, one edge-median batch, ONE distance batch per chunk; the estimate of
    a candidate is the sum over loci.  Returns (best estimate, (pruning, join edge), number of candidates,
    number of alignments issued); ties resolve to the first candidate in enumeration order.  `where` (a list) receives
    (index of the winning pruning in `prunings`, index of the join edge within that pruning)."""
    n = len(loci[0])
    dms = all_directions(tree, loci, backend) if dms is None else dms
    prunings = spr_prunings(tree, n) if prunings is None else list(prunings)
    nl = len(loci)
    if hasattr(backend, "median_ids"):
        return _spr_round_ids(tree, loci, backend, dms, prunings, chunk, where, merge_edges)
    best, ncand, naln = None, 0, 0
    chunk = _even_chunk(len(prunings), chunk)
    for c0 in range(0, len(prunings), chunk):
        part = prunings[c0:c0 + chunk]
        mark = backend.mark() if hasattr(backend, "mark") else None     # node store: the chunk's medians are temporaries
        ups, levels, joins = [], [], []       # per pruning: up[l][node]; nodes by depth; join edges (a, c)
        for (u, v) in part:
            x1, x2, lv, par = _side_plan(tree, u, v)
            up = [{x1: dm[(x2, u)], x2: dm[(x1, u)]} for dm in dms]
            ups.append(up); levels.append((lv, par))
            joins.append([(par[c][0], c) for l_ in lv for c in l_])
        # merge_edges: level d of the incremental medians and the edge medians of the join edges whose up[] became
        # available at level d-1 go into ONE batch (filler for the latency-bound dependent chain when the chunk is small;
        # measured slower for large chunks, where one big edge-median batch keeps every round of the band schedule full)
        ems = {}
        depth = max(len(lv) for lv, _ in levels)
        for d in range(depth + 1):
            batch, own = [], []
            for k, (lv, par) in enumerate(levels):
                if d < len(lv):
                    for c in lv[d]:
                        a, s = par[c]
                        for l in range(nl):
                            batch.append((ups[k][l][a][0], dms[l][(s, a)][0])); own.append((0, k, l, c, a, s))
                if (merge_edges and 0 < d <= len(lv)) or (not merge_edges and d == depth):
                    for c in (lv[d - 1] if merge_edges else [c for l_ in lv for c in l_]):
                        a = par[c][0]
                        for l in range(nl):
                            batch.append((ups[k][l][c][0], dms[l][(c, a)][0])); own.append((1, k, l, c, a, None))
            naln += len(batch)
            for (kind, k, l, c, a, s), (seq, c2) in zip(own, backend.median(batch)):
                if kind == 0:
                    ups[k][l][c] = (seq, c2 + ups[k][l][a][1] + dms[l][(s, a)][1])
                else:
                    ems[(k, l, a, c)] = (seq, c2 + ups[k][l][c][1] + dms[l][(c, a)][1])
        cand, meta = [], []
        for k, ((u, v), js) in enumerate(zip(part, joins)):
            for ji, (a, c) in enumerate(js):
                base = 0
                for l in range(nl):
                    em = ems[(k, l, a, c)]
                    cand.append((dms[l][(v, u)][0], em[0])); base += em[1] + dms[l][(v, u)][1]
                meta.append(((u, v), (a, c), base, (c0 + k, ji)))
        naln += len(cand)
        d = np.asarray(backend.distance(cand), np.int64).reshape(len(meta), nl).sum(axis=1)
        est = d + np.array([m[2] for m in meta], np.int64)
        if len(est):
            q = int(np.argmin(est))
            if best is None or int(est[q]) < best[0]:
                best = (int(est[q]), meta[q][0], meta[q][1], ncand + q, meta[q][3])
        ncand += len(meta)
        if mark is not None:
            backend.release(mark)
    if where is not None:
        where.append(None if best is None else best[4])
    if best is None:
        return None, None, 0, naln
    return best[0], (best[1], best[2]), ncand, naln


def _spr_round_ids(tree, loci, backend, dms, prunings, chunk, where, merge_edges=False):
    """`spr_round` for a backend whose sequences are ids into a device-resident node store: the same batches (one median
    batch per level of the chunk, one edge-median batch, one distance batch), assembled with numpy index arrays.
    Per pruning the rest tree is numbered locally (x1, x2, then BFS levels); `up_*[l, node]` hold id / length /
    accumulated cost of the median that looks away from the cut, `dm_*[l, edge]` those of the unbroken tree."""
    nl = len(loci)
    edges = list(dms[0].keys())
    eidx = {e: i for i, e in enumerate(edges)}
    dm_id = np.array([[dm[e][0].id for e in edges] for dm in dms], np.int32)
    dm_len = np.array([[dm[e][0].n for e in edges] for dm in dms], np.int32)
    dm_cost = np.array([[dm[e][1] for e in edges] for dm in dms], np.int64)
    best, ncand, naln = None, 0, 0
    chunk = _even_chunk(len(prunings), chunk)
    for c0 in range(0, len(prunings), chunk):
        part = prunings[c0:c0 + chunk]
        mark = backend.mark()
        plans = []
        for (u, v) in part:
            x1, x2, lv, par = _side_plan(tree, u, v)
            loc = {x1: 0, x2: 1}
            for lvl in lv:
                for c in lvl:
                    loc[c] = len(loc)
            nn = len(loc)
            up_id = np.zeros((nl, nn), np.int32); up_len = np.zeros((nl, nn), np.int32); up_cost = np.zeros((nl, nn), np.int64)
            for q, e in ((0, eidx[(x2, u)]), (1, eidx[(x1, u)])):
                up_id[:, q] = dm_id[:, e]; up_len[:, q] = dm_len[:, e]; up_cost[:, q] = dm_cost[:, e]
            levels = [(np.array([loc[c] for c in lvl], np.int64), np.array([loc[par[c][0]] for c in lvl], np.int64),
                       np.array([eidx[(par[c][1], par[c][0])] for c in lvl], np.int64)) for lvl in lv]
            joins = [(par[c][0], c) for lvl in lv for c in lvl]
            jc = np.array([loc[c] for _, c in joins], np.int64)
            je = np.array([eidx[(c, a)] for a, c in joins], np.int64)
            plans.append(dict(up_id=up_id, up_len=up_len, up_cost=up_cost, levels=levels, joins=joins, jc=jc, je=je, ev=eidx[(v, u)]))
        depth = max(len(p["levels"]) for p in plans)
        for p in plans:
            nj = len(p["jc"])
            p["em_id"] = np.zeros((nl, nj), np.int32); p["em_len"] = np.zeros((nl, nj), np.int32); p["em_cost"] = np.zeros((nl, nj), np.int64)
            p["jstart"] = np.concatenate([[0], np.cumsum([len(l_[0]) for l_ in p["levels"]])]).astype(np.int64)   # joins are in level order
        for d in range(depth + 1):
            # level d of the dependent medians; with merge_edges also the edge medians of the join edges below level d-1
            # (filler for the latency-bound chain), otherwise all edge medians in one last batch
            act = [p for p in plans if d < len(p["levels"])]
            if merge_edges:
                eact = [(p, int(p["jstart"][d - 1]), int(p["jstart"][d])) for p in plans if 0 < d <= len(p["levels"])]
            else:
                eact = [(p, 0, len(p["jc"])) for p in plans] if d == depth else []
            parts_a = [p["up_id"][:, p["levels"][d][1]].ravel() for p in act] + [p["up_id"][:, p["jc"][j0:j1]].ravel() for p, j0, j1 in eact]
            parts_b = [dm_id[:, p["levels"][d][2]].ravel() for p in act] + [dm_id[:, p["je"][j0:j1]].ravel() for p, j0, j1 in eact]
            if not parts_a:
                continue
            a = np.concatenate(parts_a); b = np.concatenate(parts_b)
            if len(a) == 0:
                continue
            ids, ln, c2 = backend.median_ids(a, b)
            naln += len(a)
            at = 0
            for p in act:
                ci, pi, si = p["levels"][d]
                m = nl * len(ci)
                p["up_id"][:, ci] = ids[at:at + m].reshape(nl, -1); p["up_len"][:, ci] = ln[at:at + m].reshape(nl, -1)
                p["up_cost"][:, ci] = c2[at:at + m].reshape(nl, -1).astype(np.int64) + p["up_cost"][:, pi] + dm_cost[:, si]
                at += m
            for p, j0, j1 in eact:
                m = nl * (j1 - j0)
                p["em_id"][:, j0:j1] = ids[at:at + m].reshape(nl, -1); p["em_len"][:, j0:j1] = ln[at:at + m].reshape(nl, -1)
                p["em_cost"][:, j0:j1] = (c2[at:at + m].reshape(nl, -1).astype(np.int64) + p["up_cost"][:, p["jc"][j0:j1]]
                                          + dm_cost[:, p["je"][j0:j1]])
                at += m
        ca = np.concatenate([np.repeat(dm_id[:, p["ev"]][:, None], len(p["jc"]), axis=1).ravel() for p in plans])
        la = np.concatenate([np.repeat(dm_len[:, p["ev"]][:, None], len(p["jc"]), axis=1).ravel() for p in plans])
        cb = np.concatenate([p["em_id"].ravel() for p in plans]); lb = np.concatenate([p["em_len"].ravel() for p in plans])
        dist = np.asarray(backend.distance_ids(ca, cb, la, lb), np.int64)
        naln += len(ca)
        at = 0
        for k, (p, (u, v)) in enumerate(zip(plans, part)):
            nj = len(p["jc"])
            m = nl * nj
            est = dist[at:at + m].reshape(nl, nj).sum(axis=0) + p["em_cost"].sum(axis=0) + int(dm_cost[:, p["ev"]].sum())
            at += m
            if nj:
                q = int(np.argmin(est))                    # first minimum: ties resolve to the first join edge
                if best is None or int(est[q]) < best[0]:
                    best = (int(est[q]), (u, v), p["joins"][q], ncand + q, (c0 + k, q))
            ncand += nj
        backend.release(mark)
    if where is not None:
        where.append(None if best is None else best[4])
    if best is None:
        return None, None, 0, naln
    return best[0], (best[1], best[2]), ncand, naln


def spr_round_sharded(tree, loci, backend, dms, prunings, chunk=64, rank=0, world=1, device=None, lanes=None, merge_edges=False,
                      timing=None):
    """One SPR neighbourhood strong-scaled over `world` ranks (the Parmap / MPI seam of src/ptree.ml:1356-1408,
    src/allDirChar.ml:2132-2177): the PRUNINGS are dealt to the ranks by estimated work (LPT on the size of the rest
    tree), so every candidate of a pruning -- all its join edges, all loci -- is evaluated on one GPU and the
    incremental medians it needs are computed where they are used; node sequences and the all-direction medians of the
    unbroken tree are replicated.  The only exchange is the reduction of the best candidate: one MIN all-reduce of the
    estimate, one of the (pruning, join edge) ordinal among the ranks that hold it (ties resolve to the first candidate
    in the global enumeration order, whatever the rank count), plus one SUM for the counters.

    `lanes`: a list of (backend, loci, dms) triples, one per concurrent host thread of THIS rank (each with its own
    context / stream / node store holding a replica of the loci and of the all-direction medians).  A rank's prunings
    are dealt to its lanes the same way they are dealt to ranks: the dependent median levels of one pruning are latency
    bound (a level lasts as long as the widest pair's wavefronts), so several independent chains in flight are what
    keeps the GPU busy -- the role of the reference's Parmap worker processes.
    Returns (best estimate, (pruning, join edge), candidates, alignments) -- identical on every rank and for every N."""
    from . import shard
    prunings = list(prunings)
    if lanes is None:
        lanes = [(backend, loci, dms)]
    nlanes = len(lanes)
    if world == 1 and nlanes == 1:
        return spr_round(tree, loci, backend, dms=dms, prunings=prunings, chunk=chunk, merge_edges=merge_edges)
    work = [len(tree.component(u, v)) for (u, v) in prunings]
    parts = shard.lpt_partition(work, world * nlanes)
    big = (1 << 62)

    def run_lane(q):
        b, lc, dm = lanes[q]
        mine = parts[rank * nlanes + q]
        where = []
        est, move, ncand, naln = spr_round(tree, lc, b, dms=dm, prunings=[prunings[i] for i in mine], chunk=chunk, where=where,
                                            merge_edges=merge_edges)
        if est is None:
            return big, big, 0, naln
        pi, ji = where[0]
        return int(est), (int(mine[pi]) << 24) | ji, ncand, naln        # join edges per pruning < 2^24

    import time as _time
    _t0 = _time.perf_counter()
    if nlanes == 1:
        res = [run_lane(0)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nlanes) as ex:      # the ctypes calls release the GIL while the GPU works
            res = list(ex.map(run_lane, range(nlanes)))
    if timing is not None:      # this rank's own share, before it meets the others in the reduction
        timing["local_seconds"] = _time.perf_counter() - _t0
        timing["local_prunings"] = int(sum(len(parts[rank * nlanes + q]) for q in range(nlanes)))
    est, mine_ord = min((r[0], r[1]) for r in res)
    ncand, naln = sum(r[2] for r in res), sum(r[3] for r in res)
    gmin, gord = est, mine_ord
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([est], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gmin = int(t.item())
        o = torch.tensor([mine_ord if est == gmin else big], dtype=torch.int64, device=device)
        dist.all_reduce(o, op=dist.ReduceOp.MIN)
        gord = int(o.item())
        cnt = torch.tensor([ncand, naln], dtype=torch.int64, device=device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ncand, naln = int(cnt[0].item()), int(cnt[1].item())
    if gmin == big:
        return None, None, 0, 0
    # the winner's move is recomputed from the ordinal on every rank (host-only tree walk, no communication)
    gp, gj = gord >> 24, gord & ((1 << 24) - 1)
    u, v = prunings[gp]
    _, _, lv, par = _side_plan(tree, u, v)
    joins = [(par[c][0], c) for l_ in lv for c in l_]
    return gmin, ((u, v), joins[gj]), ncand, naln


def apply_spr(tree, move):
    """prune v's clade at (u, v), suppress u, split the join edge with u and hang the clade back on it"""
    (u, v), (a, c) = move
    t = tree.copy()
    t.remove_edge(u, v)
    x1, x2 = t.adj[u]
    t.remove_edge(u, x1); t.remove_edge(u, x2)
    t.add_edge(x1, x2)
    t.remove_edge(a, c)
    t.add_edge(a, u); t.add_edge(u, c); t.add_edge(u, v)
    return t


# ------------------------------------------------------------------------------------------------------------
# TBR round over several loci with incremental medians (BASELINE config #4)
# ------------------------------------------------------------------------------------------------------------

def _side_plan(tree, s, t):
    """The component on s's side of the cut edge s-t with s suppressed (its neighbours x1, x2 joined), as BFS levels
    from the cut: (x1, x2, levels of nodes, node -> (parent, sibling)); None when s is a leaf."""
    nb = [w for w in tree.adj[s] if w != t]
    if not nb:
        return None
    x1, x2 = nb
    lv, frontier, par = [], [(x1, s), (x2, s)], {}
    while frontier:
        nxt = []
        for a, pa in frontier:
            ch = [w for w in tree.adj[a] if w != pa]
            if ch:
                nxt += [(ch[0], a), (ch[1], a)]
                par[ch[0]] = (a, ch[1]); par[ch[1]] = (a, ch[0])
        frontier = nxt
        if nxt:
            lv.append([c for c, _ in nxt])
    return x1, x2, lv, par


def tbr_round_multi(tree, loci, backend, dms=None, breaks=None, chunk=16):
    """One TBR neighbourhood over several loci (src/ptree.ml:1356-1453 with reroots): for the break u-v BOTH
    components are re-rooted, i.e. every edge of u's side (u suppressed) is paired with every edge of v's side
    (v suppressed); the edge medians of either side are obtained incrementally exactly like in `spr_round`
    (the merged edge x1-x2 reuses two medians of the unbroken tree), and all (edge, edge, locus) triples of a chunk
    of breaks go into ONE cost-only batch.  The pair (merged edge, merged edge) is the unbroken tree and is left out.
    Edges are named 'm' (merged), None (the lone leaf of a one-leaf side) or (parent, child).
    Returns (best estimate, (break, edge on u's side, edge on v's side), candidates, alignments issued)."""
    nl = len(loci)
    dms = all_directions(tree, loci, backend) if dms is None else dms
    breaks = tree.edges() if breaks is None else list(breaks)
    best, ncand, naln = None, 0, 0
    for c0 in range(0, len(breaks), chunk):
        part = breaks[c0:c0 + chunk]
        mark = backend.mark() if hasattr(backend, "mark") else None     # node store: the chunk's medians are temporaries
        sides = [(s, t, _side_plan(tree, s, t)) for (u, v) in part for (s, t) in ((u, v), (v, u))]
        ups = [None if pl is None else [{pl[0]: dm[(pl[1], s)], pl[1]: dm[(pl[0], s)]} for dm in dms] for s, t, pl in sides]
        depth = max([len(pl[2]) for _, _, pl in sides if pl is not None], default=0)
        for d in range(depth):
            batch, own = [], []
            for k, (s, t, pl) in enumerate(sides):
                if pl is not None and d < len(pl[2]):
                    for c in pl[2][d]:
                        a, sib = pl[3][c]
                        for l in range(nl):
                            batch.append((ups[k][l][a][0], dms[l][(sib, a)][0])); own.append((k, l, c, a, sib))
            naln += len(batch)
            for (k, l, c, a, sib), (seq, c2) in zip(own, backend.median(batch)):
                ups[k][l][c] = (seq, c2 + ups[k][l][a][1] + dms[l][(sib, a)][1])
        # edge medians of every side: the merged edge, then the edges below it
        batch, own = [], []
        for k, (s, t, pl) in enumerate(sides):
            if pl is None:
                continue
            x1, x2, lv, par = pl
            for l in range(nl):
                batch.append((dms[l][(x1, s)][0], dms[l][(x2, s)][0])); own.append((k, l, "m", dms[l][(x1, s)][1] + dms[l][(x2, s)][1]))
            for lvl in lv:
                for c in lvl:
                    a = par[c][0]
                    for l in range(nl):
                        batch.append((ups[k][l][c][0], dms[l][(c, a)][0]))
                        own.append((k, l, (a, c), ups[k][l][c][1] + dms[l][(c, a)][1]))
        naln += len(batch)
        ems = [dict() for _ in sides]
        for (k, l, e, below), (seq, c2) in zip(own, backend.median(batch)):
            ems[k].setdefault(e, [None] * nl)[l] = (seq, c2 + below)
        for k, (s, t, pl) in enumerate(sides):
            if pl is None:
                ems[k][None] = [(ls[s], 0) for ls in loci]
        cand, meta = [], []
        for bi, brk in enumerate(part):
            ea_all, eb_all = ems[2 * bi], ems[2 * bi + 1]
            root_a = None if sides[2 * bi][2] is None else "m"
            root_b = None if sides[2 * bi + 1][2] is None else "m"
            for ea, ma in ea_all.items():
                for eb, mb in eb_all.items():
                    if ea == root_a and eb == root_b:
                        continue
                    base = 0
                    for l in range(nl):
                        cand.append((ma[l][0], mb[l][0])); base += ma[l][1] + mb[l][1]
                    meta.append((brk, ea, eb, base))
        naln += len(cand)
        if meta:
            d = np.asarray(backend.distance(cand), np.int64).reshape(len(meta), nl).sum(axis=1)
            est = d + np.array([m[3] for m in meta], np.int64)
            q = int(np.argmin(est))
            if best is None or int(est[q]) < best[0]:
                best = (int(est[q]), meta[q][:3])
        ncand += len(meta)
        if mark is not None:
            backend.release(mark)
    if best is None:
        return None, None, 0, naln
    return best[0], best[1], ncand, naln


def apply_tbr_multi(tree, move):
    """cut u-v, suppress the internal ones of u and v, split the chosen edge of either side (re-using the ids u, v)
    and join the two attachment points"""
    (u, v), ea, eb = move
    t = tree.copy()
    t.remove_edge(u, v)

    def attach(s, e):
        if e is None:                      # one-leaf side: the leaf itself
            return s
        x1, x2 = t.adj[s]
        t.remove_edge(s, x1); t.remove_edge(s, x2)
        a, c = (x1, x2) if e == "m" else e
        if e != "m":
            t.add_edge(x1, x2)
            t.remove_edge(a, c)
        t.add_edge(a, s); t.add_edge(s, c)
        return s
    pa, pb = attach(u, ea), attach(v, eb)
    t.add_edge(pa, pb)
    return t


# ------------------------------------------------------------------------------------------------------------
# Single assignment: the pre-order pass that follows a downpass
# ------------------------------------------------------------------------------------------------------------

def single_assignment(tree, loci, backend, root=None):
    """Pre-order single-assignment pass: the root median is resolved against itself, every other node's towards-root
    median against its parent's single assignment (SeqCS.DOS.to_single = Sequence.Align.closest parent mine,
    src/seqCS.ml:950-982, src/sequence.ml:1180-1237), level by level from the root, all loci in the same batches.
    Needs a backend with ``single(pairs of (parent, mine)) -> [(sequence, cost), ...]``.
    Returns (downpass cost, single-assignment cost = sum of the re-costed parent/child distances, per-locus dict
    node -> single sequence; the root median is stored under the key 'root')."""
    root = tree.edges()[0] if root is None else root
    cost, dms = downpass(tree, loci, backend, root)
    a, b = root
    res = backend.median([(dm[(a, b)][0], dm[(b, a)][0]) for dm in dms])
    singles = [dict() for _ in loci]
    total = 0
    for l, ((seq, _), (s1, c1)) in enumerate(zip(res, backend.single([(seq, seq) for seq, _ in res]))):
        singles[l]["root"] = s1; total += c1
    up = {a: b, b: a}                       # neighbour towards the root edge
    above = {a: "root", b: "root"}          # whose single assignment a node is resolved against
    level = [a, b]
    while level:
        out = backend.single([(singles[l][above[x]], dms[l][(x, up[x])][0]) for l in range(len(loci)) for x in level])
        for l in range(len(loci)):
            for q, x in enumerate(level):
                s1, c1 = out[l * len(level) + q]
                singles[l][x] = s1; total += c1
        nxt = []
        for x in level:
            for y in tree.adj[x]:
                if y != up[x]:
                    up[y] = x; above[y] = x; nxt.append(y)
        level = nxt
    return cost, int(total), singles
