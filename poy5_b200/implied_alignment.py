"""Implied alignment of dynamic-homology sequence characters: the host-side bookkeeping of
src/impliedAlignment.ml around ONE batch of pairwise alignments on the GPU.

Reference path (state `Seq, `calculate_median = false`, src/impliedAlignment.ml:1595-1825): starting at one end of the root
edge, every vertex gets an `ias` (its single-assignment sequence -- observed sequence for a leaf -- with a fresh code
per position, `create_ias` :169-200) and the subtrees hanging off it are joined into it one after the other
(`tree_traverser` / `join_2_nodes` -> `ancestor ~calc_m:false` :355-557): the vertex' sequence is aligned against the
child's (`Sequence.Align.align_2`, the banded alignment + traceback hot path) and the two homology tables are merged
column by column; `invert_codes` / `convert_a_taxon` (:1828-1905) finally turn the root's ordered codes into one gapped
row per taxon.

What runs where: every alignment of the traversal is between the two ORIGINAL sequences of a tree edge unless an earlier
join dropped a column of the vertex (possible only for symbols that carry the gap bit), so all (2n-3) x loci edge
alignments go to the device in one batch (`align2`), and the merge -- O(total length), pointer chasing over hash
tables in the reference -- stays on the host.  A join whose operands differ from the speculated ones is re-aligned
on its own.  OCaml cannot run in this image: the merge below is a restatement, pinned by the properties every implied
alignment must have (tests/test_implied_alignment.py) and by equality between the GPU-driven and the checker-driven
composition, not by reference output.  Neighbour order: the reference visits `Tree.Interior (_, a, b, c)` in stored
order; here neighbours are visited in increasing id."""
import numpy as np

GAP = 16


class Ias:
    """`ias` of src/impliedAlignment.ml:48-65 restricted to what `Seq needs: seq, codes (pos -> code), homologous
    (code -> list of codes), order (codes, LAST column first)."""
    __slots__ = ("seq", "codes", "hom", "order")

    def __init__(self, seq, codes, hom, order):
        self.seq, self.codes, self.hom, self.order = seq, codes, hom, order


def create_ias(seq, cg):
    """create_ias (src/impliedAlignment.ml:169-200): position 0 (the leading gap) gets no code"""
    seq = np.asarray(seq, np.uint8)
    codes, hom, order = {}, {}, []
    for pos in range(1, len(seq)):
        c = cg()
        codes[pos] = c; hom[c] = [c]; order.append(c)
    order.reverse()
    return Ias(seq, codes, hom, order)


def ancestor(a, b, rows, cost):
    """`ancestor ~calc_m:false` (src/impliedAlignment.ml:355-557): `a` is the ancestor of `b`; `rows` = the two aligned
    rows of (a.seq, b.seq); `cost` = the 32x32 Two_D cost table.  Returns the merged ias (a's and b's tables are
    consumed)."""
    ra, rb = rows
    lena, lenb = len(a.seq), len(b.seq)
    a_pos, b_pos = lena - 1, lenb - 1
    anc, anc_pos, codes, hom = [], 0, {}, {}
    a_hom, b_hom = a.hom, b.hom
    a_or, b_or, ia, ib = a.order, b.order, 0, 0
    acc = []                                   # the reference's res_or, reversed

    def flush(src, i, it):                     # prepend_until_shared
        while src[i] != it:
            acc.append(src[i]); i += 1
        return i + 1

    for position in range(len(ra) - 1, -1, -1):
        it_a, it_b = int(ra[position]), int(rb[position])
        med = it_a                             # nogap = `B: whatever is assigned to the true ancestor
        if it_a != GAP and it_b != GAP:
            is_gap_median = int(cost[it_a, it_b]) < int(cost[it_a & 15, it_b & 15]) or med == GAP
        else:
            is_gap_median = med == GAP
        code = -1
        if it_a != GAP and it_b != GAP:
            codea, codeb = a.codes[a_pos], b.codes[b_pos]
            hom_a, hom_b = a_hom.pop(codea), b_hom.pop(codeb)
            if not is_gap_median:
                ia = flush(a_or, ia, codea); ib = flush(b_or, ib, codeb)
                hom[codea] = hom_a + hom_b
                acc.append(codea)
            else:
                ib = flush(b_or, ib, codeb); ia = flush(a_or, ia, codea)
                hom[codeb] = hom_b; hom[codea] = hom_a
                acc.append(codeb); acc.append(codea)
            code = codea; a_pos -= 1; b_pos -= 1
        elif it_a == GAP and it_b != GAP:
            codeb = b.codes[b_pos]
            hom[codeb] = b_hom.pop(codeb)
            ib = flush(b_or, ib, codeb)
            acc.append(codeb); code = codeb; b_pos -= 1
        elif it_a != GAP:
            codea = a.codes[a_pos]
            hom[codea] = a_hom.pop(codea)
            ia = flush(a_or, ia, codea)
            acc.append(codea); code = codea; a_pos -= 1
        if not is_gap_median:
            anc.append(med); codes[anc_pos] = code; anc_pos += 1
    seq = np.array([GAP] + anc[::-1], np.uint8)
    codes = {anc_pos - k: c for k, c in codes.items()}
    a_hom.update(hom); a_hom.update(b_hom)
    acc.extend(a_or[ia:]); acc.extend(b_or[ib:])
    return Ias(seq, codes, a_hom, acc)


def _rows_for(a_seq, b_seq, align2_one):
    """the operands of `ancestor` (:376-411): empty sequences are not aligned"""
    aempty, bempty = bool((a_seq == GAP).all()), bool((b_seq == GAP).all())
    if aempty and bempty:
        s = a_seq if len(a_seq) > len(b_seq) else b_seq
        return s, s
    if aempty:
        return np.full(len(b_seq), GAP, np.uint8), b_seq
    if bempty:
        return a_seq, np.full(len(a_seq), GAP, np.uint8)
    return align2_one(a_seq, b_seq)


def implied_alignment(tree, root, seqs, cost, align2):
    """One character (locus).  tree: treesearch.Tree; root = (self, other), the root edge; seqs: node -> uint8 sequence
    (leading gap included; leaves: observed, interior vertices: single assignment); cost: 32x32 Two_D cost table;
    align2(list of (a, b)) -> list of (row_a, row_b) = Sequence.Align.align_2, batched.
    Returns (matrix uint8[n_leaves, columns + 1] with gap = 16 in column 0 and in unfilled cells, leaf ids in row
    order, number of alignments re-done because a join had changed an operand)."""
    self_, other = root
    # the traversal as a list of joins (parent vertex, child vertex), children in the order they are joined
    plan, stack_children = [], {}

    def visit(parent, me):
        kids = sorted(y for y in tree.adj[me] if y != parent)
        stack_children[me] = kids
        for k in kids:
            visit(me, k)
            plan.append((me, k))
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * len(tree.adj) + 100))
    kids0 = sorted(tree.adj[self_])
    for k in kids0:
        visit(self_, k)
        plan.append((self_, k))
    seqs = {x: np.asarray(s, np.uint8) for x, s in seqs.items()}
    spec = {}
    need = [(p, c) for (p, c) in plan if not ((seqs[p] == GAP).all() or (seqs[c] == GAP).all())]
    for (p, c), r in zip(need, align2([(seqs[p], seqs[c]) for p, c in need])):
        spec[(p, c)] = r
    counter = [0]

    def cg():
        counter[0] += 1
        return counter[0]
    ias, leaf_ias = {}, {}
    # codes are handed out when the traversal first reaches a vertex (convert_node), i.e. in pre-order
    pre = [self_]

    def pre_walk(me):
        pre.append(me)
        for k in stack_children.get(me, []):
            pre_walk(k)
    for k in kids0:
        pre_walk(k)
    for x in pre:
        ias[x] = create_ias(seqs[x], cg)
        if len(tree.adj[x]) == 1:
            leaf_ias[x] = Ias(ias[x].seq, dict(ias[x].codes), None, None)
    redone = 0
    for (p, c) in plan:
        a, b = ias[p], ias[c]
        r = spec.get((p, c))
        same = len(a.seq) == len(seqs[p]) and len(b.seq) == len(seqs[c]) and bool((a.seq == seqs[p]).all()) and bool((b.seq == seqs[c]).all())
        if r is None or not same:
            if r is not None:
                redone += 1
            r = _rows_for(a.seq, b.seq, lambda x, y: align2([(x, y)])[0])
        ias[p] = ancestor(a, b, r, cost)
        del ias[c]
    fin = ias[self_]
    # invert_codes + convert_a_taxon (src/impliedAlignment.ml:1828-1905)
    n = len(fin.order)
    remap, recode = {}, {}
    for ccol, code in enumerate(fin.order):
        remap[code] = ccol
        for hc in fin.hom[code]:
            recode[hc] = code
    leaves = sorted(leaf_ias)
    out = np.zeros((len(leaves), n + 1), np.uint8)
    for row, x in enumerate(leaves):
        li = leaf_ias[x]
        for pos, code in li.codes.items():
            out[row, n - remap[recode[code]]] = li.seq[pos]
    out[out == 0] = GAP
    return out, leaves, redone


# ---- output: the symbols of Alphabet.nucleotides (src/alphabet.ml:270-315; the FIRST name bound to a code is the one
# printed: T not U, N not X, * not ?) and a fasta writer for the matrix returned by implied_alignment --------------------
NUCLEOTIDE_SYMBOLS = {1: "A", 2: "C", 4: "G", 8: "T", 3: "M", 5: "R", 9: "W", 6: "S", 10: "Y", 12: "K", 7: "V", 11: "H", 13: "D",
                      14: "B", 15: "N", 16: "-", 17: "1", 18: "2", 19: "3", 20: "4", 21: "5", 22: "6", 23: "7", 24: "8", 25: "9",
                      26: "0", 27: "!", 28: "^", 29: "$", 30: "#", 31: "*"}
_SYM = np.array([ord(NUCLEOTIDE_SYMBOLS.get(c, "?")) for c in range(32)], np.uint8)


def to_strings(matrix):
    """rows of an implied-alignment matrix as strings; column 0 (the leading gap every sequence carries) is dropped"""
    m = np.asarray(matrix, np.uint8)
    return [_SYM[row[1:] & 31].tobytes().decode("ascii") for row in m]


def write_fasta(fh, matrix, names, width=0):
    """one record per taxon, in row order (report (implied_alignments) of the reference writes the same content through
    its fasta formatter; the line layout here is not pinned to it).  width > 0 wraps the sequence lines."""
    for name, row in zip(names, to_strings(matrix)):
        fh.write(">%s\n" % name)
        if width and width > 0:
            for k in range(0, len(row), width):
                fh.write(row[k:k + width] + "\n")
        else:
            fh.write(row + "\n")
