"""Evidence for DESIGN.md section 7: the reference's 3-D cube (algn_fill_cube, src/algn.c:2593-2763) is dead code
AND wrong.  Run in the build container (needs oracle/_ref).  Three identical sequences must align at cost 0 under
any sum-of-pairs cost; the reference returns 6, and about twice the optimum on random triples.  Cause: diag_m /
upper_m / prev_m are advanced s2_len-1 rows per plane while mm advances s2_len rows (src/algn.c:2686-2733), so
from the first plane on the three predecessor rows lag behind the row being filled."""
import ctypes as C, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.refbind import RefLib

R = RefLib(); L = R.lib
cm3 = L.ref_cm3d_new(5, 31)
codes = [1, 2, 4, 8, 16]
c3 = lambda a, b, c: (a != b) + (a != c) + (b != c)
for a in range(1, 32):
    for b in range(1, 32):
        for c in range(1, 32):
            best = min(c3(x, y, z) for x in codes if a & x for y in codes if b & y for z in codes if c & z)
            L.ref_cm3d_set(cm3, a, b, c, best, a | b | c)

def ref3(s1, s2, s3):
    s1, s2, s3 = (np.array(s, np.uint8) for s in (s1, s2, s3))
    cap = len(s1) + len(s2) + len(s3) + 3
    o = [np.zeros(cap, np.uint8) for _ in range(3)]; lens = (C.c_int * 3)()
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_ubyte))
    return L.ref_align_3d(cm3, R.mat, p(s1), len(s1), p(s2), len(s2), p(s3), len(s3), p(o[0]), p(o[1]), p(o[2]), lens)

def dp3(s1, s2, s3):
    D = np.full((len(s1), len(s2), len(s3)), 10**9); D[0, 0, 0] = 0
    for i, j, k in itertools.product(range(len(s1)), range(len(s2)), range(len(s3))):
        if i == j == k == 0: continue
        for di, dj, dk in itertools.product((0, 1), repeat=3):
            if (di, dj, dk) == (0, 0, 0) or i - di < 0 or j - dj < 0 or k - dk < 0: continue
            D[i, j, k] = min(D[i, j, k], D[i - di, j - dj, k - dk] + c3(s1[i] if di else 16, s2[j] if dj else 16, s3[k] if dk else 16))
    return int(D[-1, -1, -1])

s = [16, 1, 2, 4, 8, 1]
print("identical triple: reference", ref3(s, s, s), "optimum", dp3(s, s, s))
rng = np.random.default_rng(0)
for t in range(5):
    x = [[16] + list(rng.choice([1, 2, 4, 8], int(rng.integers(3, 8)))) for _ in range(3)]
    print("random triple: reference", ref3(*x), "optimum", dp3(*x))
