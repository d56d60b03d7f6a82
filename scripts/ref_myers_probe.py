"""Evidence for DESIGN.md section 7: the reference's Myers O(ND) distance (algn_myers, src/algn.c:3697-3743,
`Sequence.Align.myers`, no caller in src/*.ml) does not compute an edit distance.  Run in the build container
(needs oracle/_ref).  zarr_test_pos (src/zarr.c:31-39) accepts a negative index only if it is >= the (positive)
length, i.e. never, so every read of V[k] with k < 0 leaves the caller's variable untouched and every write is
dropped: the lower half of Myers' V array does not exist.  In addition zarr_clear (src/zarr.c:94-102) zeroes
arr[0 .. 2*max+1] counted from the START of the buffer while reads and writes are centred at arr[length + k], so
once the static scratch has been sized by a longer pair a shorter pair runs on the previous call's values: the
result depends on the call history of the process."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.refbind import RefLib

R = RefLib(); L = R.lib
L.ref_myers.restype = C.c_int
p = lambda a: a.ctypes.data_as(C.POINTER(C.c_ubyte))

def ref(a, b):
    a = np.array(a, np.uint8); b = np.array(b, np.uint8)
    return L.ref_myers(p(a), len(a), p(b), len(b))

def indel_distance(a, b):          # len a + len b - 2 LCS over the bases (index 0 is the leading gap)
    a, b = a[1:], b[1:]
    D = np.zeros((len(a) + 1, len(b) + 1), int)
    for i in range(1, len(a) + 1):
        for j in range(1, len(b) + 1):
            D[i, j] = D[i - 1, j - 1] + 1 if a[i - 1] == b[j - 1] else max(D[i - 1, j], D[i, j - 1])
    return len(a) + len(b) - 2 * int(D[-1, -1])

# fresh process: the scratch is sized by the first pair
for a, b in [([16, 1], [16, 2]), ([16, 1, 2, 4, 8], [16, 8, 4, 2, 1]), ([16, 1, 1, 1, 1, 1, 1], [16, 2, 2, 2, 2, 2, 2, 2, 2])]:
    print("a=%s b=%s reference %d, insertion/deletion distance %d" % (a[1:], b[1:], ref(a, b), indel_distance(a, b)))
probe = ([16, 8], [16, 2, 4, 4, 2, 2, 4])
before = ref(*probe)
rng = np.random.default_rng(0)
big = [16] + [int(x) for x in rng.choice([1, 2, 4, 8], 60)]
ref(big, big)                       # a longer pair re-sizes the static scratch
print("the same pair before / after a longer pair was aligned: %d / %d (distance %d)" % (before, ref(*probe), indel_distance(*probe)))
bad = 0
for t in range(200):
    a = [16] + [int(x) for x in rng.choice([1, 2, 4, 8], int(rng.integers(1, 12)))]
    b = [16] + [int(x) for x in rng.choice([1, 2, 4, 8], int(rng.integers(1, 12)))]
    bad += ref(a, b) != indel_distance(a, b)
print("%d of 200 random short pairs differ from the insertion/deletion distance" % bad)
