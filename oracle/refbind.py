"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libpoyref[_long].so,
the UNMODIFIED reference C (src/algn.c etc.) behind oracle/ref_driver.c."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_FAIL = -2**31
_u8p = C.POINTER(C.c_ubyte)
_i32p = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_longlong)


def _p8(a):
    return a.ctypes.data_as(_u8p)


def available(long_sequences=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libpoyref_long.so" if long_sequences else "libpoyref.so"))


class RefLib:
    def __init__(self, long_sequences=False):
        name = "libpoyref_long.so" if long_sequences else "libpoyref.so"
        self.lib = L = C.CDLL(os.path.join(_HERE, "_ref", name))
        L.ref_mat_new.restype = C.c_void_p
        L.ref_mat_free.argtypes = [C.c_void_p]
        L.ref_mat_reserve.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_cm_new.restype = C.c_void_p
        L.ref_cm_new.argtypes = [C.c_int] * 4
        L.ref_cm_free.argtypes = [C.c_void_p]
        L.ref_cm_load.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _u8p, _i32p, _i32p]
        L.ref_cm_min_non0.argtypes = [C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_cost_affine.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int]
        L.ref_align_affine.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int,
                                       _u8p, _u8p, _u8p, _u8p, _i32p]
        L.ref_cost_linear.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int]
        L.ref_align_linear.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int,
                                       _u8p, _u8p, _i32p]
        L.ref_ancestor_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, _u8p]
        L.ref_median_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, _u8p]
        L.ref_union.argtypes = [_u8p, _u8p, C.c_int, _u8p]
        L.ref_worst_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int]
        L.ref_verify_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int]
        L.ref_cm3d_new.restype = C.c_void_p
        L.ref_cm3d_new.argtypes = [C.c_int, C.c_int]
        L.ref_cm3d_set.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.ref_align_3d.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int,
                                   _u8p, _u8p, _u8p, _i32p]
        L.ref_batch_affine.restype = C.c_double
        L.ref_batch_affine.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p, _i64p, _i32p, _i64p, _i32p, _u8p,
                                       _i32p, C.c_int]
        if hasattr(L, "ref_newkk_new"):
            L.ref_newkk_new.restype = C.c_void_p
            L.ref_newkk_free.argtypes = [C.c_void_p]
            L.ref_newkk_get_k.argtypes = [C.c_void_p]
            L.ref_newkk_cost.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
            L.ref_newkk_align.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _i32p]
            L.ref_powell_3d.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _u8p, _i32p]
            self.ukkm = L.ref_newkk_new()
        self.mat = L.ref_mat_new()
        L.ref_mat_reserve(self.mat, 64, 64)

    # -- cost matrices ------------------------------------------------------
    def cm(self, m):
        """Build a reference ``struct cm`` from an oracle.cost_matrix_oracle.CostMatrix2D
        (or anything with the same fields)."""
        h = self.lib.ref_cm_new(m.a_sz_letters, m.cost_model_type, m.gap_open, m.all_elements)
        cost = np.ascontiguousarray(m.cost, np.int32)
        worst = np.ascontiguousarray(m.worst, np.int32)
        med = np.ascontiguousarray(m.median, np.uint8)
        pre = np.ascontiguousarray(m.prepend, np.int32)
        tail = np.ascontiguousarray(m.tail, np.int32)
        self.lib.ref_cm_load(h, m.n, cost.ctypes.data_as(_i32p), worst.ctypes.data_as(_i32p), _p8(med),
                             pre.ctypes.data_as(_i32p), tail.ctypes.data_as(_i32p))
        return h

    def _err(self):
        return RuntimeError("reference Failure: " + self.lib.ref_last_error().decode())

    # -- affine ---------------------------------------------------------------
    def cost_affine(self, cm, s1, s2):
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        r = self.lib.ref_cost_affine(cm, self.mat, _p8(s1), len(s1), _p8(s2), len(s2))
        if r == REF_FAIL:
            raise self._err()
        return r

    def align_affine(self, cm, si, sj, swaped=0):
        """-> (cost, median, medianwg, resi, resj); requires len(si) <= len(sj)."""
        si = np.ascontiguousarray(si, np.uint8); sj = np.ascontiguousarray(sj, np.uint8)
        cap = len(si) + len(sj) + 2
        outs = [np.zeros(cap, np.uint8) for _ in range(4)]
        lens = (C.c_int * 4)()
        r = self.lib.ref_align_affine(cm, self.mat, _p8(si), len(si), _p8(sj), len(sj), int(swaped),
                                      _p8(outs[0]), _p8(outs[1]), _p8(outs[2]), _p8(outs[3]), lens)
        if r == REF_FAIL:
            raise self._err()
        return (r,) + tuple(o[:n].copy() for o, n in zip(outs, lens))

    # -- linear ---------------------------------------------------------------
    def cost_linear(self, cm, s1, s2, deltawh):
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        r = self.lib.ref_cost_linear(cm, self.mat, _p8(s1), len(s1), _p8(s2), len(s2), int(deltawh))
        if r == REF_FAIL:
            raise self._err()
        return r

    def align_linear(self, cm, s1, s2, deltawh, swaped=0):
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        cap = len(s1) + len(s2) + 2
        outs = [np.zeros(cap, np.uint8) for _ in range(2)]
        lens = (C.c_int * 2)()
        r = self.lib.ref_align_linear(cm, self.mat, _p8(s1), len(s1), _p8(s2), len(s2), int(deltawh),
                                      int(swaped), _p8(outs[0]), _p8(outs[1]), lens)
        if r == REF_FAIL:
            raise self._err()
        return (r,) + tuple(o[:n].copy() for o, n in zip(outs, lens))

    # -- O(L) helpers -----------------------------------------------------------
    def ancestor_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.lib.ref_ancestor_2(cm, _p8(a), _p8(b), len(a), _p8(out))
        if n == REF_FAIL:
            raise self._err()
        return out[:n].copy()

    def median_2(self, cm, a, b, with_gaps):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.lib.ref_median_2(cm, _p8(a), _p8(b), len(a), int(with_gaps), _p8(out))
        return out[:n].copy()

    def union(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a) + 1, np.uint8)
        n = self.lib.ref_union(_p8(a), _p8(b), len(a), _p8(out))
        return out[:n].copy()

    def worst_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.ref_worst_2(cm, _p8(a), _p8(b), len(a))

    def verify_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.ref_verify_2(cm, _p8(a), _p8(b), len(a))

    # -- Sequence.NewkkAlign (src/newkkonen.c) ---------------------------------------
    def newkk_fresh(self):
        """a new, never used newkkmat scratch (the reference's is process global: NewkkAlign.default_ukkm)"""
        self.lib.ref_newkk_free(self.ukkm)
        self.ukkm = self.lib.ref_newkk_new()

    def newkk_align(self, cm, s1, s2, affine, swaped=0):
        """-> (cost, aligned s1, aligned s2, final k); requires len(s1) <= len(s2)"""
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        cap = len(s1) + len(s2) + 2
        outs = [np.zeros(cap, np.uint8) for _ in range(2)]
        lens = (C.c_int * 2)()
        r = self.lib.ref_newkk_align(cm, self.ukkm, _p8(s1), len(s1), _p8(s2), len(s2), int(affine), int(swaped),
                                     _p8(outs[0]), _p8(outs[1]), lens)
        if r == REF_FAIL:
            raise self._err()
        return (r,) + tuple(o[:n].copy() for o, n in zip(outs, lens)) + (self.lib.ref_newkk_get_k(self.ukkm),)

    def newkk_cost(self, cm, s1, s2, affine, swaped=0):
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        r = self.lib.ref_newkk_cost(cm, self.ukkm, _p8(s1), len(s1), _p8(s2), len(s2), int(affine), int(swaped))
        if r == REF_FAIL:
            raise self._err()
        return r

    # -- powell_3D_align (src/ukkCommon.c:109) ------------------------------------------
    def powell_3d(self, s1, s2, s3, mm, go, ge):
        """inputs WITH the leading gap (copySequence skips element 0); -> (cost, row1, row2, row3)"""
        ss = [np.ascontiguousarray(x, np.uint8) for x in (s1, s2, s3)]
        cap = sum(len(x) for x in ss) + 3
        outs = [np.zeros(cap, np.uint8) for _ in range(3)]
        lens = (C.c_int * 3)()
        r = self.lib.ref_powell_3d(_p8(ss[0]), len(ss[0]), _p8(ss[1]), len(ss[1]), _p8(ss[2]), len(ss[2]), int(mm), int(go), int(ge),
                                   _p8(outs[0]), _p8(outs[1]), _p8(outs[2]), lens)
        if r == REF_FAIL:
            raise self._err()
        return (r,) + tuple(o[:n].copy() for o, n in zip(outs, lens))

    # -- threaded batch (CPU baseline) ---------------------------------------------
    def batch_affine(self, cm, mode, seqs, off_i, len_i, off_j, len_j, swaped=None, nthreads=1):
        """mode 0 = cost only, 1 = align+traceback.  -> (seconds, int32 costs)."""
        seqs = np.ascontiguousarray(seqs, np.uint8)
        off_i = np.ascontiguousarray(off_i, np.int64); off_j = np.ascontiguousarray(off_j, np.int64)
        len_i = np.ascontiguousarray(len_i, np.int32); len_j = np.ascontiguousarray(len_j, np.int32)
        n = len(len_i)
        cost = np.zeros(n, np.int32)
        sw = None if swaped is None else np.ascontiguousarray(swaped, np.uint8)
        t = self.lib.ref_batch_affine(cm, mode, n, _p8(seqs), off_i.ctypes.data_as(_i64p),
                                      len_i.ctypes.data_as(_i32p), off_j.ctypes.data_as(_i64p),
                                      len_j.ctypes.data_as(_i32p), None if sw is None else _p8(sw),
                                      cost.ctypes.data_as(_i32p), int(nthreads))
        return t, cost
