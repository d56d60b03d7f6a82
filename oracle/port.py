"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libdooracle.so (the plain-C
restatement oracle/do_oracle.c)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdooracle.so")
_u8p = C.POINTER(C.c_ubyte)
_i32p = C.POINTER(C.c_int)
INT_MIN = -2**31


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("do_oracle.c", "newkk_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", "-o", _SO] + srcs + ["-lpthread"])
    return _SO


class AlignStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("final_T", C.c_int), ("final_k", C.c_int), ("cells", C.c_longlong)]


def _p8(a):
    return a.ctypes.data_as(_u8p)


class Port:
    def __init__(self):
        self.lib = L = C.CDLL(build())
        L.do_cm_sizeof.restype = C.c_int
        L.do_cm_init.argtypes = [C.c_void_p, _i32p, _i32p, _u8p, _i32p, _i32p, C.c_int, C.c_int]
        L.do_cost_affine.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int]
        L.do_scratch_new.restype = C.c_void_p
        L.do_scratch_free.argtypes = [C.c_void_p]
        L.do_align_affine.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int,
                                      _u8p, _u8p, _u8p, _u8p, _i32p, C.POINTER(AlignStats)]
        L.do_batch_affine.restype = C.c_double
        L.do_batch_affine.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p, C.POINTER(C.c_longlong), _i32p,
                                      C.POINTER(C.c_longlong), _i32p, _u8p, _i32p, C.c_int]
        L.do_lin_scratch_new.restype = C.c_void_p
        L.do_cost_linear.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.POINTER(AlignStats)]
        L.do_backtrace_linear.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, _u8p, _u8p, _i32p]
        L.do_median_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, _u8p]
        L.do_union.argtypes = [_u8p, _u8p, C.c_int, _u8p]
        L.do_worst_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int]
        L.do_verify_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int]
        L.do_ancestor_2.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, _u8p]
        L.do_newkk_align_affine.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, C.c_int, _u8p, _u8p, _i32p, C.POINTER(AlignStats)]
        self.lin = L.do_lin_scratch_new()
        self.scratch = L.do_scratch_new()

    def cm(self, m):
        """do_cm from a CostMatrix2D-like object (32x32 tables)."""
        buf = C.create_string_buffer(self.lib.do_cm_sizeof())
        cost = np.ascontiguousarray(m.cost, np.int32); worst = np.ascontiguousarray(m.worst, np.int32)
        med = np.ascontiguousarray(m.median, np.uint8)
        pre = np.ascontiguousarray(m.prepend, np.int32); tail = np.ascontiguousarray(m.tail, np.int32)
        self.lib.do_cm_init(buf, cost.ctypes.data_as(_i32p), worst.ctypes.data_as(_i32p), _p8(med),
                            pre.ctypes.data_as(_i32p), tail.ctypes.data_as(_i32p), m.gap_open, m.cost_model_type)
        return buf

    def cost_affine(self, cm, s1, s2):
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        return self.lib.do_cost_affine(cm, _p8(s1), len(s1), _p8(s2), len(s2))

    def align_affine(self, cm, si, sj, swaped=0, with_stats=False):
        si = np.ascontiguousarray(si, np.uint8); sj = np.ascontiguousarray(sj, np.uint8)
        cap = len(si) + len(sj) + 2
        outs = [np.zeros(cap, np.uint8) for _ in range(4)]
        lens = (C.c_int * 4)()
        st = AlignStats()
        r = self.lib.do_align_affine(cm, self.scratch, _p8(si), len(si), _p8(sj), len(sj), int(swaped),
                                     _p8(outs[0]), _p8(outs[1]), _p8(outs[2]), _p8(outs[3]), lens, C.byref(st))
        if r == INT_MIN:
            raise RuntimeError("pass the shorter one as first")
        res = (r,) + tuple(o[:n].copy() for o, n in zip(outs, lens))
        return res + (st,) if with_stats else res

    def batch_affine(self, cm, mode, seqs, off_i, len_i, off_j, len_j, swaped=None, nthreads=1):
        """mode 0 = cost only, 1 = align+traceback.  -> (seconds, int32 costs)."""
        seqs = np.ascontiguousarray(seqs, np.uint8)
        off_i = np.ascontiguousarray(off_i, np.int64); off_j = np.ascontiguousarray(off_j, np.int64)
        len_i = np.ascontiguousarray(len_i, np.int32); len_j = np.ascontiguousarray(len_j, np.int32)
        n = len(len_i)
        cost = np.zeros(n, np.int32)
        sw = None if swaped is None else np.ascontiguousarray(swaped, np.uint8)
        i64p = C.POINTER(C.c_longlong)
        t = self.lib.do_batch_affine(cm, mode, n, _p8(seqs), off_i.ctypes.data_as(i64p), len_i.ctypes.data_as(_i32p),
                                     off_j.ctypes.data_as(i64p), len_j.ctypes.data_as(_i32p),
                                     None if sw is None else _p8(sw), cost.ctypes.data_as(_i32p), int(nthreads))
        return t, cost

    # -- Sequence.NewkkAlign, affine (src/newkkonen.c) ------------------------------------------
    def newkk_align(self, cm, s1, s2, swaped=0, with_stats=False):
        """newkkonen_CAML_algn_affine + newkkonen_CAML_backtrace_affine; s1 must be the shorter sequence.
        -> (cost, aligned s1, aligned s2[, stats])"""
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        cap = len(s1) + len(s2) + 2
        o1 = np.zeros(cap, np.uint8); o2 = np.zeros(cap, np.uint8)
        lens = (C.c_int * 2)()
        st = AlignStats()
        r = self.lib.do_newkk_align_affine(cm, _p8(s1), len(s1), _p8(s2), len(s2), int(swaped), _p8(o1), _p8(o2), lens, C.byref(st))
        if r == INT_MIN:
            raise RuntimeError("newkkonen: pass the shorter one as first / invalid dir")
        res = (r, o1[:lens[0]].copy(), o2[:lens[1]].copy())
        return res + (st,) if with_stats else res

    # -- linear gap ---------------------------------------------------------------------------
    def cost_linear(self, cm, s1, s2, deltawh, with_stats=False):
        """algn_CAML_simple_2 semantics: s1 must be the shorter sequence."""
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        st = AlignStats()
        r = self.lib.do_cost_linear(cm, self.lin, _p8(s1), len(s1), _p8(s2), len(s2), int(deltawh), C.byref(st))
        if r == INT_MIN:
            raise RuntimeError("pass the shorter one as first")
        return (r, st) if with_stats else r

    def align_linear(self, cm, s1, s2, deltawh, swaped=0):
        """algn_CAML_align_2d = simple_2 + backtrace_2d.  -> (cost, r1, r2)"""
        s1 = np.ascontiguousarray(s1, np.uint8); s2 = np.ascontiguousarray(s2, np.uint8)
        cost = self.cost_linear(cm, s1, s2, deltawh)
        cap = len(s1) + len(s2)
        o1 = np.zeros(cap, np.uint8); o2 = np.zeros(cap, np.uint8)
        lens = (C.c_int * 2)()
        self.lib.do_backtrace_linear(self.lin, _p8(s1), _p8(s2), int(swaped), _p8(o1), _p8(o2), lens)
        return cost, o1[:lens[0]].copy(), o2[:lens[1]].copy()

    # -- O(L) helpers -------------------------------------------------------------------------------
    def median_2(self, cm, a, b, with_gaps):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.lib.do_median_2(cm, _p8(a), _p8(b), len(a), int(with_gaps), _p8(out))
        return out[:n].copy()

    def union(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a), np.uint8)
        self.lib.do_union(_p8(a), _p8(b), len(a), _p8(out))
        return out

    def worst_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.do_worst_2(cm, _p8(a), _p8(b), len(a))

    def verify_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.do_verify_2(cm, _p8(a), _p8(b), len(a))

    def ancestor_2(self, cm, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.lib.do_ancestor_2(cm, _p8(a), _p8(b), len(a), _p8(out))
        if n == INT_MIN:
            raise RuntimeError("median should not be 0")
        return out[:n].copy()
