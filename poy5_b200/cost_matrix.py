"""Host mirror of ``Cost_matrix.Two_D`` (src/cost_matrix.ml): table construction is done by the
C++ side of the library (poy_cm_fill); this module only names things the way the reference does."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import CmHost, Cm3dHost


class Two_D:
    """A pair (c2_full, c2_original) of host-side cost models, like the tuple the reference
    threads through ``Data`` (src/seqCS.ml:51-55)."""

    def __init__(self, full, original):
        self.full = full
        self.original = original

    @staticmethod
    def of_list(rows, gap_opening=None):
        """Cost_matrix.Two_D.of_list (src/cost_matrix.ml:1257-1266) followed, when gap_opening is
        given, by set_cost_model (Affine go) on both matrices (src/data.ml:5937-5964)."""
        L = _lib.load()
        single = (C.c_int32 * 25)(*[int(v) for r in rows for v in r])
        full, orig = CmHost(), CmHost()
        st = L.poy_cm_fill(single, -1 if gap_opening is None else int(gap_opening), C.byref(full), C.byref(orig))
        if st != 0:
            raise _lib.PoyError(st, L.poy_status_string(st).decode())
        return Two_D(full, orig)

    @staticmethod
    def of_transformations_and_gaps(trans, gaps, gap_opening=None):
        """src/cost_matrix.ml:1333-1344 for the 5-letter DNA alphabet."""
        n = 5
        rows = [[0 if x == p else (gaps if (x == n - 1 or p == n - 1) else trans) for x in range(n)] for p in range(n)]
        return Two_D.of_list(rows, gap_opening)


def tables(cm_host):
    """numpy views (32x32) of a CmHost, for tests."""
    return dict(cost=np.ctypeslib.as_array(cm_host.cost).reshape(32, 32).copy(),
                worst=np.ctypeslib.as_array(cm_host.worst).reshape(32, 32).copy(),
                median=np.ctypeslib.as_array(cm_host.median).reshape(32, 32).copy(),
                prepend=np.ctypeslib.as_array(cm_host.prepend).copy(),
                tail=np.ctypeslib.as_array(cm_host.tail).copy(),
                gap_open=cm_host.gap_open, cost_model_type=cm_host.cost_model_type)


def min_non0(cm_host):
    return _lib.load().poy_cm_min_non0(C.byref(cm_host))


def get_closest(cm_host, a, b):
    return _lib.load().poy_cm_get_closest(C.byref(cm_host), int(a), int(b))


class Three_D:
    """Cost_matrix.Three_D (src/cost_matrix.ml:1440-1760) for the DNA bitset alphabet: host tables + device handle."""

    def __init__(self, host):
        self.host = host

    @staticmethod
    def of_two_dim(cm_host):
        """Cost_matrix.Three_D.of_two_dim -> of_two_dim_comb (src/cost_matrix.ml:1605-1652, 1707-1724)"""
        L = _lib.load()
        out = Cm3dHost()
        st = L.poy_cm3d_fill(C.byref(cm_host), C.byref(out))
        if st != 0:
            raise _lib.PoyError(st, L.poy_status_string(st).decode())
        return Three_D(out)

    def tables(self):
        return dict(cost=np.ctypeslib.as_array(self.host.cost).reshape(32, 32, 32).copy(),
                    median=np.ctypeslib.as_array(self.host.median).reshape(32, 32, 32).copy())
