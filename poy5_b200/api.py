"""Object wrappers over the C ABI: Context (one per GPU), CostModel, Pool."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import PoyError, CmHost


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """poy_ctx: stream + scratch arenas on one GPU.  Raises PoyError(POY_ERR_NO_DEVICE) without a GPU."""

    def __init__(self, device=0, stream=0):
        self.L = _lib.load()
        h = C.c_void_p()
        st = self.L.poy_ctx_create(int(device), C.c_void_p(stream), C.byref(h))
        if st != 0:
            raise PoyError(st, self.L.poy_status_string(st).decode())
        self.h = h
        self.device = device

    def check(self, st):
        if st != 0:
            raise PoyError(st, self.L.poy_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.L.poy_ctx_synchronize(self.h))

    def trim(self):
        """give the grow-only device scratch back to the driver (re-allocated on demand)"""
        self.check(self.L.poy_ctx_trim(self.h))

    def set_arena_limit(self, nbytes):
        self.check(self.L.poy_ctx_set_arena_limit(self.h, int(nbytes)))

    @property
    def launches(self):
        return int(self.L.poy_ctx_launch_count(self.h))

    def stats(self):
        """poy_ctx_stats: dict of cumulative counters (band cells computed, probe / full fills, rounds, ...)"""
        out = np.zeros(8, np.int64)
        self.check(self.L.poy_ctx_stats(self.h, _ptr(out)))
        return dict(launches=int(out[0]), band_cells=int(out[1]), probe_fills=int(out[2]), full_fills=int(out[3]),
                    repeated=int(out[4]), rounds=int(out[5]), pairs=int(out[6]))

    def microbench(self, kind):
        ops = C.c_double(); mhz = C.c_double()
        self.check(self.L.poy_microbench_int(self.h, kind, C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value

    def close(self):
        if getattr(self, "h", None):
            self.L.poy_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CostModel:
    """Device-resident 2-D cost model (poy_cm).  `host` is a _lib.CmHost."""

    def __init__(self, ctx, host):
        self.ctx = ctx
        self.host = host
        h = C.c_void_p()
        ctx.check(ctx.L.poy_cm_upload(ctx.h, C.byref(host), C.byref(h)))
        self.h = h

    @property
    def gap_open(self):
        return self.host.gap_open

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.poy_cm_free(self.ctx.h, self.h)
            self.h = None


class CostModel3D:
    """Device-resident 3-D cost model (poy_cm3d, struct cm_3d).  `three_d` is a cost_matrix.Three_D."""

    def __init__(self, ctx, three_d):
        self.ctx = ctx
        self.host = three_d.host
        h = C.c_void_p()
        ctx.check(ctx.L.poy_cm3d_upload(ctx.h, C.byref(self.host), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.poy_cm3d_free(self.ctx.h, self.h)
            self.h = None


class Pool:
    """Device-resident packed sequences (poy_pool).  `seqs` is a list of uint8 arrays, each starting
    with the gap code 16; or pass (data, offsets) directly."""

    def __init__(self, ctx, seqs=None, data=None, offsets=None):
        self.ctx = ctx
        if seqs is not None:
            lens = np.fromiter((len(s) for s in seqs), np.int64, len(seqs))
            offsets = np.zeros(len(seqs) + 1, np.int64)
            np.cumsum(lens, out=offsets[1:])
            data = np.concatenate([np.asarray(s, np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
        self.data = np.ascontiguousarray(data, np.uint8)
        self.offsets = np.ascontiguousarray(offsets, np.int64)
        self.nseq = len(self.offsets) - 1
        self.lens = np.diff(self.offsets)
        h = C.c_void_p()
        ctx.check(ctx.L.poy_pool_upload(ctx.h, _ptr(self.data), _ptr(self.offsets), self.nseq, C.byref(h)))
        self.h = h

    def seq(self, s):
        return self.data[self.offsets[s]:self.offsets[s + 1]]

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.poy_pool_free(self.ctx.h, self.h)
            self.h = None


class Store:
    """Device-resident node store (poy_store): immutable sequences named by int ids, medians appended in HBM."""

    def __init__(self, ctx, cap_bytes=1 << 24, cap_seqs=4096):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.L.poy_store_create(ctx.h, int(cap_bytes), int(cap_seqs), C.byref(h)))
        self.h = h

    def __len__(self):
        return int(self.ctx.L.poy_store_count(self.h))

    @property
    def nbytes(self):
        return int(self.ctx.L.poy_store_bytes(self.h))

    def append(self, seqs):
        """host sequences -> int32 ids"""
        if not len(seqs):
            return np.zeros(0, np.int32)
        lens = np.fromiter((len(s) for s in seqs), np.int64, len(seqs))
        off = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        data = np.ascontiguousarray(np.concatenate([np.asarray(s, np.uint8) for s in seqs]))
        first = C.c_int32()
        self.ctx.check(self.ctx.L.poy_store_append(self.ctx.h, self.h, _ptr(data), _ptr(off), len(seqs), C.byref(first)))
        return np.arange(first.value, first.value + len(seqs), dtype=np.int32)

    def lengths(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(len(ids), np.int32)
        self.ctx.check(self.ctx.L.poy_store_lengths(self.h, len(ids), _ptr(ids), _ptr(out)))
        return out

    def read(self, ids):
        """ids -> list of uint8 arrays (device -> host)"""
        ids = np.ascontiguousarray(ids, np.int32)
        if not len(ids):
            return []
        lens = self.lengths(ids).astype(np.int64)
        off = np.zeros(len(ids) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        buf = np.zeros(max(1, int(off[-1])), np.uint8)
        self.ctx.check(self.ctx.L.poy_store_read(self.ctx.h, self.h, len(ids), _ptr(ids), _ptr(off), _ptr(buf)))
        return [buf[off[q]:off[q + 1]] for q in range(len(ids))]

    def median(self, h, a, b):
        """SeqCS.DOS.median over ids; medians are appended on the device.  -> (ids, lengths, cost2)"""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        n = len(a)
        ids = np.zeros(n, np.int32); ln = np.zeros(n, np.int32); c2 = np.zeros(n, np.int32)
        self.ctx.check(self.ctx.L.poy_store_median(self.ctx.h, self.h, h.c2_full.h, h.c2_original.h, n, _ptr(a), _ptr(b),
                                                   _ptr(ids), _ptr(ln), _ptr(c2)))
        return ids, ln, c2

    def distance(self, h, a, b, missing_distance=0):
        """SeqCS.DOS.distance over ids -> int32 costs"""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        cost = np.zeros(len(a), np.int32)
        self.ctx.check(self.ctx.L.poy_store_distance(self.ctx.h, self.h, h.c2_original.h, len(a), _ptr(a), _ptr(b),
                                                     int(missing_distance), _ptr(cost)))
        return cost

    def truncate(self, nseq):
        self.ctx.check(self.ctx.L.poy_store_truncate(self.ctx.h, self.h, int(nseq)))

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.poy_store_free(self.ctx.h, self.h)
            self.h = None
