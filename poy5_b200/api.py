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

    def set_arena_limit(self, nbytes):
        self.check(self.L.poy_ctx_set_arena_limit(self.h, int(nbytes)))

    @property
    def launches(self):
        return int(self.L.poy_ctx_launch_count(self.h))

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
