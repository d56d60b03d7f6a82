"""ctypes loader for libpoy5b200.so -- fails loudly when the CUDA extension is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)

POY_OK = 0
STATUS = {0: "POY_OK", -1: "POY_ERR_CUDA", -2: "POY_ERR_NO_DEVICE", -3: "POY_ERR_ARG", -4: "POY_ERR_ORDER",
          -5: "POY_ERR_COST_RANGE", -6: "POY_ERR_MODEL", -7: "POY_ERR_NOMEM"}


class PoyError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__("%s (%d): %s" % (STATUS.get(status, "?"), status, message))


class CmHost(C.Structure):
    """poy_cm_host (include/poy5_b200.h) == struct cm tables (src/cm.h:33-76)."""
    _fields_ = [("cost", C.c_int32 * 1024), ("worst", C.c_int32 * 1024), ("median", C.c_uint8 * 1024),
                ("prepend", C.c_int32 * 32), ("tail", C.c_int32 * 32), ("gap_open", C.c_int32),
                ("cost_model_type", C.c_int32), ("is_identity", C.c_int32), ("is_metric", C.c_int32)]


class Cm3dHost(C.Structure):
    """poy_cm3d_host == struct cm_3d tables (src/cm.h:253-280), indexed (((a << 5) + b) << 5) + c."""
    _fields_ = [("cost", C.c_int32 * 32768), ("median", C.c_uint8 * 32768)]


def lib_path():
    return os.path.join(_HERE, "libpoy5b200.so")


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("libpoy5b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C poy5_b200/csrc`. There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.poy_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.poy_ctx_destroy.argtypes = [vp]
    L.poy_ctx_destroy.restype = None
    L.poy_last_error.argtypes = [vp]
    L.poy_last_error.restype = C.c_char_p
    L.poy_status_string.argtypes = [C.c_int]
    L.poy_status_string.restype = C.c_char_p
    L.poy_ctx_set_arena_limit.argtypes = [vp, C.c_uint64]
    L.poy_ctx_synchronize.argtypes = [vp]
    L.poy_ctx_trim.argtypes = [vp]
    L.poy_ctx_launch_count.argtypes = [vp]
    L.poy_ctx_launch_count.restype = C.c_uint64
    L.poy_ctx_stats.argtypes = [vp, vp]
    L.poy_cm_fill.argtypes = [i32p, C.c_int32, C.POINTER(CmHost), C.POINTER(CmHost)]
    L.poy_cm_min_non0.argtypes = [C.POINTER(CmHost)]
    L.poy_cm_get_closest.argtypes = [C.POINTER(CmHost), C.c_int32, C.c_int32]
    L.poy_cm_upload.argtypes = [vp, C.POINTER(CmHost), C.POINTER(vp)]
    L.poy_cm_free.argtypes = [vp, vp]
    L.poy_cm_free.restype = None
    L.poy_pool_upload.argtypes = [vp, vp, vp, C.c_int32, C.POINTER(vp)]
    L.poy_pool_from_device.argtypes = [vp, vp, vp, vp, C.c_int32, C.POINTER(vp)]
    L.poy_pool_free.argtypes = [vp, vp]
    L.poy_pool_free.restype = None
    L.poy_batch_cost_affine.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp]
    L.poy_batch_cost_affine_dev.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp]
    L.poy_batch_align_affine.argtypes = [vp, vp, vp, C.c_int32] + [vp] * 11
    L.poy_batch_align_affine_dev.argtypes = [vp, vp, vp, C.c_int32] + [vp] * 13
    L.poy_batch_cost_linear.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, vp]
    L.poy_batch_align_linear.argtypes = [vp, vp, vp, C.c_int32] + [vp] * 10
    L.poy_batch_median_2.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, vp]
    L.poy_batch_union.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    L.poy_batch_aligned_cost.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, vp]
    L.poy_batch_ancestor_2.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp]
    L.poy_batch_closest.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp]
    L.poy_dos_distance.argtypes = [vp, vp, vp, C.c_int32, vp, vp, C.c_int32, vp]
    L.poy_dos_median.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.poy_dos_median2.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.poy_store_create.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(vp)]
    L.poy_store_free.argtypes = [vp, vp]
    L.poy_store_free.restype = None
    L.poy_store_pool.argtypes = [vp]
    L.poy_store_pool.restype = vp
    L.poy_store_count.argtypes = [vp]
    L.poy_store_count.restype = C.c_int32
    L.poy_store_bytes.argtypes = [vp]
    L.poy_store_bytes.restype = C.c_int64
    L.poy_store_append.argtypes = [vp, vp, vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    L.poy_store_truncate.argtypes = [vp, vp, C.c_int32]
    L.poy_store_lengths.argtypes = [vp, C.c_int32, vp, vp]
    L.poy_store_read.argtypes = [vp, vp, C.c_int32, vp, vp, vp]
    L.poy_store_median.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    L.poy_store_distance.argtypes = [vp, vp, vp, C.c_int32, vp, vp, C.c_int32, vp]
    L.poy_cm3d_fill.argtypes = [C.POINTER(CmHost), C.POINTER(Cm3dHost)]
    L.poy_cm3d_upload.argtypes = [vp, C.POINTER(Cm3dHost), C.POINTER(vp)]
    L.poy_cm3d_free.argtypes = [vp, vp]
    L.poy_cm3d_free.restype = None
    L.poy_batch_median_3.argtypes = [vp, vp, C.c_int32] + [vp] * 10
    L.poy_batch_newkk_align.argtypes = [vp, vp, vp, C.c_int32] + [vp] * 9
    L.poy_microbench_int.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _LIB = L
    return L


EXPORTS = ["poy_ctx_create", "poy_ctx_destroy", "poy_last_error", "poy_status_string", "poy_ctx_set_arena_limit",
           "poy_ctx_synchronize", "poy_ctx_trim", "poy_ctx_launch_count", "poy_ctx_stats", "poy_cm_fill", "poy_cm_min_non0", "poy_cm_get_closest",
           "poy_cm_upload", "poy_cm_free", "poy_pool_upload", "poy_pool_from_device", "poy_pool_free",
           "poy_batch_cost_affine", "poy_batch_cost_affine_dev", "poy_batch_align_affine",
           "poy_batch_align_affine_dev", "poy_batch_cost_linear", "poy_batch_align_linear", "poy_batch_median_2", "poy_batch_union",
           "poy_batch_aligned_cost", "poy_batch_ancestor_2", "poy_batch_closest", "poy_dos_distance", "poy_dos_median", "poy_dos_median2",
           "poy_store_create", "poy_store_free", "poy_store_pool", "poy_store_count", "poy_store_bytes", "poy_store_append",
           "poy_store_truncate", "poy_store_lengths", "poy_store_read", "poy_store_median", "poy_store_distance",
           "poy_batch_newkk_align", "poy_cm3d_fill", "poy_cm3d_upload", "poy_cm3d_free",
           "poy_batch_median_3", "poy_microbench_int"]
