"""poy5_b200 -- B200-native batched direct-optimization alignment (the DO hot path of amnh/poy5).

The product is the CUDA library ``libpoy5b200.so`` (C ABI: ``include/poy5_b200.h``).  This
package is a thin ctypes mirror of that ABI plus host-side classes named after the reference's
OCaml modules (``Cost_matrix.Two_D``, ``Sequence.Align``, ``SeqCS.DOS``).  There is no CPU
implementation of any alignment here: importing works without a GPU (so the build can be
checked), but creating a :class:`Context` raises unless a CUDA device is present.
"""
from ._lib import PoyError, lib_path, load  # noqa: F401
from .api import Context, CostModel, CostModel3D, Pool, Store  # noqa: F401
from . import cost_matrix, sequence, seqcs  # noqa: F401
