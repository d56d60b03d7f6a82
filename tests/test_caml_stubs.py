"""The reference-shaped single-pair entry points (integration/poy_caml_stubs.c: algn_CAML_cost_affine_3,
algn_CAML_align_affine_3[_bc] with the reference's OCaml-value signatures, each a batch of one through the C ABI).
OCaml values (struct seq / struct cm custom blocks) are made by the compiled reference itself (oracle/_ref), handed
to BOTH implementations of the same symbol, and the results compared."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cost_matrix_oracle as cmo
from poy5_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "integration", "_build", "libpoycamlstubs.so")
REF = os.path.join(ROOT, "oracle", "_ref", "libpoyref.so")
_u8p = C.POINTER(C.c_ubyte)


def test_stubs_library_exports_reference_symbols():
    if not os.path.exists(STUBS):
        pytest.skip("integration/_build/libpoycamlstubs.so not built (needs /root/reference for seq.h / cm.h)")
    out = os.popen("nm -D --defined-only %s" % STUBS).read()
    for sym in ("algn_CAML_cost_affine_3", "algn_CAML_align_affine_3", "algn_CAML_align_affine_3_bc"):
        assert (" T " + sym) in out


@pytest.mark.gpu
def test_stubs_match_reference_stubs():
    if not (os.path.exists(STUBS) and os.path.exists(REF)):
        pytest.skip("needs the prebuilt stubs and oracle/_ref")
    from oracle import refbind
    R = refbind.RefLib(False)
    ref = C.CDLL(REF, mode=C.RTLD_GLOBAL)      # provides caml_failwith for the stubs (the OCaml runtime's role)
    if not hasattr(ref, "ref_seq_new"):
        pytest.skip("oracle/_ref predates ref_seq_new")
    stubs = C.CDLL(STUBS)
    vp, lg = C.c_void_p, C.c_long
    ref.ref_seq_new.restype = vp; ref.ref_seq_new.argtypes = [_u8p, C.c_int, C.c_int]
    ref.ref_seq_read.argtypes = [vp, _u8p]; ref.ref_val_free.argtypes = [vp]
    ref.ref_val_int.restype = lg; ref.ref_int_val.argtypes = [lg]
    for L in (stubs, ref):
        L.algn_CAML_cost_affine_3.restype = lg; L.algn_CAML_cost_affine_3.argtypes = [vp, vp, vp, vp]
        L.algn_CAML_align_affine_3.restype = lg; L.algn_CAML_align_affine_3.argtypes = [vp] * 8 + [lg]

    def mk(s, cap=None):
        s = np.ascontiguousarray(s, np.uint8)
        return ref.ref_seq_new(s.ctypes.data_as(_u8p), len(s), len(s) if cap is None else cap)

    def rd(v, cap):
        buf = np.zeros(cap + 4, np.uint8)
        n = ref.ref_seq_read(v, buf.ctypes.data_as(_u8p))
        return buf[:n].copy()

    checked = 0
    for reg in [(1, 1, 3), (2, 1, 5)]:
        full, _ = cmo.dna_matrices(*reg)
        cm = R.cm(full)
        seqs, ia, ib = synth.pair_batch(31 + reg[2], 12, 180, frac_decorated=0.4, jitter=0.3)
        for p in range(len(ia)):
            a, b = seqs[ia[p]], seqs[ib[p]]
            va, vb = mk(a), mk(b)
            c_ref = ref.ref_int_val(ref.algn_CAML_cost_affine_3(va, vb, cm, R.mat))
            c_new = ref.ref_int_val(stubs.algn_CAML_cost_affine_3(va, vb, cm, R.mat))
            assert c_ref == c_new
            sw = int(len(a) > len(b))
            vi, vj = (vb, va) if sw else (va, vb)
            cap = len(a) + len(b) + 2
            outs = {}
            for name, L in (("ref", ref), ("new", stubs)):
                res = [mk(np.zeros(0, np.uint8), cap) for _ in range(4)]     # resi, resj, median, medianwg
                cost = ref.ref_int_val(L.algn_CAML_align_affine_3(vi, vj, cm, R.mat, res[0], res[1], res[2], res[3], ref.ref_val_int(sw)))
                outs[name] = (cost, [rd(v, cap) for v in res])
                for v in res:
                    ref.ref_val_free(v)
            assert outs["ref"][0] == outs["new"][0]
            for x, y in zip(outs["ref"][1], outs["new"][1]):
                assert np.array_equal(x, y)
            ref.ref_val_free(va); ref.ref_val_free(vb)
            checked += 1
    assert checked == 24
