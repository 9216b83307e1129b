"""The Python glue against the reference's OWN Python code.

tests/golden/ref_python_glue_n96.npz holds what python/janusx/pyBLUP/assoc.py (imported unmodified) produces when its
native module `janusx.janusx` is served by the CPU oracle (tests/golden/make_ref_python_golden.py).  Here
janusx_b200/assoc.py runs on the same oracle-backed functions (a stand-in for the device model: test infrastructure
only) and must reproduce those numbers -- null-model bookkeeping, pve, bounds branches, the scipy null-ML fit and the
gwas() call conventions are then the reference's, whatever the native numerics underneath.
A second test imports the reference's Python layer on top of janusx_b200.jxrs itself (import only) to prove that every
name it pulls from `janusx.janusx` on this path resolves; it needs /root/reference and is skipped where that is absent.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden" / "ref_python_glue_n96.npz"


def _fake_device(O):
    class FakeDev:
        def __init__(self, s, xcov, y_rot, u_t=None, device=0, u_t_on_device=False):
            self.s, self.xcov, self.y = np.asarray(s, float), np.asarray(xcov, float), np.asarray(y_rot, float).reshape(-1)
            self.ut = None if u_t is None else np.ascontiguousarray(u_t, dtype=np.float32)
            self.n, self.p = self.y.shape[0], self.xcov.shape[1]

        def rotate_xy(self, x, y):
            return O.lmm_rotate_x_y_with_ut_f64(self.ut, x, y)

        def set_xy(self, xcov, y):
            self.xcov, self.y = np.asarray(xcov, float), np.asarray(y, float).reshape(-1)

        def reml_null(self, low, high, max_iter=50, tol=1e-2):
            return O.lmm_reml_null_f32(self.s, self.xcov, self.y, low, high, max_iter, tol)

        def ml_loglike_null(self, x):
            return O.ml_loglike_null_f32(self.s, self.xcov, self.y, x)

        def lmm_reml_chunk(self, g, low, high, max_iter=50, tol=1e-2, nullml=None, rotated=True, init=None, return_evals=False):
            if rotated:
                return O.lmm_reml_chunk_f32(self.s, self.xcov, self.y, low, high, g, max_iter, tol, 0, nullml)
            return O.lmm_reml_chunk_from_snp_f32(self.s, self.xcov, self.y, low, high, g, self.ut, max_iter, tol, 0, nullml)

        def lmm2_chunk(self, g, low, high, nullml, max_iter=50, tol=1e-2, rotated=False, init=None, return_evals=False):
            assert not rotated
            return O.lmm_reml_lmm2_chunk_from_snp_f32(self.s, self.xcov, self.y, low, high, g, self.ut, nullml, max_iter, tol)

        def fixed_chunk(self, g, log10_lbd, nullml=None, rotated=False, return_meta=False):
            res = (O.lmm_assoc_chunk_f32(self.s, self.xcov, self.y, log10_lbd, g, 0, nullml) if rotated
                   else O.lmm_assoc_chunk_from_snp_f32(self.s, self.xcov, self.y, log10_lbd, g, self.ut))
            return res[0] if isinstance(res, tuple) else res

    return FakeDev


def test_glue_matches_reference_python(oracle, monkeypatch):
    from janusx_b200 import assoc
    G = np.load(GOLDEN)
    monkeypatch.setattr(assoc, "DeviceModel", _fake_device(oracle))
    monkeypatch.setattr(assoc, "_eigh", lambda k, device=0: np.linalg.eigh(np.asarray(k, dtype=np.float64)))
    # _lmm_profile_exact_vc is pure numpy on both sides
    vc = assoc._lmm_profile_exact_vc(G["attr_S"], G["attr_Xcov"], G["attr_y"], float(G["attr_lbd_null"]))
    np.testing.assert_allclose(vc, G["vc"], rtol=1e-13)
    lmm = assoc.LMM(G["y"], G["cov"], G["K"])
    np.testing.assert_allclose(lmm.S, G["attr_S"], rtol=1e-13, atol=1e-15)
    assert np.array_equal(np.abs(lmm.Dh), G["attr_Dh_abs"]) and lmm.Dh.dtype == np.float32
    np.testing.assert_allclose(lmm.Xcov, G["attr_Xcov"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.asarray(lmm.y).reshape(-1), G["attr_y"].reshape(-1), rtol=1e-12, atol=1e-13)
    for attr in ("lbd_null", "sigma_g2_null", "sigma_e2_null", "pve", "pve_vc_ratio_raw", "trace_mean", "LL0", "ML0"):
        np.testing.assert_allclose(getattr(lmm, attr), float(G[f"attr_{attr}"]), rtol=1e-12, err_msg=attr)
    np.testing.assert_allclose(np.asarray(lmm.bounds, dtype=float), G["attr_bounds"], rtol=1e-13)
    np.testing.assert_allclose(lmm.gwas(G["g"]), G["LMM_gwas"], rtol=1e-12, equal_nan=True)
    lbd_ml, ml0 = assoc.lmm_ml_null(lmm, lmm.bounds, max_iter=30, tol=1e-2)
    np.testing.assert_allclose([lbd_ml, ml0], G["ml_null"], rtol=1e-12)
    lmm2 = assoc.LMM2(G["y"], G["cov"], G["K"])
    np.testing.assert_allclose(lmm2.gwas(G["g"]), G["LMM2_gwas"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(lmm2._lmm2_ml0_exact, float(G["LMM2_ml0_exact"]), rtol=1e-13)
    fv = assoc.FvLMM(G["y"], G["cov"], G["K"])
    np.testing.assert_allclose(fv.gwas(G["g"]), G["FvLMM_gwas"], rtol=1e-12, equal_nan=True)
    # the pve-outside-[0.05, 0.95] branch: bounds fall back to (-5, 5)
    flat = assoc.LMM(G["flat_y"], None, G["K"])
    np.testing.assert_allclose(flat.pve, float(G["flat_pve"]), rtol=1e-10)
    assert tuple(flat.bounds) == (-5, 5) == tuple(G["flat_bounds"])


@pytest.mark.skipif(not Path("/root/reference/python/janusx/pyBLUP/assoc.py").exists(),
                    reason="needs the reference checkout (not present on the GPU box)")
def test_reference_python_layer_imports_on_top_of_jxrs():
    """`janusx.janusx` := janusx_b200.jxrs.  The reference's pyBLUP/assoc.py hard-imports ten native names and soft-imports
    ten more in ONE try block (pyBLUP/assoc.py:207-246): a single missing name silently disables the whole group."""
    from janusx_b200 import jxrs
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "janusx" or k.startswith("janusx.")}
    sys.path.insert(0, "/root/reference/python")
    try:
        for k in saved:
            sys.modules.pop(k, None)
        import janusx  # noqa: F401
        sys.modules["janusx.janusx"] = jxrs
        ref = importlib.import_module("janusx.pyBLUP.assoc")
        for name in ("_lmm_reml_chunk_from_snp_f32", "_lmm_reml_lmm2_chunk_from_snp_f32", "_lmm_assoc_chunk_from_snp_f32",
                     "_fvlmm_assoc_chunk_f32", "_fvlmm_assoc_chunk_from_snp_f32", "_fvlmm_assoc_chunk_from_snp_to_tsv_f32",
                     "_fvlmm_assoc_bed_to_tsv_f32", "_fvlmm_assoc_prepare_cache_f32", "_fvlmm_assoc_chunk_with_cache_f32",
                     "_fvlmm_assoc_chunk_from_snp_with_cache_f32", "_lmm_rotate_x_y_with_ut_f64", "_rust_eigh_from_array_f64",
                     "_rust_eigh_from_array_f64_inplace"):
            assert getattr(ref, name) is not None, name
        assert ref.lmm_reml_chunk_f32 is jxrs.lmm_reml_chunk_f32 and ref.lmm_reml_null_f32 is jxrs.lmm_reml_null_f32
        # positional call conventions of the reference wrappers (pyBLUP/assoc.py:777-790, 832-845, 1512-1522)
        import inspect
        assert len(inspect.signature(jxrs.lmm_reml_chunk_from_snp_f32).parameters) == 12
        assert len(inspect.signature(jxrs.lmm_reml_lmm2_chunk_from_snp_f32).parameters) == 12
        assert len(inspect.signature(jxrs.fvlmm_assoc_chunk_from_snp_f32).parameters) == 9
        assert len(inspect.signature(jxrs.fvlmm_assoc_chunk_from_snp_with_cache_f32).parameters) == 6
        with pytest.raises(NotImplementedError):
            jxrs.fastlmm_reml_null_f32()
    finally:
        sys.path.remove("/root/reference/python")
        for k in [k for k in sys.modules if k == "janusx" or k.startswith("janusx.")]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
