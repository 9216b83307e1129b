"""Pin the Python glue against the reference's OWN Python code.

`python/janusx/pyBLUP/assoc.py` (LMM / LMM2 / FvLMM: null-model bookkeeping, pve, bounds, the scipy null-ML fit, the
gwas() call conventions) is pure Python on top of the native module `janusx.janusx`, which cannot be built here (Rust).
This script imports that reference file UNMODIFIED with `janusx.janusx` replaced by a stub whose functions are the CPU
oracle (oracle/oracle.py) and numpy's eigh, runs the reference classes on the small synthetic case, and stores what they
produce.  tests/test_assoc_glue_cpu.py then runs janusx_b200/assoc.py on the same oracle-backed functions and must get
the same numbers: whatever the native numerics are, the glue on top of them is the reference's.  (The native numerics
themselves stay "parity unpinned")

Run here only (needs /root/reference):  python tests/golden/make_ref_python_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import oracle as O  # noqa: E402
from janusx_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def oracle_backed_native():
    """The subset of `janusx.janusx` the LMM classes touch, served by the oracle (same signatures)."""
    m = types.ModuleType("janusx.janusx")

    def _missing(*a, **k):
        raise RuntimeError("not part of the exact-LMM path")

    for name in ("fastlmm_prepare_lowrank_f64", "fastlmm_assoc_from_snp_f32", "lm_block_assoc_f32", "fastlmm_reml_chunk_f32",
                 "fastlmm_reml_null_f32", "fastlmm_assoc_chunk_f32", "fvlmm_assoc_chunk_from_snp_to_tsv_f32",
                 "fvlmm_assoc_bed_to_tsv_f32"):
        setattr(m, name, _missing)
    # importable but None: the reference then takes its uncached fixed-lambda route (same arithmetic)
    for name in ("fvlmm_assoc_prepare_cache_f32", "fvlmm_assoc_chunk_with_cache_f32", "fvlmm_assoc_chunk_from_snp_with_cache_f32"):
        setattr(m, name, None)

    def rust_eigh_from_array_f64(a, threads=0, driver=None, jobz="V", require_lapack=False):
        w, v = np.linalg.eigh(np.asarray(a, dtype=np.float64))
        return w, v, "numpy", "lapack_numpy_eigh", int(w.shape[0]), 0, 0, 0, True, 0.0

    m.rust_eigh_from_array_f64 = rust_eigh_from_array_f64
    m.rust_eigh_from_array_f64_inplace = rust_eigh_from_array_f64
    m.lmm_rotate_x_y_with_ut_f64 = lambda u_t, x, y, threads=0: O.lmm_rotate_x_y_with_ut_f64(u_t, x, y)
    m.lmm_reml_null_f32 = lambda s, xc, y, low, high, max_iter=50, tol=1e-2: O.lmm_reml_null_f32(s, xc, y, low, high, max_iter, tol)
    m.ml_loglike_null_f32 = lambda s, xc, y, l10: O.ml_loglike_null_f32(s, xc, y, l10)
    m.lmm_reml_chunk_f32 = lambda s, xc, y, low, high, g, max_iter=50, tol=1e-2, threads=0, nullml=None: \
        O.lmm_reml_chunk_f32(s, xc, y, low, high, g, max_iter, tol, threads, nullml)
    m.lmm_reml_chunk_from_snp_f32 = lambda s, xc, y, low, high, g, ut, max_iter=50, tol=1e-2, threads=0, nullml=None, \
        rotate_block_rows=256: O.lmm_reml_chunk_from_snp_f32(s, xc, y, low, high, g, ut, max_iter, tol, threads, nullml)
    m.lmm_reml_lmm2_chunk_from_snp_f32 = lambda s, xc, y, low, high, g, ut, nullml, max_iter=50, tol=1e-2, threads=0, \
        rotate_block_rows=256: O.lmm_reml_lmm2_chunk_from_snp_f32(s, xc, y, low, high, g, ut, nullml, max_iter, tol, threads)

    def _fixed(res):
        return res[0] if isinstance(res, tuple) else res

    m.lmm_assoc_chunk_f32 = lambda s, xc, y, l10, g, threads=0, nullml=None: _fixed(O.lmm_assoc_chunk_f32(s, xc, y, l10, g, threads, nullml))
    m.lmm_assoc_chunk_from_snp_f32 = lambda s, xc, y, l10, g, ut, threads=0, nullml=None, rotate_block_rows=512: \
        _fixed(O.lmm_assoc_chunk_from_snp_f32(s, xc, y, l10, g, ut))
    m.fvlmm_assoc_chunk_f32 = m.lmm_assoc_chunk_f32
    m.fvlmm_assoc_chunk_from_snp_f32 = m.lmm_assoc_chunk_from_snp_f32
    return m


def load_reference_assoc():
    sys.path.insert(0, "/root/reference/python")
    import janusx  # noqa: F401  (pure-Python package root)
    sys.modules["janusx.janusx"] = oracle_backed_native()
    import importlib
    return importlib.import_module("janusx.pyBLUP.assoc")


def main():
    ref = load_reference_assoc()
    case = synth.make_case(n=96, m=40, q=2, seed=11, missing_rate=0.03)
    K = case.u @ np.diag(case.s) @ case.u.T
    K = 0.5 * (K + K.T)
    keep, af, _, _ = O.count_qc_block(case.packed, case.n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = O.decode_centered_block(case.packed, case.n, af[idx], row_indices=idx)
    out = {"K": K, "y": case.y, "cov": case.cov, "g": g}
    for cls_name in ("LMM", "LMM2", "FvLMM"):
        obj = getattr(ref, cls_name)(case.y, case.cov, K)
        res = np.asarray(obj.gwas(g, threads=1), dtype=np.float64)
        out[f"{cls_name}_gwas"] = res
        if cls_name == "LMM":
            for attr in ("S", "Xcov", "y", "lbd_null", "sigma_g2_null", "sigma_e2_null", "pve", "pve_vc_ratio_raw", "trace_mean",
                         "LL0", "ML0"):
                out[f"attr_{attr}"] = np.asarray(getattr(obj, attr), dtype=np.float64)
            out["attr_bounds"] = np.asarray(obj.bounds, dtype=np.float64)
            out["attr_Dh_abs"] = np.abs(np.asarray(obj.Dh, dtype=np.float32))     # eigenvector signs are arbitrary
            out["vc"] = np.asarray(ref._lmm_profile_exact_vc(obj.S, obj.Xcov, obj.y, obj.lbd_null), dtype=np.float64)
            out["ml_null"] = np.asarray(ref.lmm_ml_null(obj.S, obj.Xcov, obj.y, obj.bounds, max_iter=30, tol=1e-2), dtype=np.float64)
        if cls_name == "LMM2":
            out["LMM2_ml0_exact"] = np.float64(getattr(obj, "_lmm2_ml0_exact", np.nan))
    # a second trait whose pve leaves [0.05, 0.95]: the (-5, 5) bounds branch
    y_flat = np.random.default_rng(5).normal(size=case.n)
    obj = ref.LMM(y_flat, None, K)
    out["flat_y"] = y_flat
    out["flat_pve"], out["flat_bounds"], out["flat_lbd"] = np.float64(obj.pve), np.asarray(obj.bounds, dtype=np.float64), np.float64(obj.lbd_null)
    np.savez_compressed(OUT / "ref_python_glue_n96.npz", **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if k.startswith(("attr_l", "attr_p", "attr_b", "flat_p", "flat_b", "ml_null", "LMM2_ml0"))})


if __name__ == "__main__":
    main()
