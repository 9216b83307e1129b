"""LMM / LMM2 / FvLMM model objects with the reference's Python interface, device-backed.

Mirrors python/janusx/pyBLUP/assoc.py:1586-2182 (class LMM, LMM2, FvLMM): same constructor arguments,
attributes (`S, Dh, Xcov, y, lbd_null, pve, LL0, ML0, bounds`) and `.gwas(snp_chunk, threads)` return
layout.  The host-side bookkeeping (bounds, pve) is plain numpy like the reference; every n-length or
n x n computation (X/y rotation, null REML/ML fits, chunk scans) runs on the GPU through jxrs.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
from scipy.optimize import minimize_scalar

from . import jxrs
from .jxrs import DeviceModel


def _lmm_profile_exact_vc(S, Xcov, y_rot, lbd):
    """pyBLUP/assoc.py:907-951 -- profile REML variance components at the null lambda (p x p algebra)."""
    lbd_f = float(lbd)
    if not np.isfinite(lbd_f) or lbd_f <= 0.0:
        return float("nan"), float("nan")
    s = np.maximum(np.asarray(S, dtype=np.float64).reshape(-1), 0.0)
    x = np.asarray(Xcov, dtype=np.float64)
    y = np.asarray(y_rot, dtype=np.float64).reshape(-1)
    n, p = x.shape
    if n - p <= 0:
        return float("nan"), float("nan")
    v_inv = 1.0 / np.maximum(s + lbd_f, 1e-30)
    xtx = (x.T * v_inv) @ x
    xty = (x.T * v_inv) @ y
    try:
        beta = np.linalg.solve(xtx, xty)
    except np.linalg.LinAlgError:
        beta = np.linalg.lstsq(xtx, xty, rcond=None)[0]
    resid = y - x @ beta
    q = float(np.dot(v_inv, np.square(resid)))
    if not np.isfinite(q) or q <= 0.0:
        return float("nan"), float("nan")
    sg2 = q / float(n - p)
    return float(sg2), float(lbd_f * sg2)


class LMM:
    """Exact LMM GWAS with a ridge-stabilised full-rank spectral GRM (pyBLUP/assoc.py:1586-1994)."""

    _GRM_EIGH_RIDGE = 1e-6

    def __init__(self, y: np.ndarray, X: Optional[np.ndarray], kinship: np.ndarray, device: int = 0):
        y_arr = np.asarray(y).reshape(-1, 1)
        X_design = (np.concatenate([np.ones((y_arr.shape[0], 1)), X], axis=1) if X is not None
                    else np.ones((y_arr.shape[0], 1)))
        t0 = time.time()
        k = np.array(kinship, dtype=np.float64, copy=True)
        k.flat[:: k.shape[0] + 1] += float(self._GRM_EIGH_RIDGE)   # assoc.py:1626-1629
        evals, evecs = _eigh(k, device)
        self._initialize_from_spectral(y=y_arr, X=X_design, eigvals=evals, eigvecs=evecs,
                                       evd_secs=time.time() - t0, device=device)

    @classmethod
    def from_spectral(cls, y, X, eigvals, eigvecs, evd_secs: float = 0.0, device: int = 0) -> "LMM":
        y_arr = np.asarray(y).reshape(-1, 1)
        X_design = (np.concatenate([np.ones((y_arr.shape[0], 1)), X], axis=1) if X is not None
                    else np.ones((y_arr.shape[0], 1)))
        obj = cls.__new__(cls)
        obj._initialize_from_spectral(y=y_arr, X=X_design, eigvals=np.asarray(eigvals, dtype=np.float64),
                                      eigvecs=np.asarray(eigvecs, dtype=np.float64), evd_secs=float(evd_secs),
                                      device=device)
        return obj

    def _initialize_from_spectral(self, *, y, X, eigvals, eigvecs, evd_secs, device=0):
        s_full = np.ascontiguousarray(np.asarray(eigvals, dtype=np.float64).reshape(-1))
        u_full = np.asarray(eigvecs, dtype=np.float64)
        y_vec = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        X_design = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        n = int(y_vec.shape[0])
        if s_full.shape[0] != n:
            raise ValueError(f"eigvals length mismatch: got {s_full.shape[0]}, expected {n}")
        if u_full.shape != (n, n):
            raise ValueError(f"eigvecs shape mismatch: got {u_full.shape}, expected ({n}, {n})")
        if X_design.shape[0] != n:
            raise ValueError(f"design row mismatch: got {X_design.shape[0]}, expected {n}")
        self.n = n
        self.rank = n
        self.lowrank = False
        self.full_rank = True
        self.evd_secs = float(evd_secs)
        self.trace_mean = float(np.sum(np.clip(s_full, 0.0, None), dtype=np.float64) / float(max(1, n)))
        self.S = s_full
        # assoc.py:1818: U^T is stored as float32
        self.Dh = np.ascontiguousarray(u_full.T.astype(np.float32))
        # one resident device model per LMM object: U^T uploaded once
        self._dev = DeviceModel(self.S, np.ones((n, X_design.shape[1])), np.zeros(n), self.Dh, device=device)
        xcov_rot, y_rot = self._dev.rotate_xy(X_design, y_vec)             # assoc.py:1824-1831
        self.Xcov = np.ascontiguousarray(xcov_rot)
        self.y = np.ascontiguousarray(y_rot)
        self._dev.set_xy(self.Xcov, self.y.reshape(-1))
        lbd_null, ml0, reml = self._dev.reml_null(-5.0, 5.0, max_iter=50, tol=1e-3)  # assoc.py:1832-1839
        sg2, se2 = _lmm_profile_exact_vc(self.S, self.Xcov, self.y, lbd_null)
        vg_null = float(np.mean(np.clip(self.S, 0.0, None)))
        self.lbd_null = float(lbd_null)
        self.sigma_g2_null, self.sigma_e2_null = float(sg2), float(se2)
        ssum = self.sigma_g2_null + self.sigma_e2_null
        if np.isfinite(ssum) and ssum > 0.0:                                  # assoc.py:1848-1861
            self.pve_vc_ratio_raw = float(self.sigma_g2_null / ssum)
            var_g = self.sigma_g2_null * max(self.trace_mean, 0.0)
            denom = var_g + self.sigma_e2_null
            self.pve = float(var_g / denom) if np.isfinite(denom) and denom > 0.0 else self.pve_vc_ratio_raw
        else:
            self.pve = (float(vg_null / (vg_null + self.lbd_null)) if (vg_null + self.lbd_null) > 0
                        else float("nan"))
            self.pve_vc_ratio_raw = float("nan")
        self.pve_component_ratio_raw = self.pve_vc_ratio_raw
        self.pve_pheno_scale = float(self.pve)
        self.LL0 = float(reml)
        self.ML0 = float(ml0)
        if self.pve > 0.95 or self.pve < 0.05 or (not np.isfinite(self.lbd_null)) or self.lbd_null <= 0.0:
            self.bounds = (-5, 5)                                             # assoc.py:1873-1876
        else:
            self.bounds = (np.log10(self.lbd_null) - 2, np.log10(self.lbd_null) + 2)

    @property
    def device_model(self) -> DeviceModel:
        return self._dev

    def gwas(self, snp: np.ndarray, threads: int = 1) -> np.ndarray:
        """assoc.py:1962-1994 -> f64[m, 3] beta, se, pwald (max_iter=30, tol=1e-2, midpoint start)."""
        return self._dev.lmm_reml_chunk(snp, self.bounds[0], self.bounds[1], max_iter=30, tol=1e-2, nullml=None,
                                        rotated=False)


class LMM2(LMM):
    """Exact LMM scan with Wald beta/se plus per-SNP ML for the LRT (assoc.py:1997-2033)."""

    def gwas(self, snp: np.ndarray, threads: int = 1) -> np.ndarray:
        ml0_exact = getattr(self, "_lmm2_ml0_exact", None)
        if ml0_exact is None or not np.isfinite(float(ml0_exact)):
            lbd_ml, ml0_exact = lmm_ml_null(self, self.bounds, max_iter=30, tol=1e-2)
            self._lmm2_lbd_null_ml = float(lbd_ml)
            self._lmm2_ml0_exact = float(ml0_exact)
        return self._dev.lmm2_chunk(snp, self.bounds[0], self.bounds[1], float(ml0_exact), max_iter=30, tol=1e-2,
                                    rotated=False)


def lmm_ml_null(model: LMM, bounds, max_iter: int = 30, tol: float = 1e-2):
    """assoc.py:954-991: scipy bounded scalar minimisation of -ML(null); each evaluation is a device call."""
    low, high = float(bounds[0]), float(bounds[1])
    if not (np.isfinite(low) and np.isfinite(high) and low < high):
        raise ValueError(f"Invalid bounds for null ML optimization: {bounds}")

    def _objective(x: float) -> float:
        ml = model.device_model.ml_loglike_null(float(x))
        return -ml if np.isfinite(ml) else 1e300

    opt = minimize_scalar(_objective, bounds=(low, high), method="bounded",
                          options={"maxiter": int(max_iter), "xatol": float(tol)})
    best = float(opt.x)
    return float(10.0 ** best), float(model.device_model.ml_loglike_null(best))


class FvLMM(LMM):
    """Fixed-variance LMM: the null lambda for the whole scan (assoc.py:2079-2182)."""

    def gwas_rotated(self, utsnp_chunk: np.ndarray, threads: int = 1) -> np.ndarray:
        if not (np.isfinite(self.lbd_null) and self.lbd_null > 0.0):
            raise RuntimeError("FvLMM.gwas_rotated requires a finite positive null lambda.")
        return self._dev.fixed_chunk(utsnp_chunk, float(np.log10(self.lbd_null)), nullml=None, rotated=True)

    def gwas(self, snp: np.ndarray, threads: int = 1) -> np.ndarray:
        if not (np.isfinite(self.lbd_null) and self.lbd_null > 0.0):
            return super().gwas(snp, threads=threads)
        return self._dev.fixed_chunk(snp, float(np.log10(self.lbd_null)), nullml=None, rotated=False)


def _eigh(k: np.ndarray, device: int = 0):
    """Null-model eigendecomposition through the library's own entry point (jxb_eigh: one cuSOLVER Xsyevd call;
    SURVEY 8a A17: LAPACK dsyevd in the reference, src/math/eigh.rs:1320 -- 'library call on GPU, not a hand
    kernel').  Ascending eigenvalues, eigenvectors in columns."""
    from ._cabi import check, lib, ptr, require_gpu

    require_gpu()
    k = np.ascontiguousarray(k, dtype=np.float64)
    n = k.shape[0]
    w = np.empty(n, dtype=np.float64)
    ut = np.empty((n, n), dtype=np.float64)
    check(lib().jxb_eigh(int(device), n, ptr(k), 0.0, ptr(w), ptr(ut), None))
    return w, ut.T
