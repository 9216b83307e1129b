"""Generate tests/golden/*.npz from the CPU oracle (oracle/jx_oracle.c).

PARITY UNPINNED: the reference (Rust) cannot be built or imported here, and its own tests hold no
expected beta/se/p/lambda values for this path, so these goldens pin the ORACLE,
not the reference binary.  The toy case restates the reference's embedded recipe
(python/janusx/assoc/api.py:617-655: default_rng(42), n=8, m=5, K = I + three off-diagonals).

Run:  python tests/golden/make_golden.py      (deterministic; rewrites the .npz files)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from janusx_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def small_case():
    case = synth.make_case(n=96, m=40, q=2, seed=11, missing_rate=0.03)
    n = case.n
    ut = np.ascontiguousarray(case.u.T.astype(np.float32))
    X = np.concatenate([np.ones((n, 1)), case.cov], axis=1)
    xr, yr = O.lmm_rotate_x_y_with_ut_f64(ut, X, case.y)
    y = yr[:, 0].copy()
    lbd, ml0, reml0 = O.lmm_reml_null_f32(case.s, xr, y, -5.0, 5.0, 50, 1e-3)
    lo, hi = float(np.log10(lbd) - 2.0), float(np.log10(lbd) + 2.0)
    keep, af, mr, missing = O.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = O.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    rot = O.rotate_block(g, ut, mode=0)
    lmm, ev = O.lmm_reml_chunk_f32(case.s, xr, y, lo, hi, rot, 30, 1e-2, 1, None, return_evals=True)
    lmm4 = O.lmm_reml_chunk_f32(case.s, xr, y, lo, hi, rot, 30, 1e-2, 1, ml0)
    lml, ml_null = O.lmm_ml_null_brent(case.s, xr, y, lo, hi, 30, 1e-2, None)
    lmm2 = O.lmm_reml_lmm2_chunk_f32(case.s, xr, y, lo, hi, rot, ml_null, 30, 1e-2, 1)
    fixed, meta = O.lmm_assoc_chunk_f32(case.s, xr, y, float(np.log10(lbd)), rot, 1, None)
    # sample subset: every third sample dropped
    sidx = np.array([j for j in range(n) if j % 3 != 1], dtype=np.int64)
    keep_s, af_s, mr_s, missing_s = O.count_qc_block(case.packed, n, sidx, 0.02, 0.05, 1.0)
    idx_s = np.nonzero(keep_s)[0]
    g_s = O.decode_centered_block(case.packed, n, af_s[idx_s], sample_idx=sidx, row_indices=idx_s)
    np.savez_compressed(
        OUT / "small_n96.npz",
        packed=case.packed, n=np.int64(n), y_raw=case.y, cov=case.cov, s=case.s, u=case.u,
        xcov=xr, y=y, null=np.array([lbd, ml0, reml0]), bounds=np.array([lo, hi]),
        keep=keep, af=af, miss_rate=mr, missing=missing, g=g, rot=rot,
        lmm=lmm, lmm_evals=ev, lmm4=lmm4, ml_null=np.array([lml, ml_null]), lmm2=lmm2,
        fixed=fixed, fixed_meta=np.array([meta["ypy"], meta["log_det_v"], meta["df"]]),
        sidx=sidx, keep_s=keep_s, af_s=af_s, missing_s=missing_s, g_s=g_s,
    )


def toy_case():
    rng = np.random.default_rng(42)
    n, m = 8, 5
    y = rng.normal(size=n)
    X = np.stack([rng.normal(size=n), np.linspace(-1.0, 1.0, n, dtype=np.float64)], axis=1)
    G = rng.normal(size=(n, m))
    K = np.eye(n)
    K[0, 1] = K[1, 0] = 0.15
    K[2, 3] = K[3, 2] = 0.08
    K[4, 5] = K[5, 4] = 0.05
    Kr = K.copy()
    Kr[np.diag_indices(n)] += 1e-6
    s, u = np.linalg.eigh(Kr)
    ut = np.ascontiguousarray(u.T.astype(np.float32))
    Xd = np.concatenate([np.ones((n, 1)), X], axis=1)
    xr, yr = O.lmm_rotate_x_y_with_ut_f64(ut, Xd, y)
    yv = yr[:, 0].copy()
    lbd, ml0, reml0 = O.lmm_reml_null_f32(s, xr, yv, -5.0, 5.0, 50, 1e-3)
    snp = np.ascontiguousarray(G.T.astype(np.float32))  # API passes G as given (api.py:530-531)
    # bounds rule pyBLUP/assoc.py:1873-1876 needs pve; the toy K ~ I gives pve inside (0.05,0.95) or not:
    res_wide = O.lmm_reml_chunk_from_snp_f32(s, xr, yv, -5.0, 5.0, snp, ut, 30, 1e-2, 1)
    lo, hi = float(np.log10(lbd) - 2.0), float(np.log10(lbd) + 2.0)
    res_narrow = O.lmm_reml_chunk_from_snp_f32(s, xr, yv, lo, hi, snp, ut, 30, 1e-2, 1)
    np.savez_compressed(OUT / "toy_n8.npz", y=y, X=X, G=G, K=K, s=s, u=u, xcov=xr, yrot=yv,
                        null=np.array([lbd, ml0, reml0]), res_wide=res_wide, res_narrow=res_narrow)


if __name__ == "__main__":
    small_case()
    toy_case()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
