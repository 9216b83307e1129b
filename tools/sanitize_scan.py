"""Small packed scans through every kernel family, for compute-sanitizer: tcgen05 int8 rotation + lane solve (forced),
warp solve, fixed-lambda (lane and warp), FP64 DMMA path (dominance coding), X/y rotation, null fits."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import make_problem  # noqa: E402

from janusx_b200 import jxrs  # noqa: E402

# small on purpose (memcheck runs kernels 10-100x slower); the spectral decomposition comes from the host so that no
# library kernel is instrumented
case = make_problem(n=160, m=300, q=2, seed=3, missing_rate=0.03)
n = case.n
ut = np.ascontiguousarray(case.u.T.astype(np.float32))
X = np.concatenate([np.ones((n, 1)), case.cov], axis=1)
mdl = jxrs.DeviceModel(case.s, np.ones((n, X.shape[1])), np.zeros(n), ut, device=0)
xcov, yrot = mdl.rotate_xy(X, case.y)
mdl.set_xy(xcov, yrot[:, 0])
lbd, _, _ = mdl.reml_null(-5.0, 5.0, 50, 1e-3)
l10 = float(np.log10(lbd))
kw = dict(low=l10 - 2.0, high=l10 + 2.0)
_, nullml = mdl.ml_null(kw["low"], kw["high"], 30, 1e-2, l10)
jxrs.set_thread_solve_min_rows(1)            # lane kernel on a small batch
jxrs._cabi.lib().jxb_set_fixed_lane_min_rows(1)
a = mdl.scan_packed(case.packed, n, mode="lmm2", init=l10, nullml=nullml, **kw)[3]
f = mdl.scan_packed(case.packed, n, mode="fvlmm", log10_lbd=l10)[3]
jxrs.set_thread_solve_min_rows(1 << 30)      # warp kernels
jxrs._cabi.lib().jxb_set_fixed_lane_min_rows(1 << 40)
b = mdl.scan_packed(case.packed, n, mode="lmm2", init=l10, nullml=nullml, **kw)[3]
f2 = mdl.scan_packed(case.packed, n, mode="fvlmm", log10_lbd=l10)[3]
d = mdl.scan_packed(case.packed, n, genetic_model="dom", **kw)[3]      # FP64 DMMA rotation
# lane and warp kernels take ln|V| differently (table log of 16-sample products vs log per sample): same statistics to ~1e-9
ok = ~np.isnan(b[:, 0])
assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[ok, :2], b[ok, :2], rtol=1e-7, atol=0)
assert np.array_equal(f, f2, equal_nan=True) and np.isfinite(d[:, 1]).any()
print("sanitize_scan ok", a.shape, f.shape, d.shape)
