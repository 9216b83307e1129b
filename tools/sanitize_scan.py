"""Small packed scans through every kernel family, for compute-sanitizer: tcgen05 int8 rotation + lane solve (forced),
warp solve, fixed-lambda (lane and warp), FP64 DMMA path (dominance coding), GRM + eigh."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import make_problem  # noqa: E402

from janusx_b200 import assoc, jxrs  # noqa: E402

case = make_problem(n=300, m=700, q=2, seed=3, missing_rate=0.03)
n = case.n
g = jxrs.DeviceGrm(n, None, 1, 0)
g.update(case.packed, None, qc=(0.02, 0.05, 1.0))
K, _ = g.finish()
g.close()
m = assoc.LMM(case.y, case.cov, K, device=0)
mdl = m.device_model
l10 = float(np.log10(m.lbd_null))
kw = dict(low=float(m.bounds[0]), high=float(m.bounds[1]))
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
assert np.array_equal(a, b, equal_nan=True) and np.array_equal(f, f2, equal_nan=True) and np.isfinite(d[:, 1]).any()
print("sanitize_scan ok", a.shape, f.shape, d.shape)
