"""Run-to-run bit reproducibility of the null-model front steps (GRM, eigendecomposition, X/y rotation, null fit) and of
one scan batch: the multi-GPU CLI test compares TSVs of separate processes byte for byte."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from conftest import make_problem  # noqa: E402

from janusx_b200 import assoc, jxrs  # noqa: E402


def main():
    case = make_problem(n=400, m=3000, q=0, seed=123, missing_rate=0.02)
    n = case.n
    ks, ws, us = [], [], []
    for rep in range(3):
        g = jxrs.DeviceGrm(n, None, 1, 0)
        g.update(case.packed, None, qc=(0.02, 0.05, 1.0))
        k, _ = g.finish()
        g.close()
        ks.append(k)
        w, u = assoc._eigh(k + 1e-6 * np.eye(n), 0)
        ws.append(w)
        us.append(u)
    print("GRM reproducible:", all(np.array_equal(ks[0], k) for k in ks[1:]))
    print("eigh values reproducible:", all(np.array_equal(ws[0], w) for w in ws[1:]),
          "vectors:", all(np.array_equal(us[0], u) for u in us[1:]),
          "max |du|:", max(float(np.abs(us[0] - u).max()) for u in us[1:]))
    # same matrix, same process, eigh twice
    w1, u1 = assoc._eigh(ks[0] + 1e-6 * np.eye(n), 0)
    w2, u2 = assoc._eigh(ks[0] + 1e-6 * np.eye(n), 0)
    print("eigh(same K) reproducible:", np.array_equal(w1, w2), np.array_equal(u1, u2))
    outs = []
    for rep in range(2):
        m = assoc.LMM(case.y, None, ks[0], device=0)
        l10 = float(np.log10(m.lbd_null))
        keep, af, miss, out = m.device_model.scan_packed(case.packed, n, low=m.bounds[0], high=m.bounds[1], init=l10)
        outs.append((m.Xcov.copy(), m.y.copy(), m.lbd_null, out))
        m.device_model.close()
    print("null model reproducible:", np.array_equal(outs[0][0], outs[1][0]), np.array_equal(outs[0][1], outs[1][1]),
          outs[0][2] == outs[1][2], "scan rows:", np.array_equal(outs[0][3], outs[1][3], equal_nan=True))
    # shard dependence: the same rows scanned as one batch and as two halves
    m = assoc.LMM(case.y, None, ks[0], device=0)
    l10 = float(np.log10(m.lbd_null))
    kw = dict(low=m.bounds[0], high=m.bounds[1], init=l10)
    _, _, _, full = m.device_model.scan_packed(case.packed, n, **kw)
    _, _, _, a = m.device_model.scan_packed(case.packed[:1500], n, **kw)
    _, _, _, b = m.device_model.scan_packed(case.packed[1500:], n, **kw)
    ab = np.concatenate([a, b])
    print("batch split independent:", np.array_equal(full, ab, equal_nan=True),
          "rows differing:", int(np.sum(np.any((full != ab) & ~(np.isnan(full) & np.isnan(ab)), axis=1))))


if __name__ == "__main__":
    main()
