"""Does a scan row depend on the batch it is scanned in?  Exact CLI null model (2 PCs), full batch vs halves, repeated."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from conftest import make_problem  # noqa: E402

from janusx_b200 import _cabi, assoc, jxrs  # noqa: E402


def fetch_rot(mdl, rows):
    ldc = (mdl.n + 31) // 32 * 32
    out = np.zeros((rows, ldc), dtype=np.float32)
    _cabi.check(_cabi.lib().jxb_debug_fetch_rot(mdl.handle, 0, rows, _cabi.ptr(out)))
    return out


def main():
    case = make_problem(n=400, m=3000, q=0, seed=123, missing_rate=0.02)
    n = case.n
    g = jxrs.DeviceGrm(n, None, 1, 0)
    g.update(case.packed, None, qc=(0.02, 0.05, 1.0))
    K, _ = g.finish()
    g.close()
    evals, evecs = assoc._eigh(K + 1e-6 * np.eye(n), 0)
    X = evecs[:, ::-1][:, :2] * np.sqrt(np.maximum(evals[::-1][:2], 0.0))
    base = assoc.LMM(case.y, X, K, device=0)
    mdl = base.device_model
    l10 = float(np.log10(base.lbd_null))
    kw = dict(low=float(base.bounds[0]), high=float(base.bounds[1]), init=l10)
    for rep in range(4):
        keep_f, _, _, full = mdl.scan_packed(case.packed, n, **kw)
        rot_full = fetch_rot(mdl, full.shape[0])
        ka, _, _, a = mdl.scan_packed(case.packed[:1500], n, **kw)
        rot_a = fetch_rot(mdl, a.shape[0])
        kb, _, _, b = mdl.scan_packed(case.packed[1500:], n, **kw)
        rot_b = fetch_rot(mdl, b.shape[0])
        ab = np.concatenate([a, b])
        rot_ab = np.concatenate([rot_a, rot_b])
        bad = np.nonzero(np.any((full != ab) & ~(np.isnan(full) & np.isnan(ab)), axis=1))[0]
        rbad = np.nonzero(np.any(rot_full != rot_ab, axis=1))[0]
        print(f"rep {rep}: result rows differing {bad.tolist()[:8]} rot rows differing {rbad.tolist()[:8]} "
              f"padding nonzero {int(np.count_nonzero(rot_full[:, n:]))}")
        for r in bad[:3]:
            print("   full", full[r], "split", ab[r])
            d = np.nonzero(rot_full[r] != rot_ab[r])[0]
            print("   rot cols differing", d.tolist()[:8], rot_full[r, d[:4]], rot_ab[r, d[:4]])
        # same batch twice
        _, _, _, full2 = mdl.scan_packed(case.packed, n, **kw)
        print("   same batch again identical:", np.array_equal(full, full2, equal_nan=True))


if __name__ == "__main__":
    main()
