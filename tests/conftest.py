import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_small():
    return dict(np.load(GOLDEN / "small_n96.npz", allow_pickle=False))


def make_problem(n=96, m=40, q=2, seed=11, missing_rate=0.03):
    """A small synthetic case shared by CPU and GPU tests (deterministic; host numpy only)."""
    from janusx_b200 import synth
    return synth.make_case(n=n, m=m, q=q, seed=seed, missing_rate=missing_rate)


def null_model(O, case):
    """Reference null-model bookkeeping (pyBLUP/assoc.py:1818-1876) computed with the oracle."""
    n = case.n
    ut = np.ascontiguousarray(case.u.T.astype(np.float32))
    X = np.concatenate([np.ones((n, 1)), case.cov], axis=1)
    xr, yr = O.lmm_rotate_x_y_with_ut_f64(ut, X, case.y)
    lbd, ml0, reml0 = O.lmm_reml_null_f32(case.s, xr, yr[:, 0], -5.0, 5.0, 50, 1e-3)
    lo, hi = float(np.log10(lbd) - 2.0), float(np.log10(lbd) + 2.0)
    return dict(ut=ut, X=X, xcov=xr, y=yr[:, 0].copy(), lbd=lbd, ml0=ml0, reml0=reml0, low=lo, high=hi)
