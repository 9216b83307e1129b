"""GPU parity for the two steps in front of the scan: N1 the centred additive GRM on the int8 tensor
cores (csrc/grm.cu) and N2 its eigendecomposition (csrc/eigh.cu, cuSOLVER).

GRM tolerances: against the numpy restatement with mu on the kernel's 2^-21 grid the contraction is exact integer
work, so only f64 accumulation order differs (1e-13 of max|K|); against the restatement with the reference's f32
genotype values the grid rounding shows up at ~3e-8 of max|K| (the reference's own f32 SYRK is order-dependent at
~1e-6, so 2e-7 is inside its noise).
"""
import numpy as np
import pytest

from conftest import make_problem, null_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jx():
    from janusx_b200 import jxrs
    yield jxrs
    jxrs.clear_model_cache()


def _af(oracle, packed, n_full, sidx=None):
    keep, af, mr, missing = oracle.count_qc_block(packed, n_full, sidx, 0.0, 1.0, 0.0)
    return af


@pytest.mark.parametrize("n,m,miss", [(300, 2500, 0.02), (257, 700, 0.0), (64, 130, 0.2), (1000, 1200, 0.01)])
def test_grm_matches_oracle(jx, oracle, n, m, miss):
    from janusx_b200 import synth
    packed, _ = synth.draw_genotypes(m, n, seed=7 + n, missing_rate=miss)
    packed[0] = 0                                    # monomorphic row: contributes nothing, var = 0
    af = _af(oracle, packed, n)
    flip = np.zeros(m, bool)
    want_grid, varsum = oracle.grm_packed_f64(packed, n, flip, af, mu_grid_bits=21)
    want_f32z, _ = oracle.grm_packed_f64(packed, n, flip, af)
    got = jx.grm_packed_f64(packed, n, flip, af)
    scale = np.abs(want_grid).max()
    assert got.shape == (n, n) and np.array_equal(got, got.T)
    assert np.abs(got - want_grid).max() <= 1e-13 * scale
    assert np.abs(got - want_f32z).max() <= 2e-7 * scale
    got32 = jx.grm_packed_f32(packed, n, flip, af)
    assert got32.dtype == np.float32 and np.array_equal(got32, got.astype(np.float32))
    # streaming in uneven pieces, device-computed allele frequency: same matrix up to f64 accumulation order
    g = jx.DeviceGrm(n)
    for r0, r1 in ((0, 1), (1, 130), (130, m)):
        if r1 > r0:
            g.update(packed[r0:min(r1, m)], None)
    k2, vs2 = g.finish()
    g.close()
    assert vs2 == varsum and np.abs(k2 - got).max() <= 1e-13 * scale


def test_grm_sample_subset_qc_mask_and_errors(jx, oracle):
    from janusx_b200 import synth
    n_full, m = 310, 900
    packed, _ = synth.draw_genotypes(m, n_full, seed=21, missing_rate=0.03)
    sidx = np.array(sorted(np.random.default_rng(3).choice(n_full, size=201, replace=False)), dtype=np.int64)
    af = _af(oracle, packed, n_full, sidx)
    want, _ = oracle.grm_packed_f64(packed, n_full, np.zeros(m, bool), af, sample_indices=sidx, mu_grid_bits=21)
    got = jx.grm_packed_f64(packed, n_full, np.zeros(m, bool), af, sample_indices=sidx)
    assert got.shape == (201, 201) and np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    # QC thresholds on the device == filtering the rows first
    keep, af_q, _, _ = oracle.count_qc_block(packed, n_full, sidx, 0.05, 0.04, 1.0)
    assert 0 < keep.sum() < m
    g = jx.DeviceGrm(n_full, sidx)
    g.update(packed, None, qc=(0.05, 0.04, 1.0))
    assert g.rows_used == int(keep.sum())
    k_q, _ = g.finish()
    g.close()
    want_q, _ = oracle.grm_packed_f64(packed[keep], n_full, np.zeros(int(keep.sum()), bool), af_q[keep],
                                      sample_indices=sidx, mu_grid_bits=21)
    assert np.abs(k_q - want_q).max() <= 1e-13 * np.abs(want_q).max()
    with pytest.raises(RuntimeError, match="packed second dimension mismatch"):
        jx.grm_packed_f64(packed[:, :-1], n_full, np.zeros(m, bool), af)
    with pytest.raises(RuntimeError, match="row_maf length mismatch"):
        jx.grm_packed_f64(packed, n_full, np.zeros(m, bool), af[:-1])
    with pytest.raises(NotImplementedError):
        jx.grm_packed_f64(packed, n_full, np.zeros(m, bool), af, method=2)
    with pytest.raises(RuntimeError, match="invalid centered GRM denominator"):
        jx.grm_packed_f64(np.zeros((4, (n_full + 3) // 4), np.uint8), n_full, np.zeros(4, bool), np.zeros(4, np.float32))


def test_eigh_matches_lapack_properties(jx, oracle):
    rng = np.random.default_rng(5)
    for n in (1, 2, 97, 400):
        a = rng.normal(size=(n, n))
        a = a @ a.T / n + 1e-6 * np.eye(n)
        res = jx.rust_eigh_from_array_f64(a)
        w, v = res[0], res[1]
        assert len(res) == 10 and res[4] == n and res[3] == "cusolver_xsyevd"
        w_np = np.linalg.eigvalsh(a)
        assert np.all(np.diff(w) >= 0) and np.abs(w - w_np).max() <= 1e-12 * max(1.0, np.abs(w_np).max())
        assert np.abs(v.T @ v - np.eye(n)).max() <= 1e-12
        assert np.abs((v * w) @ v.T - a).max() <= 1e-12 * max(1.0, np.abs(a).max())
    # the reference's own known-answer test: src/math/eigh.rs:1981-1998 ([[2,1],[1,2]] -> 1, 3)
    w, v = jx.rust_eigh_from_array_f64(np.array([[2.0, 1.0], [1.0, 2.0]]))[:2]
    assert np.allclose(w, [1.0, 3.0], atol=1e-12) and np.allclose(np.abs(v), np.sqrt(0.5), atol=1e-12)
    assert jx.rust_eigh_from_array_f64(np.eye(3), jobz="N")[1] is None
    with pytest.raises(RuntimeError, match="non-empty square matrix"):
        jx.rust_eigh_from_array_f64(np.zeros((2, 3)))


def test_eigh_cusolvermg_path(jx, monkeypatch):
    """The eigensolver behind n > 46,340 (BASELINE configs[3]) and behind jxb_set_eigh_devices(k > 1): cusolverMgSyevd on
    block-cyclic column panels.  Exercised here at small n on one device (forced) and, when the box has them, on two
    devices: eigenvalues and the reconstruction match LAPACK like the single-call path."""
    import torch
    from janusx_b200 import _cabi
    rng = np.random.default_rng(9)
    n = 777                                             # not a multiple of the 256-column panel
    a = rng.normal(size=(n, n))
    a = a @ a.T / n + 1e-6 * np.eye(n)
    w_np = np.linalg.eigvalsh(a)
    monkeypatch.setenv("JXB_EIGH_FORCE_MG", "1")
    try:
        for k in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
            _cabi.lib().jxb_set_eigh_devices(k)
            w, v = jx.rust_eigh_from_array_f64(a)[:2]
            assert np.all(np.diff(w) >= 0) and np.abs(w - w_np).max() <= 1e-12 * np.abs(w_np).max()
            assert np.abs(v.T @ v - np.eye(n)).max() <= 1e-11
            assert np.abs((v * w) @ v.T - a).max() <= 1e-11 * np.abs(a).max()
    finally:
        _cabi.lib().jxb_set_eigh_devices(0)


def test_grm_eigh_scan_pipeline_on_device(jx, oracle):
    """packed rows -> GRM -> eigh (chained on the device) -> null model -> scan, against the oracle run on the same
    spectral decomposition."""
    case = make_problem(n=280, m=500, q=2, seed=31, missing_rate=0.02)
    n = case.n
    g = jx.DeviceGrm(n)
    g.update(case.packed, None, qc=(0.02, 0.05, 1.0))
    s, ut32 = g.eigh(1e-6)
    g.close()
    keep, af, _, _ = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    K, _ = oracle.grm_packed_f64(case.packed[keep], n, np.zeros(int(keep.sum()), bool), af[keep], mu_grid_bits=21)
    w_np = np.linalg.eigvalsh(K + 1e-6 * np.eye(n))
    assert np.abs(s - w_np).max() <= 1e-10 * np.abs(w_np).max()
    u = ut32.astype(np.float64)
    assert np.abs(u @ u.T - np.eye(n)).max() <= 5e-6          # f32-rounded eigenvectors
    assert np.abs((u.T * s) @ u - (K + 1e-6 * np.eye(n))).max() <= 5e-6 * np.abs(K).max()
    # scan with this decomposition == oracle with the same (S, U^T)
    import copy
    c2 = copy.copy(case)
    c2.s, c2.u = s, np.ascontiguousarray(ut32.T.astype(np.float64))
    nm = null_model(oracle, c2)
    mdl = jx.DeviceModel(s, nm["xcov"], nm["y"], ut32)
    k_d, af_d, miss_d, out_d = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"])
    idx = np.nonzero(keep)[0]
    gdec = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    want = oracle.lmm_reml_chunk_from_snp_f32(s, nm["xcov"], nm["y"], nm["low"], nm["high"], gdec, ut32, 30, 1e-2)
    assert np.array_equal(k_d, keep)
    ok = ~np.isnan(want[:, 0])
    np.testing.assert_allclose(out_d[ok, :2], want[ok, :2], rtol=1e-8)
    assert np.max(np.abs(np.log10(out_d[:, 2]) - np.log10(want[:, 2]))) <= 1e-6
