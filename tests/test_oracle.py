"""CPU tests of the oracle: reference-held known answers, independent cross-checks, committed goldens.

The reference's own tests hold no beta/se/p/lambda values for this path, so what can be
pinned against the reference is pinned here (decode codes, value LUT, chi-square tails, p-value
sanitising, chunked == unchunked); everything else is cross-checked against independent numpy algebra.
"""
import math

import numpy as np
import pytest

from conftest import make_problem, null_model


def test_decode_kat_codes(oracle):
    # src/stats/packed.rs:2613-2643: byte 0x03 -> first sample 2.0 ; byte 0x02 -> 1.0 (before centring)
    # src/math/bedmath.rs:1536-1573: value LUT by code [00,01,10,11] = [0, 2*maf, 1, 2]
    packed = np.array([[0b11100100]], dtype=np.uint8)  # samples: code 00, 01, 10, 11
    maf = np.array([0.25], dtype=np.float32)
    g = oracle.decode_centered_block(packed, 4, maf)
    raw = np.array([0.0, 0.5, 1.0, 2.0], dtype=np.float32)
    mean = np.float32(raw.astype(np.float64).sum() / 4.0)
    assert np.array_equal(g[0], raw - mean)
    for byte, want in ((0x03, 2.0), (0x02, 1.0)):
        gg = oracle.decode_centered_block(np.array([[byte]], dtype=np.uint8), 1, np.array([0.5], dtype=np.float32))
        # one sample: centred value is 0, so check through a 2-sample row with a known second sample (code 00)
        gg = oracle.decode_centered_block(np.array([[byte]], dtype=np.uint8), 2, np.array([0.5], dtype=np.float32))
        assert gg[0, 0] - gg[0, 1] == np.float32(want)


def test_counts_and_qc_rules(oracle):
    n = 10
    codes = np.array([[0, 1, 2, 3, 3, 2, 0, 0, 1, 0]], dtype=np.uint8)  # raw 2-bit codes
    packed = np.zeros((1, 3), dtype=np.uint8)
    for j in range(n):
        packed[0, j >> 2] |= codes[0, j] << ((j & 3) * 2)
    keep, af, mr, missing = oracle.count_qc_block(packed, n, None, 0.0, 1.0, 1.0)
    assert missing[0] == 2 and keep[0]
    assert mr[0] == np.float32(2) / np.float32(10)
    # het=2, hom_alt=2 -> alt_sum 6 over 2*8
    assert af[0] == np.float32(6) / (np.float32(2.0) * np.float32(8))
    # miss threshold is strict '>' (lmm.rs:1284)
    assert oracle.count_qc_block(packed, n, None, 0.0, 0.2, 1.0)[0][0]
    assert not oracle.count_qc_block(packed, n, None, 0.0, 0.19, 1.0)[0][0]
    # maf filter uses the folded frequency but the stored af stays unfolded (lmm.rs:1311-1321)
    assert not oracle.count_qc_block(packed, n, None, 0.4, 1.0, 1.0)[0][0]
    # selected samples: drop both missing calls
    sel = np.array([0, 2, 3, 4, 5, 6, 7, 9], dtype=np.int64)
    k2, af2, mr2, ms2 = oracle.count_qc_block(packed, n, sel, 0.0, 1.0, 1.0)
    assert ms2[0] == 0 and mr2[0] == 0.0 and af2[0] == np.float32(6) / (np.float32(2.0) * np.float32(8))
    # all-missing row: kept only when maf_thr <= 0 (lmm.rs:1288-1299)
    allmiss = np.full((1, 3), 0b01010101, dtype=np.uint8)
    assert oracle.count_qc_block(allmiss, n, None, 0.0, 1.0, 1.0)[0][0]
    assert not oracle.count_qc_block(allmiss, n, None, 0.01, 1.0, 1.0)[0][0]


def test_chi2_and_normal_tails(oracle):
    # src/math/linalg.rs:373-377: inverse sf at p=1.8885e-19 is ~81.8 -> forward sf at 81.8 is ~that p
    # (the reference asserts |stat - 81.8| < 0.5, i.e. sf(81.3) > 1.8885e-19 > sf(82.3))
    assert oracle.chi2_sf_df1(81.3) > 1.8885e-19 > oracle.chi2_sf_df1(82.3)
    assert oracle.chi2_sf_df1(0.0) == 1.0 and oracle.chi2_sf_df1(float("nan")) == 1.0
    assert oracle.chi2_sf_df1(-3.0) == 1.0
    # df=1 identity: chi2_sf(z^2) == 2*normal_sf(|z|)
    for z in (0.3, 1.0, 2.5, 6.0):
        assert math.isclose(oracle.chi2_sf_df1(z * z), 2.0 * oracle.normal_sf(z), rel_tol=1e-14)
    assert oracle.chi2_sf_df1(1e6) == 2.2250738585072014e-308  # clamp to f64::MIN_POSITIVE


def test_tsv_row_format_and_sanitize(oracle):
    # src/math/linalg.rs:398-405 sanitize rules + Rust float formatting
    row = oracle.format_row("1", 123, ".", "A", "T", 0.31234, 0.0125, [float("nan"), 1.0, 1e-12]).decode()
    f = row.rstrip("\n").split("\t")
    assert f[:5] == ["1", "123", "1_123", "A", "T"]
    assert f[5:] == ["0.3123", "0.0125", "NaN", "1.0000", "NaN", "1.0000e0"]
    row = oracle.format_row("2", 5, "rs1", "G", "C", 0.5, 0.0, [1.0, 0.0, 1e-12]).decode().split("\t")
    assert row[9] == "NaN" and row[10].strip() == "1.0000e0"       # se == 0 -> invalid
    row = oracle.format_row("2", 5, "rs1", "G", "C", 0.5, 0.0, [0.5, 0.25, 6.1e-5, 2.5, -1234.5678, 3e-300])
    f = row.decode().rstrip("\n").split("\t")
    assert f[7:] == ["0.5000", "0.2500", "4.0000e0", "6.1000e-5", "2.500000e0", "-1.234568e3", "3.0000e-300"]
    assert oracle.fmt_exp(0.0, 4) == "0.0000e0" and oracle.fmt_exp(float("inf"), 4) == "inf"
    assert oracle.fmt_fixed(-0.0, 4) == "-0.0000" and oracle.fmt_fixed(float("nan"), 4) == "NaN"


def test_brent_restatement(oracle):
    # quadratic: converges to the vertex; the reference's loose tolerance semantics (tol*|x| + eps)
    x, fx, ne = oracle.brent_minimize(lambda t: (t - 1.3) ** 2 + 2.0, -5.0, 5.0, 1e-2, 30)
    assert abs(x - 1.3) < 2e-2 and ne <= 31
    # init outside the bracket falls back to the midpoint (brent.rs:36-38)
    x0, _, _ = oracle.brent_minimize(lambda t: (t - 1.3) ** 2, -5.0, 5.0, 1e-2, 0, init_x=9.0)
    assert x0 == 0.0
    x1, _, _ = oracle.brent_minimize(lambda t: (t - 1.3) ** 2, -5.0, 5.0, 1e-2, 0, init_x=2.0)
    assert x1 == 2.0
    # near x = 0 the tolerance collapses to eps and the search runs to max_iter (SURVEY finding 3)
    _, _, ne0 = oracle.brent_minimize(lambda t: t * t, -5.0, 4.0, 1e-2, 30)
    assert ne0 == 31


def test_objective_against_dense_algebra(oracle):
    case = make_problem(n=80, m=12, q=2, seed=5, missing_rate=0.0)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, _, _ = oracle.count_qc_block(case.packed, n, None, 0.0, 1.0, 1.0)
    g = oracle.decode_centered_block(case.packed, n, af)
    rot = oracle.rotate_block(g, nm["ut"]).astype(np.float64)
    x = -0.3
    lam = 10.0 ** x
    for j in range(3):
        Z = np.concatenate([nm["xcov"], rot[j][:, None]], axis=1)
        w = 1.0 / (case.s + lam)
        A = Z.T @ (w[:, None] * Z) + 1e-6 * np.eye(Z.shape[1])
        b = Z.T @ (w * nm["y"])
        beta = np.linalg.solve(A, b)
        r = nm["y"] - Z @ beta
        Q = float((w * r * r).sum())
        d = Z.shape[1]
        reml = ((n - d) * (math.log(n - d) - 1 - math.log(2 * math.pi)) / 2
                - 0.5 * ((n - d) * math.log(Q) + np.log(case.s + lam).sum() + np.linalg.slogdet(A)[1]))
        ml = n * (math.log(n) - 1 - math.log(2 * math.pi)) / 2 - 0.5 * (n * math.log(Q) + np.log(case.s + lam).sum())
        assert math.isclose(oracle.reml_loglike(x, case.s, nm["xcov"], nm["y"], rot[j]), reml, rel_tol=1e-12)
        assert math.isclose(oracle.ml_loglike(x, case.s, nm["xcov"], nm["y"], rot[j]), ml, rel_tol=1e-12)
        bk, se, lo = oracle.final_beta_se(x, case.s, nm["xcov"], nm["y"], rot[j])
        assert math.isclose(bk, beta[-1], rel_tol=1e-10)
        assert math.isclose(se, math.sqrt(Q / (n - d) * np.linalg.inv(A)[-1, -1]), rel_tol=1e-10)
        assert lo == lam


def test_rotation_modes_and_xy(oracle):
    case = make_problem(n=64, m=8, q=1, seed=3, missing_rate=0.0)
    nm = null_model(oracle, case)
    keep, af, _, _ = oracle.count_qc_block(case.packed, case.n, None, 0.0, 1.0, 1.0)
    g = oracle.decode_centered_block(case.packed, case.n, af)
    r0 = oracle.rotate_block(g, nm["ut"], mode=0)
    exact = (g.astype(np.float64) @ nm["ut"].astype(np.float64).T).astype(np.float32)
    assert np.max(np.abs(r0.astype(np.float64) - exact)) <= np.spacing(np.abs(exact).max())
    r1 = oracle.rotate_block(g, nm["ut"], mode=1)
    assert np.allclose(r0, r1, rtol=0, atol=1e-4)   # the reference's own Metal-projector tolerance
    xr = nm["ut"].astype(np.float64) @ nm["X"]
    assert np.allclose(nm["xcov"], xr, rtol=1e-12, atol=1e-12)


def test_invalid_rows_and_chunk_invariance(oracle):
    # python/janusx/assoc/smoke.py:38-46: chunked == unchunked; lmm.rs:121-125: zero-variance SNP -> NaN,NaN,1
    case = make_problem(n=72, m=16, q=2, seed=9, missing_rate=0.01)
    nm = null_model(oracle, case)
    keep, af, _, _ = oracle.count_qc_block(case.packed, case.n, None, 0.0, 1.0, 1.0)
    g = oracle.decode_centered_block(case.packed, case.n, af)
    g[3] = 0.0
    full = oracle.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30, 1e-2)
    assert np.isnan(full[3, 0]) and np.isnan(full[3, 1]) and full[3, 2] == 1.0
    parts = [oracle.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g[i:i + 5],
                                                nm["ut"], 30, 1e-2, threads=2) for i in range(0, 16, 5)]
    assert np.array_equal(np.concatenate(parts), full, equal_nan=True)
    with pytest.raises(RuntimeError, match="low must be < high"):
        oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], 1.0, 1.0, g)
    with pytest.raises(RuntimeError, match=r"g_rot_chunk must be \(m_chunk, n\)"):
        oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], -1.0, 1.0, g[:, :-1])


def test_goldens_reproduce(oracle, golden_small):
    G = golden_small
    n = int(G["n"])
    keep, af, mr, missing = oracle.count_qc_block(G["packed"], n, None, 0.02, 0.05, 1.0)
    assert np.array_equal(keep, G["keep"]) and np.array_equal(af, G["af"]) and np.array_equal(missing, G["missing"])
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(G["packed"], n, af[idx], row_indices=idx)
    assert np.array_equal(g, G["g"])
    ut = np.ascontiguousarray(G["u"].T.astype(np.float32))
    rot = oracle.rotate_block(g, ut)
    assert np.array_equal(rot, G["rot"])
    lo, hi = G["bounds"]
    lmm = oracle.lmm_reml_chunk_f32(G["s"], G["xcov"], G["y"], lo, hi, rot, 30, 1e-2)
    assert np.allclose(lmm, G["lmm"], rtol=1e-12, atol=0, equal_nan=True)
    lmm2 = oracle.lmm_reml_lmm2_chunk_f32(G["s"], G["xcov"], G["y"], lo, hi, rot, float(G["ml_null"][1]), 30, 1e-2)
    assert np.allclose(lmm2, G["lmm2"], rtol=1e-12, atol=0, equal_nan=True)
    # 15-19 objective evaluations per SNP at tol=1e-2, max_iter=30 (SURVEY 3.4 probe)
    assert 8 <= G["lmm_evals"].min() and G["lmm_evals"].max() <= 32


def test_toy_recipe_golden(oracle):
    import numpy as np
    from conftest import GOLDEN
    T = dict(np.load(GOLDEN / "toy_n8.npz"))
    ut = np.ascontiguousarray(T["u"].T.astype(np.float32))
    snp = np.ascontiguousarray(T["G"].T.astype(np.float32))
    res = oracle.lmm_reml_chunk_from_snp_f32(T["s"], T["xcov"], T["yrot"], -5.0, 5.0, snp, ut, 30, 1e-2)
    assert res.shape == (5, 3)                                 # smoke.py:41
    assert np.allclose(res, T["res_wide"], rtol=1e-12, equal_nan=True)
    chunked = np.concatenate([oracle.lmm_reml_chunk_from_snp_f32(T["s"], T["xcov"], T["yrot"], -5.0, 5.0,
                                                                 snp[i:i + 2], ut, 30, 1e-2, threads=2)
                              for i in range(0, 5, 2)])
    assert np.allclose(res, chunked, equal_nan=True)            # smoke.py:45-46


def test_bed_scan_composition(oracle, tmp_path):
    from janusx_b200 import synth
    case = make_problem(n=60, m=30, q=1, seed=21, missing_rate=0.04)
    nm = null_model(oracle, case)
    prefix = str(tmp_path / "toy")
    ids = [f"snp{i}" if i % 7 else "." for i in range(30)]
    synth.write_plink(prefix, case.packed, case.n, snp_ids=ids)
    out = tmp_path / "o.tsv"
    rows = oracle.scan_bed_to_tsv(prefix, str(out), case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05, 1.0,
                                  low=nm["low"], high=nm["high"], rotate_block_rows=8)
    lines = out.read_bytes().split(b"\n")
    assert lines[0] == oracle.HEADERS[3].rstrip(b"\n")
    assert len(lines) - 2 == rows and rows > 0
    assert any(l.split(b"\t")[2].startswith(b"1_") for l in lines[1:-1])   # '.' ids become chrom_pos
    # block size must not change the output (no warm start)
    out2 = tmp_path / "o2.tsv"
    oracle.scan_bed_to_tsv(prefix, str(out2), case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05, 1.0,
                           low=nm["low"], high=nm["high"], rotate_block_rows=512)
    assert out.read_bytes() == out2.read_bytes()


def test_route_b_row_decisions_match_restatement(oracle):
    """BedChunkReader.next_chunk_prepared (route B): the host-side f64 QC of janusx_b200/gfreader.py on integer counts
    against the row-by-row restatement of src/io/gfcore.rs:405-480 + src/io/gfreader.rs:3660-3681."""
    from janusx_b200 import synth
    from janusx_b200.gfreader import prepared_row_decisions
    n_full, m = 123, 400
    packed, _ = synth.draw_genotypes(m, n_full, seed=9, missing_rate=0.08)
    packed[3] = 0b01010101        # all missing
    packed[4] = 0                 # monomorphic
    sidx = np.array(sorted(np.random.default_rng(0).choice(n_full, size=77, replace=False)), dtype=np.int64)
    for idx in (None, sidx):
        sel = np.arange(n_full) if idx is None else idx
        codes = (packed[:, sel >> 2] >> (2 * (sel & 3)).astype(np.uint8)) & 3
        het, hom, mis = (codes == 2).sum(1), (codes == 3).sum(1), (codes == 1).sum(1)
        for thr in ((0.0, 1.0, 1.0), (0.05, 0.1, 1.0), (0.02, 0.05, 0.4), (0.1, 0.02, 0.0)):
            keep_o, g, af, miss = oracle.bed_chunk_prepared_rows(packed, n_full, idx, *thr)
            keep, imputed = prepared_row_decisions(mis, het, hom, sel.shape[0], *thr)
            assert np.array_equal(keep, keep_o), thr
            k = np.nonzero(keep)[0]
            total = (het[k] + 2 * hom[k]).astype(np.float64) + mis[k].astype(np.float64) * imputed[k].astype(np.float64)
            coded_mean = (total / float(sel.shape[0])).astype(np.float32)
            assert np.array_equal((coded_mean * np.float32(0.5)).view(np.uint32), af.view(np.uint32))
            assert np.array_equal(miss, mis[k].astype(np.float32))
            # the restated rows are centred: f32 row sums vanish to rounding
            if k.size:
                assert np.abs(g.astype(np.float64).sum(axis=1)).max() < 1e-3
