"""GPU parity tests: every call goes through the C ABI (libjxb200.so) and is compared with the CPU oracle.

Tolerances are the north-star gates: counts / af / SNP order bit-exact; lambda 1e-6
relative; beta, se 1e-8 relative; |delta(-log10 p)| <= 1e-6.  The rotated block is f32: the device sums
in a different (tiled) order than the oracle's sequential-j loop, so an entry may land on the other side
of an f32 rounding boundary -- allowed up to 1 ulp on a tiny fraction of entries.
"""
import math

import numpy as np
import pytest

from conftest import make_problem, null_model

pytestmark = pytest.mark.gpu

RTOL_BETA = 1e-8
RTOL_LAMBDA = 1e-6
ATOL_LOGP = 1e-6


@pytest.fixture(scope="module")
def jx():
    from janusx_b200 import jxrs
    yield jxrs
    jxrs.clear_model_cache()


def assert_results_close(got, want, cols_p=(2,), cols_lambda=(), cols_rel=(0, 1)):
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want[:, 0])
    for c in cols_rel:
        np.testing.assert_allclose(got[ok, c], want[ok, c], rtol=RTOL_BETA, atol=0)
    for c in cols_lambda:
        np.testing.assert_allclose(got[ok, c], want[ok, c], rtol=RTOL_LAMBDA, atol=0)
    for c in cols_p:
        assert np.max(np.abs(np.log10(got[:, c]) - np.log10(want[:, c]))) <= ATOL_LOGP


def assert_rot_close(got, want):
    """f32-rounded f64 dot products: equal except where the two summation orders straddle an f32 rounding
    boundary (<= 1 ulp), or where the entry is ~0 and f64 summation noise (~1e-13 of the row scale) exceeds
    its ulp (centred genotypes projected on the near-constant eigenvector)."""
    diff = got != want
    frac = diff.mean()
    if frac > 0:
        ulp = np.spacing(np.abs(want).astype(np.float32))
        noise = 1e-13 * np.abs(want).max(axis=1, keepdims=True)
        assert np.all(np.abs(got - want) <= ulp + noise), "beyond 1 f32 ulp + f64 summation noise"
    big = np.abs(want) > 1e-6 * np.abs(want).max()
    assert (diff & big).mean() < 1e-4, f"{(diff & big).mean():.2e} of rotated entries differ"


def test_k1_counts_qc_decode_bit_exact(jx, oracle, golden_small):
    G = golden_small
    n = int(G["n"])
    mdl = jx.DeviceModel(G["s"], G["xcov"], G["y"], np.ascontiguousarray(G["u"].T.astype(np.float32)))
    counts, af, mr, g = mdl.decode_packed(G["packed"], n, None, 0.02, 0.05, 1.0)
    assert np.array_equal(counts[:, 3].astype(bool), G["keep"])
    assert np.array_equal(counts[:, 0], G["missing"])
    assert np.array_equal(af.view(np.uint32), G["af"].view(np.uint32))
    assert np.array_equal(mr.view(np.uint32), G["miss_rate"].view(np.uint32))
    assert np.array_equal(g.view(np.uint32), G["g"].view(np.uint32))
    # sample subset (gather path), golden + freshly computed oracle
    sub = jx.DeviceModel(G["s"][: len(G["sidx"])], G["xcov"][: len(G["sidx"])], G["y"][: len(G["sidx"])])
    c2, af2, _, g2 = sub.decode_packed(G["packed"], n, G["sidx"], 0.02, 0.05, 1.0)
    assert np.array_equal(c2[:, 3].astype(bool), G["keep_s"]) and np.array_equal(c2[:, 0], G["missing_s"])
    assert np.array_equal(af2.view(np.uint32), G["af_s"].view(np.uint32))
    assert np.array_equal(g2.view(np.uint32), G["g_s"].view(np.uint32))


@pytest.mark.parametrize("n,m,miss", [(1, 3, 0.0), (5, 7, 0.3), (33, 64, 0.1), (257, 40, 0.02), (1000, 300, 0.01)])
def test_k1_ragged_shapes_and_models(jx, oracle, n, m, miss):
    from janusx_b200 import synth
    packed, _ = synth.draw_genotypes(m, n, seed=100 + n, missing_rate=miss)
    packed[0] = 0  # monomorphic row
    if m > 2:
        packed[1] = 0b01010101  # all missing (padding bits included: must be masked)
    mdl = jx.DeviceModel(np.ones(n), np.ones((n, 1)), np.zeros(n))
    for model, thr in (("add", (0.0, 1.0, 1.0)), ("dom", (0.02, 0.5, 0.9)), ("rec", (0.0, 1.0, 0.0)), ("het", (0.1, 0.2, 1.0))):
        keep, af, mr, missing = oracle.count_qc_block(packed, n, None, *thr)
        counts, af_d, mr_d, g_d = mdl.decode_packed(packed, n, None, *thr, genetic_model=model)
        assert np.array_equal(counts[:, 3].astype(bool), keep)
        assert np.array_equal(counts[:, 0], missing)
        assert np.array_equal(af_d.view(np.uint32), af.view(np.uint32))
        assert np.array_equal(mr_d.view(np.uint32), mr.view(np.uint32))
        idx = np.nonzero(keep)[0]
        g = oracle.decode_centered_block(packed, n, af[idx], row_indices=idx, model=model)
        assert g_d.shape == g.shape
        assert np.array_equal(g_d.view(np.uint32), g.view(np.uint32))


def test_k2_rotation_both_kernels(jx, oracle, golden_small):
    G = golden_small
    ut = np.ascontiguousarray(G["u"].T.astype(np.float32))
    mdl = jx.DeviceModel(G["s"], G["xcov"], G["y"], ut)
    for variant in (1, 0):
        rot = mdl.rotate_block(G["g"], variant=variant)
        assert_rot_close(rot, G["rot"])


@pytest.mark.parametrize("n,rows", [(130, 5), (257, 129), (1000, 300)])
def test_k2_rotation_edges(jx, oracle, n, rows):
    rng = np.random.default_rng(n)
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ut = np.ascontiguousarray(q.T.astype(np.float32))
    g = rng.normal(size=(rows, n)).astype(np.float32)
    want = oracle.rotate_block(g, ut, mode=0)
    mdl = jx.DeviceModel(np.ones(n), np.ones((n, 1)), np.zeros(n), ut)
    assert_rot_close(mdl.rotate_block(g, variant=0), want)
    assert_rot_close(mdl.rotate_block(g, variant=1), want)
    # linearity (size-independent property): rot(a*g1 + g2) == a*rot(g1) + rot(g2) up to f32 rounding
    g2 = rng.normal(size=(rows, n)).astype(np.float32)
    lhs = mdl.rotate_block((2.0 * g + g2).astype(np.float32))
    rhs = 2.0 * mdl.rotate_block(g) + mdl.rotate_block(g2)
    assert np.allclose(lhs, rhs, rtol=0, atol=2e-5 * np.sqrt(n))


def test_rotate_xy_and_null_fit(jx, oracle, golden_small):
    G = golden_small
    n = int(G["n"])
    ut = np.ascontiguousarray(G["u"].T.astype(np.float32))
    X = np.concatenate([np.ones((n, 1)), G["cov"]], axis=1)
    xr, yr = jx.lmm_rotate_x_y_with_ut_f64(ut, X, G["y_raw"])
    np.testing.assert_allclose(xr, G["xcov"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(yr[:, 0], G["y"], rtol=1e-12, atol=1e-13)
    lbd, ml0, reml0 = jx.lmm_reml_null_f32(G["s"], G["xcov"], G["y"], -5.0, 5.0, 50, 1e-3)
    assert math.isclose(lbd, G["null"][0], rel_tol=RTOL_LAMBDA)
    assert math.isclose(ml0, G["null"][1], rel_tol=1e-10) and math.isclose(reml0, G["null"][2], rel_tol=1e-10)
    assert math.isclose(jx.ml_loglike_null_f32(G["s"], G["xcov"], G["y"], 0.25),
                        oracle.ml_loglike_null_f32(G["s"], G["xcov"], G["y"], 0.25), rel_tol=1e-12)
    mdl = jx.DeviceModel(G["s"], G["xcov"], G["y"])
    x_ml, ml_null = mdl.ml_null(float(G["bounds"][0]), float(G["bounds"][1]), 30, 1e-2)
    assert math.isclose(x_ml, G["ml_null"][0], rel_tol=1e-6, abs_tol=1e-9)
    assert math.isclose(ml_null, G["ml_null"][1], rel_tol=1e-10)


def test_k3_solve_on_rotated_golden(jx, golden_small):
    G = golden_small
    lo, hi = map(float, G["bounds"])
    out = jx.lmm_reml_chunk_f32(G["s"], G["xcov"], G["y"], lo, hi, G["rot"], 30, 1e-2)
    assert_results_close(out, G["lmm"])
    out4 = jx.lmm_reml_chunk_f32(G["s"], G["xcov"], G["y"], lo, hi, G["rot"], 30, 1e-2, nullml=float(G["null"][1]))
    assert_results_close(out4, G["lmm4"], cols_p=(2, 3))
    mdl = jx.DeviceModel(G["s"], G["xcov"], G["y"])
    _, ev = mdl.lmm_reml_chunk(G["rot"], lo, hi, 30, 1e-2, return_evals=True)
    assert np.array_equal(ev, G["lmm_evals"])   # same Brent path, evaluation for evaluation
    out2 = mdl.lmm2_chunk(G["rot"], lo, hi, float(G["ml_null"][1]), 30, 1e-2, rotated=True)
    assert_results_close(out2, G["lmm2"], cols_p=(2, 5), cols_lambda=(3,))
    np.testing.assert_allclose(out2[:, 4], G["lmm2"][:, 4], rtol=1e-10)
    fx, meta = mdl.fixed_chunk(G["rot"], float(np.log10(G["null"][0])), rotated=True, return_meta=True)
    assert_results_close(fx, G["fixed"])
    np.testing.assert_allclose([meta["ypy"], meta["log_det_v"], meta["df"]], G["fixed_meta"], rtol=1e-10)


@pytest.mark.parametrize("p_cov", [1, 3, 5, 8, 11])
def test_k3_covariate_counts(jx, oracle, p_cov):
    # static register kernels for p<=8, runtime-p kernel above
    case = make_problem(n=150, m=24, q=p_cov - 1, seed=40 + p_cov, missing_rate=0.01)
    nm = null_model(oracle, case)
    keep, af, _, _ = oracle.count_qc_block(case.packed, case.n, None, 0.0, 1.0, 1.0)
    g = oracle.decode_centered_block(case.packed, case.n, af)
    g[2] = 0.0   # degenerate SNP -> NaN, NaN, 1 (lmm.rs:121-125)
    rot = oracle.rotate_block(g, nm["ut"])
    want = oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], rot, 30, 1e-2)
    got = jx.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], rot, 30, 1e-2)
    assert_results_close(got, want)
    _, mlnull = oracle.lmm_ml_null_brent(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], 30, 1e-2)
    want2 = oracle.lmm_reml_lmm2_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], rot, mlnull, 30, 1e-2)
    mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"])
    got2 = mdl.lmm2_chunk(rot, nm["low"], nm["high"], mlnull, 30, 1e-2, rotated=True)
    assert_results_close(got2, want2, cols_p=(2, 5), cols_lambda=(3,))


@pytest.mark.parametrize("p_cov", [1, 5, 8])
def test_big_solve_kernels_covariate_counts(jx, oracle, p_cov):
    """Large-batch solve kernels (lane-per-SNP default, thread-per-SNP) for other covariate counts than the bench's 4,
    incl. p >= 5 where they run at 3 CTAs/SM; -lmm2 through the packed scan, evaluation counts included."""
    case = make_problem(n=180, m=96, q=p_cov - 1, seed=60 + p_cov, missing_rate=0.02)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, _, _ = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    _, mlnull = oracle.lmm_ml_null_brent(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], 30, 1e-2)
    want, ev_o = oracle.lmm_reml_lmm2_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"],
                                                oracle.rotate_block(g, nm["ut"]), mlnull, 30, 1e-2, return_evals=True)
    mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
    outs = []
    try:
        jx.set_thread_solve_min_rows(1)
        for variant in (0, 1):
            jx.set_big_solve_kernel(variant)
            k, _, _, out, ev = mdl.scan_packed(case.packed, n, mode="lmm2", low=nm["low"], high=nm["high"], nullml=mlnull,
                                               return_evals=True)
            assert np.array_equal(k, keep) and np.array_equal(ev, ev_o)
            assert_results_close(out, want, cols_p=(2, 5), cols_lambda=(3,))
            outs.append(out)
    finally:
        jx.set_thread_solve_min_rows(32768)
        jx.set_big_solve_kernel(0)
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_end_to_end_from_snp_and_packed(jx, oracle):
    case = make_problem(n=400, m=600, q=3, seed=77, missing_rate=0.02)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, mr, missing = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    want = oracle.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30, 1e-2)
    got = jx.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30, 1e-2)
    assert_results_close(got, want)
    # reference-faithful comparator: f32-accumulated rotation, the reference's own 1e-5 acceptance bound
    ref32 = oracle.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30,
                                               1e-2, rot_mode=1)
    ok = ~np.isnan(ref32[:, 0])
    rel = np.abs(got[ok, :2] - ref32[ok, :2]) / np.abs(ref32[ok, :2])
    assert np.median(rel) < 1e-5
    # packed scan = K1 -> K2 -> K3 in one call; chunking must not change results
    mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
    k_d, af_d, miss_d, out_d, ev = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"], return_evals=True)
    assert np.array_equal(k_d, keep) and np.array_equal(af_d.view(np.uint32), af.view(np.uint32))
    assert np.array_equal(miss_d, missing)
    assert_results_close(out_d, want)
    parts = [mdl.scan_packed(case.packed[i:i + 128], n, low=nm["low"], high=nm["high"])[3] for i in range(0, 600, 128)]
    assert np.array_equal(np.concatenate(parts), out_d, equal_nan=True)
    # LMM2 and fixed-lambda through the same packed entry point
    _, mlnull = oracle.lmm_ml_null_brent(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], 30, 1e-2)
    want2 = oracle.lmm_reml_lmm2_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"],
                                                    mlnull, 30, 1e-2)
    out2 = mdl.scan_packed(case.packed, n, mode="lmm2", low=nm["low"], high=nm["high"], nullml=mlnull)[3]
    assert_results_close(out2, want2, cols_p=(2, 5), cols_lambda=(3,))
    l10 = float(np.log10(nm["lbd"]))
    want3 = oracle.lmm_assoc_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], l10, g, nm["ut"])
    out3 = mdl.scan_packed(case.packed, n, mode="fvlmm", log10_lbd=l10)[3]
    assert_results_close(out3, want3)


def test_packed_array_entry_points_and_tsv_writer(jx, oracle, tmp_path):
    """lmm_reml_assoc_packed_f32[_to_tsv] (lmm.rs:3040-3800), GwasAssocTsvWriter, FvLmmAssocCache."""
    case = make_problem(n=300, m=260, q=2, seed=5, missing_rate=0.02)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, mr, missing = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    l10 = float(np.log10(nm["lbd"]))
    rot = oracle.rotate_block(g, nm["ut"])
    want = oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], rot, 50, 1e-2,
                                     nullml=nm["ml0"], init_log10_lbd=l10)
    flip = np.zeros(idx.size, bool)
    seen = []
    got = jx.lmm_reml_assoc_packed_f32(case.packed, n, flip, af[idx], case.s, nm["xcov"], nm["y"], nm["ut"],
                                       row_indices=idx, low=nm["low"], high=nm["high"], nullml=nm["ml0"],
                                       init_log10_lbd=l10, progress_callback=lambda d, t: seen.append((d, t)),
                                       progress_every=100)
    assert got.shape == (idx.size, 4) and seen[-1] == (idx.size, idx.size) and len(seen) == -(-idx.size // 100)
    assert_results_close(got, want, cols_p=(2, 3))
    # already-selected rows, no row_indices: identical
    got2 = jx.lmm_reml_assoc_packed_f32(case.packed[idx], n, flip, af[idx], case.s, nm["xcov"], nm["y"], nm["ut"],
                                        low=nm["low"], high=nm["high"], nullml=nm["ml0"], init_log10_lbd=l10)
    assert np.array_equal(got, got2, equal_nan=True)
    # a row_maf that is not the frequency over these samples is used as given (test_prepared_row_maf_and_flip_are_used_as_given)
    got3 = jx.lmm_reml_assoc_packed_f32(case.packed[idx], n, flip, af[idx] * np.float32(0.5), case.s, nm["xcov"], nm["y"],
                                        nm["ut"], low=nm["low"], high=nm["high"])
    assert got3.shape == (idx.size, 3)
    with pytest.raises(RuntimeError, match="packed second dimension mismatch"):
        jx.lmm_reml_assoc_packed_f32(case.packed[idx][:, :-1], n, flip, af[idx], case.s, nm["xcov"], nm["y"], nm["ut"])
    # to_tsv: same numbers, reference row format; metadata from arrays or from the BIM
    from janusx_b200 import synth
    prefix = str(tmp_path / "pk")
    synth.write_plink(prefix, case.packed, n)
    out_tsv = tmp_path / "pk.tsv"
    rows = jx.lmm_reml_assoc_packed_f32_to_tsv(case.packed, n, flip, af[idx], mr[idx], case.s, nm["xcov"], nm["y"], nm["ut"],
                                               [], [], [], [], [], str(out_tsv), row_indices=idx, low=nm["low"],
                                               high=nm["high"], nullml=nm["ml0"], init_log10_lbd=l10, bed_prefix=prefix)
    assert rows == idx.size
    bim = oracle.read_bim(prefix)
    lines = out_tsv.read_bytes().split(b"\n")
    assert lines[0] == b"chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\tplrt" and len(lines) == rows + 2
    for k in (0, 1, rows // 2, rows - 1):
        chrom, snp_id, pos, a0, a1 = bim[int(idx[k])]
        rate = np.float32(np.float32(round(float(mr[idx[k]]) * n)) / np.float32(n))
        assert lines[1 + k] + b"\n" == oracle.format_row(chrom, pos, snp_id, a0, a1, float(af[idx[k]]), float(rate), got[k])
    # writer class fed with device results
    w = jx.GwasAssocTsvWriter(str(tmp_path / "w.tsv"))
    sites = [jx.SiteInfo(bim[int(i)][0], bim[int(i)][2], bim[int(i)][3], bim[int(i)][4]) for i in idx]
    rate_all = (np.round(mr[idx].astype(np.float64) * n).astype(np.float32) / np.float32(n))
    w.write_chunk(sites, [bim[int(i)][1] for i in idx], af[idx], rate_all, got)
    w.close()
    assert (tmp_path / "w.tsv").read_bytes() == out_tsv.read_bytes()
    # fixed-lambda cache handle
    cache = jx.fvlmm_assoc_prepare_cache_f32(case.s, nm["xcov"], nm["y"], l10)
    assert (cache.n, cache.p) == (n, nm["xcov"].shape[1]) and abs(cache.lbd - nm["lbd"]) <= 1e-12 * nm["lbd"]
    want_f = oracle.lmm_assoc_chunk_f32(case.s, nm["xcov"], nm["y"], l10, rot, nullml=nm["ml0"])
    want_f = want_f[0] if isinstance(want_f, tuple) else want_f
    got_f = jx.fvlmm_assoc_chunk_with_cache_f32(cache, rot, nullml=nm["ml0"])
    assert_results_close(got_f, want_f, cols_p=(2, 3))
    with pytest.raises(RuntimeError, match="g_rot_chunk must be"):
        jx.fvlmm_assoc_chunk_with_cache_f32(cache, rot[:, :-1])
    # unrotated chunk through the cache handle, and the variant that returns ready-made TSV text (positional calls as in
    # pyBLUP/assoc.py:1512-1522 / workflow_model_stream.py:1680-1700)
    got_s = jx.fvlmm_assoc_chunk_from_snp_with_cache_f32(cache, g, nm["ut"], 0, nm["ml0"], 512)
    assert_results_close(got_s, want_f, cols_p=(2, 3))
    got_9 = jx.fvlmm_assoc_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], l10, g, nm["ut"], 0, None, 512)
    assert_results_close(got_9, want_f[:, :3])
    names = [bim[int(i)] for i in idx]
    blocks, n_rows = jx.fvlmm_assoc_chunk_from_snp_to_tsv_f32(
        case.s, nm["xcov"], nm["y"], l10, g, nm["ut"], [b[0] for b in names], [b[2] for b in names], [b[1] for b in names],
        [b[3] for b in names], [b[4] for b in names], af[idx].tolist(), rate_all.tolist(), threads=0, nullml=None)
    assert n_rows == idx.size and len(blocks) == 1
    first = blocks[0].split(b"\n")[0] + b"\n"
    assert first == oracle.format_row(names[0][0], names[0][2], names[0][1], names[0][3], names[0][4], float(af[idx[0]]),
                                      float(rate_all[0]), got_9[0])


def test_sample_subset_scan(jx, oracle):
    case = make_problem(n=300, m=120, q=2, seed=91, missing_rate=0.03)
    sidx = np.array(sorted(np.random.default_rng(1).choice(300, size=211, replace=False)), dtype=np.int64)
    import copy
    sub = copy.copy(case)
    # null model on the subset: recompute K on those samples
    from janusx_b200 import synth
    K = synth.vanraden_grm(case.packed, case.n)[np.ix_(sidx, sidx)]
    K[np.diag_indices(len(sidx))] += 1e-6
    s, u = np.linalg.eigh(K)
    sub.s, sub.u, sub.n, sub.y, sub.cov = s, u, len(sidx), case.y[sidx], case.cov[sidx]
    nm = null_model(oracle, sub)
    keep, af, mr, missing = oracle.count_qc_block(case.packed, case.n, sidx, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, case.n, af[idx], sample_idx=sidx, row_indices=idx)
    want = oracle.lmm_reml_chunk_from_snp_f32(s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30, 1e-2)
    mdl = jx.DeviceModel(s, nm["xcov"], nm["y"], nm["ut"])
    k_d, af_d, miss_d, out_d = mdl.scan_packed(case.packed, case.n, sample_idx=sidx, low=nm["low"], high=nm["high"])
    assert np.array_equal(k_d, keep) and np.array_equal(af_d.view(np.uint32), af.view(np.uint32))
    assert np.array_equal(miss_d, missing)
    assert_results_close(out_d, want)


def _tsv_fields(path):
    lines = path.read_bytes().split(b"\n")
    assert lines[-1] == b""
    return lines[0], [l.split(b"\t") for l in lines[1:-1]]


def _last_digit_unit(text):
    """Value of one unit in the last printed digit of a TSV number (`{:.4}` / `{:.4e}` / `{:.6e}` renderings)."""
    text = text.decode() if isinstance(text, bytes) else text
    if "e" in text:
        mant, ex = text.split("e")
        digits = len(mant.split(".")[1]) if "." in mant else 0
        return 10.0 ** (int(ex) - digits)
    return 10.0 ** (-(len(text.split(".")[1]) if "." in text else 0))


def _assert_row_equiv(a, b):
    """Two TSV rows of the same SNP: identification columns byte-exact; a numeric field may differ only by ONE unit of
    its last printed digit (two doubles within the 1e-8 gate on either side of a rounding boundary)."""
    assert a[:7] == b[:7]          # chrom pos snp alleles af miss: byte-exact
    for x, y in zip(a[7:], b[7:]):
        if x == y:
            continue
        fx, fy = float(x), float(y)
        assert not (math.isnan(fx) or math.isnan(fy)), (x, y)
        assert abs(fx - fy) <= 1.0001 * max(_last_digit_unit(x), _last_digit_unit(y)), (x, y)


def _assert_tsv_equiv(got_path, want_path):
    hg, rg = _tsv_fields(got_path)
    hw, rw = _tsv_fields(want_path)
    assert hg == hw and len(rg) == len(rw)
    mismatched = 0
    for a, b in zip(rg, rw):
        if a != b:
            mismatched += 1
            _assert_row_equiv(a, b)
    assert mismatched <= max(1, len(rw) // 200)


def test_bed_to_tsv_matches_oracle(jx, oracle, tmp_path):
    from janusx_b200 import synth
    case = make_problem(n=320, m=900, q=2, seed=55, missing_rate=0.03)
    nm = null_model(oracle, case)
    prefix = str(tmp_path / "panel")
    ids = [f"rs{i}" if i % 11 else "." for i in range(900)]
    synth.write_plink(prefix, case.packed, case.n, snp_ids=ids)
    args = (case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05, 1.0)
    seen = []
    rows = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "g.tsv"), *args, low=nm["low"], high=nm["high"],
                                            rotate_block_rows=256, progress_callback=lambda d, t: seen.append((d, t)),
                                            progress_every=300)
    rows_o = oracle.scan_bed_to_tsv(prefix, str(tmp_path / "o.tsv"), *args, low=nm["low"], high=nm["high"])
    assert rows == rows_o and rows > 0 and seen and seen[-1] == (900, 900)
    _assert_tsv_equiv(tmp_path / "g.tsv", tmp_path / "o.tsv")
    # LMM2 (null ML fitted inside, seeded like the CLI) and the fixed-lambda scan
    l10 = float(np.log10(nm["lbd"]))
    rows2 = jx.lmm_reml_lmm2_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "g2.tsv"), *args, low=nm["low"],
                                                  high=nm["high"], init_log10_lbd_reml=l10)
    rows2_o = oracle.scan_bed_to_tsv(prefix, str(tmp_path / "o2.tsv"), *args, low=nm["low"], high=nm["high"],
                                     model="lmm2", init_log10_lbd=l10)
    assert rows2 == rows2_o
    _assert_tsv_equiv(tmp_path / "g2.tsv", tmp_path / "o2.tsv")
    rows3, pve, ldv = jx.fvlmm_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "g3.tsv"), case.s, nm["xcov"], nm["y"], l10,
                                                    nm["ut"], 0.02, 0.05, 1.0)
    rows3_o = oracle.scan_bed_to_tsv(prefix, str(tmp_path / "o3.tsv"), *args, model="fvlmm", log10_lbd=l10)
    assert rows3 == rows3_o and 0.0 <= pve <= 1.0 and math.isfinite(ldv)
    _assert_tsv_equiv(tmp_path / "g3.tsv", tmp_path / "o3.tsv")
    # prepared row metadata (the default CLI route of the reference): listed rows are scanned without re-applying QC
    keep_o, af_o, mr_o, _ = oracle.count_qc_block(case.packed, case.n, None, 0.02, 0.05, 1.0)
    listed = np.nonzero(keep_o)[0]
    meta = dict(row_indices=listed, row_flip=np.zeros(listed.size, bool), row_missing=mr_o[listed], row_maf=af_o[listed])
    rows_p = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "gp.tsv"), *args, low=nm["low"], high=nm["high"],
                                              rotate_block_rows=256, **meta)
    assert rows_p == rows and (tmp_path / "gp.tsv").read_bytes() == (tmp_path / "g.tsv").read_bytes()
    sub = listed[::7]
    meta = dict(row_indices=sub, row_flip=np.zeros(sub.size, bool), row_missing=mr_o[sub], row_maf=af_o[sub])
    rows_s = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "gs.tsv"), *args, low=nm["low"], high=nm["high"], **meta)
    full_lines = (tmp_path / "g.tsv").read_bytes().split(b"\n")
    sub_lines = (tmp_path / "gs.tsv").read_bytes().split(b"\n")
    assert rows_s == sub.size and sub_lines[0] == full_lines[0]
    assert sub_lines[1:-1] == [full_lines[1 + i] for i in range(0, listed.size, 7)]
    # sample subset by IID + snps_only + error paths
    some = [f"S{j}" for j in range(case.n)]
    with pytest.raises(RuntimeError, match="sample 'nobody' not found in PLINK FAM"):
        jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "e.tsv"), *args, sample_ids=some[:-1] + ["nobody"])
    (tmp_path / "bad.bed").write_bytes(b"\x00\x01\x02" + bytes(10))
    (tmp_path / "bad.bim").write_text("")
    (tmp_path / "bad.fam").write_text("".join(f"F S{j} 0 0 0 -9\n" for j in range(case.n)))
    with pytest.raises(RuntimeError, match="only SNP-major BED supported"):
        jx.lmm_reml_assoc_bed_to_tsv_f32(str(tmp_path / "bad"), str(tmp_path / "e.tsv"), *args)


def test_model_objects_and_large_property_checks(jx, oracle):
    """LMM/LMM2/FvLMM objects (pyBLUP/assoc.py interface) + size-independent properties at a larger n."""
    from janusx_b200 import assoc, synth
    case = make_problem(n=1500, m=256, q=3, seed=123, missing_rate=0.01)
    n = case.n
    K = case.u @ np.diag(case.s) @ case.u.T
    K[np.diag_indices(n)] -= 1e-6
    lmm = assoc.LMM(case.y, case.cov, K)
    assert lmm.Dh.dtype == np.float32 and lmm.Xcov.shape == (n, 4) and lmm.y.shape == (n, 1)
    keep, af, _, _ = oracle.count_qc_block(case.packed, n, None, 0.0, 1.0, 1.0)
    g = oracle.decode_centered_block(case.packed, n, af)
    res = lmm.gwas(g, threads=1)
    assert res.shape == (256, 3) and np.all(np.isfinite(res))
    # chunked == unchunked (python/janusx/assoc/smoke.py:45-46)
    chunked = np.concatenate([lmm.gwas(g[i:i + 100], threads=2) for i in range(0, 256, 100)])
    assert np.array_equal(res, chunked)
    # oracle with the SAME spectral inputs the object holds
    want = oracle.lmm_reml_chunk_from_snp_f32(lmm.S, lmm.Xcov, lmm.y[:, 0], lmm.bounds[0], lmm.bounds[1], g, lmm.Dh, 30, 1e-2)
    assert_results_close(res, want)
    # scaling a SNP by c scales beta and se by 1/c and leaves p unchanged (exact powers of two)
    res_half = lmm.gwas((g * np.float32(0.5)).astype(np.float32))
    # (not exact: the 1e-6 ridge on the diagonal of Z'V^-1 Z, reml.rs:316-323, is not scale-invariant)
    np.testing.assert_allclose(res_half[:, :2], 2.0 * res[:, :2], rtol=1e-5)
    assert np.max(np.abs(np.log10(res_half[:, 2]) - np.log10(res[:, 2]))) < 1e-5
    lmm2 = assoc.LMM2.from_spectral(case.y, case.cov, case.s, case.u)
    r2 = lmm2.gwas(g[:64])
    assert r2.shape == (64, 6) and np.all(r2[:, 5] <= 1.0) and np.all(r2[:, 3] > 0)
    # different eigensolvers (cuSOLVER vs LAPACK) can flip a loose-Brent branch on a rare SNP: compare the bulk
    assert np.median(np.abs(r2[:, :2] - res[:64, :2]) / np.abs(res[:64, :2])) < 1e-6
    fv = assoc.FvLMM.from_spectral(case.y, case.cov, case.s, case.u)
    r3 = fv.gwas(g[:64])
    assert r3.shape == (64, 3)
    # fixed-lambda vs exact: same sign, similar magnitude for null SNPs
    assert np.corrcoef(r3[:, 0], res[:64, 0])[0, 1] > 0.99


def test_int8_sliced_rotation_variant(jx, oracle):
    """variant 2: exact int8-sliced tensor-core rotation must satisfy the same gates as the DMMA path
    (with and without missing calls, identity and subset samples)."""
    try:
        for miss, seed in ((0.0, 301), (0.03, 302)):
            case = make_problem(n=520, m=700, q=3, seed=seed, missing_rate=miss)
            nm = null_model(oracle, case)
            n = case.n
            keep, af, mr, missing = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
            idx = np.nonzero(keep)[0]
            g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
            want = oracle.lmm_reml_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], 30, 1e-2)
            mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
            jx.set_rotate_variant(2)
            k2, af2, ms2, out2 = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"])
            # variant 3: the hand-written tcgen05 kernel computes exactly what the library slice GEMMs compute
            jx.set_rotate_variant(3)
            k3, af3, ms3, out3 = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"])
            assert np.array_equal(k3, keep)
            assert_results_close(out3, want)
            assert_results_close(out3, out2)
            jx.set_rotate_variant(0)
            k0, af0, ms0, out0 = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"])
            assert np.array_equal(k2, keep) and np.array_equal(ms2, missing)
            assert_results_close(out2, want)
            assert_results_close(out2, out0)
            # chunking invariance with the int8 path
            jx.set_rotate_variant(2)
            parts = [mdl.scan_packed(case.packed[i:i + 300], n, low=nm["low"], high=nm["high"])[3] for i in range(0, 700, 300)]
            assert np.array_equal(np.concatenate(parts), out2, equal_nan=True)
            # non-additive coding falls back to the DMMA kernel
            d2 = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"], genetic_model="dom")[3]
            jx.set_rotate_variant(0)
            d0 = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"], genetic_model="dom")[3]
            assert np.array_equal(d2, d0, equal_nan=True)
    finally:
        jx.set_rotate_variant(3)


@pytest.mark.parametrize("p_cov", [1, 3, 4, 6])
def test_shared_abscissa_prefix_matches_plain_search(jx, oracle, p_cov):
    """The lane-per-SNP solve takes the first three objective values of every REML search (abscissae no SNP can change:
    src/math/brent.rs:16-136) from per-batch tables.  Forced on at a small size: same gates against the oracle, the
    oracle's evaluation counts, and rows bit-identical to the search that evaluates everything itself -- also when the
    search ends inside the prefix (max_iter 1..3), starts from a given point, or runs on the compiler's divide."""
    case = make_problem(n=210, m=160, q=p_cov - 1, seed=90 + p_cov, missing_rate=0.02)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, _, _ = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    rot = oracle.rotate_block(g, nm["ut"])
    _, mlnull = oracle.lmm_ml_null_brent(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], 30, 1e-2)
    want, ev_o = oracle.lmm_reml_lmm2_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], rot, mlnull, 30, 1e-2,
                                                return_evals=True)
    mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
    kw2 = dict(mode="lmm2", low=nm["low"], high=nm["high"], nullml=mlnull, return_evals=True)
    try:
        jx.set_thread_solve_min_rows(1)
        jx.set_prefix_evals(2)
        k, _, _, out, ev = mdl.scan_packed(case.packed, n, **kw2)
        assert np.array_equal(k, keep) and np.array_equal(ev, ev_o)
        assert_results_close(out, want, cols_p=(2, 5), cols_lambda=(3,))
        variants = [dict(kw2), dict(kw2, max_iter=1), dict(kw2, max_iter=2), dict(kw2, max_iter=3), dict(kw2, max_iter=4),
                    dict(kw2, init=0.37), dict(kw2, init=nm["low"]), dict(kw2, tol=1e-6, max_iter=60),
                    dict(mode="lmm", low=nm["low"], high=nm["high"], return_evals=True),
                    dict(mode="lmm", low=nm["low"], high=nm["high"], nullml=mlnull, max_iter=2, return_evals=True)]
        for divide in (0, 1):
            jx._cabi.lib().jxb_set_generic_divide(divide)
            for v in variants:
                jx.set_prefix_evals(2)
                a = mdl.scan_packed(case.packed, n, **v)
                a2 = mdl.scan_packed(case.packed, n, **v)           # second batch of the same search set-up: cached tables
                jx.set_prefix_evals(0)
                b = mdl.scan_packed(case.packed, n, **v)
                assert np.array_equal(a[3], b[3], equal_nan=True) and np.array_equal(a[4], b[4]), (divide, v)
                assert np.array_equal(a[3], a2[3], equal_nan=True) and np.array_equal(a[4], a2[4]), (divide, v)
        # a new phenotype on the same model invalidates the tables
        jx.set_prefix_evals(2)
        y2 = nm["y"] * 1.5 + 0.25
        mdl.set_xy(nm["xcov"], y2)
        a = mdl.scan_packed(case.packed, n, **kw2)
        jx.set_prefix_evals(0)
        b = mdl.scan_packed(case.packed, n, **kw2)
        assert np.array_equal(a[3], b[3], equal_nan=True) and np.array_equal(a[4], b[4])
        assert not np.array_equal(a[3], out, equal_nan=True)
    finally:
        jx._cabi.lib().jxb_set_generic_divide(0)
        jx.set_prefix_evals(1)
        jx.set_thread_solve_min_rows(32768)


@pytest.mark.parametrize("big_kernel", [0, 1])
@pytest.mark.parametrize("mode", ["lmm", "lmm2"])
def test_thread_per_snp_solve_kernel(jx, oracle, mode, big_kernel):
    """The large-batch solve kernels -- 0: lane-per-SNP with refill on the row-major block (default), 1: one thread per
    SNP on an SNP-minor block -- must satisfy the same gates as the warp kernel, evaluation for evaluation."""
    case = make_problem(n=450, m=500, q=3, seed=411, missing_rate=0.02)
    nm = null_model(oracle, case)
    n = case.n
    keep, af, mr, missing = oracle.count_qc_block(case.packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(case.packed, n, af[idx], row_indices=idx)
    mdl = jx.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
    try:
        jx.set_thread_solve_min_rows(1)
        jx.set_big_solve_kernel(big_kernel)
        if mode == "lmm":
            want, ev_o = oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"],
                                                   oracle.rotate_block(g, nm["ut"]), 30, 1e-2, return_evals=True)
            k, _, _, out, ev = mdl.scan_packed(case.packed, n, low=nm["low"], high=nm["high"], return_evals=True)
            assert_results_close(out, want)
            assert np.array_equal(ev, ev_o)
        else:
            _, mlnull = oracle.lmm_ml_null_brent(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], 30, 1e-2)
            want = oracle.lmm_reml_lmm2_chunk_from_snp_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], g, nm["ut"], mlnull, 30, 1e-2)
            k, _, _, out = mdl.scan_packed(case.packed, n, mode="lmm2", low=nm["low"], high=nm["high"], nullml=mlnull)
            assert_results_close(out, want, cols_p=(2, 5), cols_lambda=(3,))
            np.testing.assert_allclose(out[:, 4], want[:, 4], rtol=1e-10)
        assert np.array_equal(k, keep)
        # switching back to the warp kernel on the same model (row-major block) still works
        jx.set_thread_solve_min_rows(1 << 30)
        out_w = mdl.scan_packed(case.packed, n, mode=mode, low=nm["low"], high=nm["high"],
                                nullml=(mlnull if mode == "lmm2" else None))[3]
        assert_results_close(out_w, want, cols_p=((2, 5) if mode == "lmm2" else (2,)))
    finally:
        jx.set_thread_solve_min_rows(32768)
        jx.set_big_solve_kernel(0)


def test_full_size_n20000_parity_sample(jx, oracle):
    """BASELINE.json configs[2] sample count (n = 20,000, 3 covariates, LMM2): a 96-SNP sample through the
    bench path (tcgen05 int8 rotation + thread-per-SNP solve) and through the FP64 DMMA + warp-solve path, both
    against the CPU oracle at the north-star gates."""
    import torch
    sys_path_bench = __import__("sys").path
    from pathlib import Path
    root = str(Path(__file__).resolve().parents[1])
    if root not in sys_path_bench:
        sys_path_bench.insert(0, root)
    import bench as B
    n, m, q = 20000, 96, 3
    dev = torch.device("cuda:0")
    s_np, u_t_dev, X_np, y_np = B.build_null_model(torch, n, 4096, q, dev)
    ut = u_t_dev.cpu().numpy()
    pk, _ = B.gen_packed_batch(torch, n, m, 777, dev)
    packed = pk.cpu().numpy()
    packed[5, :40] = 0b01010101          # a few missing calls -> exercises the missing-indicator pass
    mdl = jx.DeviceModel(s_np, np.ones((n, q + 1)), np.zeros(n), u_t_dev, device=0, u_t_on_device=True)
    del u_t_dev
    xcov, yrot = mdl.rotate_xy(X_np, y_np)
    xo, yo = oracle.lmm_rotate_x_y_with_ut_f64(ut, X_np, y_np)
    np.testing.assert_allclose(xcov, xo, rtol=1e-10, atol=1e-11)
    # use the ORACLE's rotated design on both sides so the comparison isolates the scan
    mdl.set_xy(xo, yo[:, 0])
    lbd, ml0, reml0 = mdl.reml_null(-5.0, 5.0, 50, 1e-3)
    lbd_o, ml0_o, reml0_o = oracle.lmm_reml_null_f32(s_np, xo, yo[:, 0], -5.0, 5.0, 50, 1e-3)
    assert math.isclose(lbd, lbd_o, rel_tol=RTOL_LAMBDA) and math.isclose(reml0, reml0_o, rel_tol=1e-10)
    l10 = float(np.log10(lbd_o))
    lo, hi = l10 - 2.0, l10 + 2.0
    _, nullml = oracle.lmm_ml_null_brent(s_np, xo, yo[:, 0], lo, hi, 30, 1e-2, l10)
    keep, af, mr, missing = oracle.count_qc_block(packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    g = oracle.decode_centered_block(packed, n, af[idx], row_indices=idx)
    want = oracle.lmm_reml_lmm2_chunk_f32(s_np, xo, yo[:, 0], lo, hi, oracle.rotate_block(g, ut), nullml, 30, 1e-2,
                                          init_reml=l10)
    kw = dict(mode="lmm2", low=lo, high=hi, init=l10, nullml=nullml)
    try:
        jx.set_thread_solve_min_rows(1)
        k1, af1, ms1, out1 = mdl.scan_packed(packed, n, **kw)          # tcgen05 + lane kernel (bench path)
        jx.set_big_solve_kernel(1)
        _, _, _, out1b = mdl.scan_packed(packed, n, **kw)              # tcgen05 + thread kernel (SNP-minor block)
        jx.set_big_solve_kernel(0)
        assert np.array_equal(out1, out1b, equal_nan=True)             # same per-SNP arithmetic, bit for bit
        jx.set_thread_solve_min_rows(1 << 30)
        k2, _, _, out2 = mdl.scan_packed(packed, n, **kw)                # tcgen05 + warp kernel
        jx.set_rotate_variant(0)
        k3, _, _, out3 = mdl.scan_packed(packed, n, **kw)                # FP64 DMMA + warp kernel
    finally:
        jx.set_rotate_variant(3)
        jx.set_thread_solve_min_rows(32768)
    assert np.array_equal(k1, keep) and np.array_equal(ms1, missing)
    assert np.array_equal(af1.view(np.uint32), af.view(np.uint32))
    for out in (out1, out2, out3):
        assert_results_close(out, want, cols_p=(2, 5), cols_lambda=(3,))
        np.testing.assert_allclose(out[:, 4], want[:, 4], rtol=1e-10)


def test_rcp_fast_equals_ieee_divide(jx):
    """The branch-free reciprocal of the large-batch solve kernels against `1.0 / x` on 2^26 values per exponent band
    (random and extreme mantissas): zero differing bit patterns inside the range the host admits it for."""
    import ctypes as C
    lib = jx._cabi.lib()
    for lo, hi in ((-30, 30), (-960, -900), (900, 960), (-1, 1)):
        bad = C.c_uint64(123)
        jx._cabi.check(lib.jxb_selftest_rcp(1 << 26, lo, hi, C.byref(bad)))
        assert bad.value == 0, (lo, hi, bad.value)


def test_prepared_row_maf_and_flip_are_used_as_given(jx, oracle, tmp_path):
    """Prepared row metadata as the reference consumes it (src/decode/decode.rs:163-219, src/stats/lmm.rs:1237-1262):
    `row_maf` is the imputation frequency even when it is NOT the frequency over the scanned samples
    (workflow_model_packed.py:1296 passes full-sample values), `row_flip` reverses the code LUT, and the BED entry point
    prints `row_maf` / round(row_missing * n) / n in the af / miss columns."""
    from janusx_b200 import synth
    case = make_problem(n=320, m=240, q=2, seed=91, missing_rate=0.04)
    nm = null_model(oracle, case)
    n, m = case.n, 240
    rng = np.random.default_rng(5)
    _, af, mr, _ = oracle.count_qc_block(case.packed, n, None, 0.0, 1.0, 0.0)
    flip = rng.random(m) < 0.3
    row_maf = np.where(flip, 1.0 - af, af).astype(np.float32)
    row_maf[::5] = np.float32(0.37)                                  # deliberately not the sample frequency
    g = oracle.decode_centered_block(case.packed, n, row_maf, flip=flip)
    want = oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], oracle.rotate_block(g, nm["ut"]), 50, 1e-2)
    got = jx.lmm_reml_assoc_packed_f32(case.packed, n, flip, row_maf, case.s, nm["xcov"], nm["y"], nm["ut"], low=nm["low"],
                                       high=nm["high"])
    assert_results_close(got, want)
    # non-additive coding takes the FP64 decode path: same override
    g_d = oracle.decode_centered_block(case.packed, n, row_maf, flip=flip, model="dom")
    want_d = oracle.lmm_reml_chunk_f32(case.s, nm["xcov"], nm["y"], nm["low"], nm["high"], oracle.rotate_block(g_d, nm["ut"]), 50, 1e-2)
    got_d = jx.lmm_reml_assoc_packed_f32(case.packed, n, flip, row_maf, case.s, nm["xcov"], nm["y"], nm["ut"], low=nm["low"],
                                         high=nm["high"], model="dom")
    assert_results_close(got_d, want_d)
    # BED entry point with the four prepared arrays
    prefix = str(tmp_path / "prep")
    synth.write_plink(prefix, case.packed, n)
    listed = np.arange(3, m, 2)
    miss_rate = mr[listed].copy()
    miss_rate[0] = np.float32(0.0301)                                # round(0.0301 * 320) = 10 -> 10 / 320
    rows = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "p.tsv"), case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05,
                                            1.0, low=nm["low"], high=nm["high"], max_iter=50, row_indices=listed,
                                            row_flip=flip[listed], row_missing=miss_rate, row_maf=row_maf[listed])
    assert rows == listed.size
    head, body = _tsv_fields(tmp_path / "p.tsv")
    bim = oracle.read_bim(prefix)
    for k, j in enumerate(listed):
        chrom, snp_id, pos, a0, a1 = bim[j]
        cnt = 0.0 if not (np.isfinite(miss_rate[k]) and miss_rate[k] > 0) else np.round(float(miss_rate[k]) * n)
        rate = float(np.float32(cnt) / np.float32(n))
        line = oracle.format_row(chrom, pos, snp_id, a0, a1, float(row_maf[j]), rate, want[j]).rstrip(b"\n").split(b"\t")
        _assert_row_equiv(body[k], line)
    assert body[0][6] == b"0.0312"


@pytest.mark.parametrize("n,model", [(20000, "lmm2"), (50000, "lmm")])
def test_full_batch_at_full_size_sampled_parity(jx, oracle, n, model):
    """One full device batch (75,776 SNP rows) at the sample counts of BASELINE.json configs[2] / [4] (n = 20,000: -lmm2,
    then -fvlmm on the same batch) and configs[3] (n = 50,000, -lmm): rows sampled from the start, middle and end of the
    batch are checked against the oracle, which catches 32-bit index overflow in any kernel (rows x n exceeds 2^31 at
    n = 50,000).  The n = 20,000 batch runs the streamed scan (rotation slabs under the persistent solve kernel) and
    is repeated with the compiler's divide and through the streamed (overlapped) scan: bit-identical."""
    import sys
    from pathlib import Path
    import torch
    root = str(Path(__file__).resolve().parents[1])
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench as B
    rows, q = 75776, 3
    dev = torch.device("cuda:0")
    # n > 46,340: two unrelated half-size populations = block-diagonal GRM / U^T (bench.build_null_model)
    s_np, u_t_dev, X_np, y_np = B.build_null_model(torch, n, 4096, q, dev)
    ut = u_t_dev.cpu().numpy()
    mdl = jx.DeviceModel(s_np, np.ones((n, q + 1)), np.zeros(n), u_t_dev, device=0, u_t_on_device=True)
    del u_t_dev
    torch.cuda.empty_cache()
    xo, yo = oracle.lmm_rotate_x_y_with_ut_f64(ut, X_np, y_np)
    mdl.set_xy(xo, yo[:, 0])
    lbd_o, _, _ = oracle.lmm_reml_null_f32(s_np, xo, yo[:, 0], -5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd_o))
    lo, hi = l10 - 2.0, l10 + 2.0
    nullml = None
    if model == "lmm2":
        _, nullml = oracle.lmm_ml_null_brent(s_np, xo, yo[:, 0], lo, hi, 30, 1e-2, l10)
    pk, _ = B.gen_packed_batch(torch, n, rows, 4242, dev)
    pk[3, :50] = 0b01010101            # missing calls in one row: the whole batch takes the missing-indicator pass
    bps = pk.shape[1]
    torch.cuda.synchronize(dev)          # the scan runs on the model's own stream
    kw = dict(mode=model, low=lo, high=hi, init=(l10 if model == "lmm2" else None), nullml=nullml)
    cols = mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, **kw)
    keep, af, missing, out, ev = mdl.scan_fetch(rows, cols)
    streamed = mdl.stage_ms()["streamed"] > 0
    assert keep.sum() == out.shape[0], (int(keep.sum()), out.shape)
    assert keep.sum() > rows * 0.99, int(keep.sum())
    pos = np.cumsum(keep) - 1                                   # source row -> compacted row
    picks = [r for r in list(range(0, 6)) + list(range(rows // 2, rows // 2 + 6)) + list(range(rows - 6, rows)) if keep[r]]
    sub = pk[picks].cpu().numpy()
    keep_o, af_o, _, miss_o = oracle.count_qc_block(sub, n, None, 0.02, 0.05, 1.0)
    assert keep_o.all() and np.array_equal(af_o.view(np.uint32), af[picks].view(np.uint32))
    assert np.array_equal(miss_o, missing[picks])
    g = oracle.decode_centered_block(sub, n, af_o)
    rot_o = oracle.rotate_block(g, ut)
    got = out[pos[picks]]
    if model == "lmm2":
        want, ev_o = oracle.lmm_reml_lmm2_chunk_f32(s_np, xo, yo[:, 0], lo, hi, rot_o, nullml, 30, 1e-2, init_reml=l10), None
        assert_results_close(got, want, cols_p=(2, 5), cols_lambda=(3,))
        np.testing.assert_allclose(got[:, 4], want[:, 4], rtol=1e-10)
    else:
        want, ev_o = oracle.lmm_reml_chunk_f32(s_np, xo, yo[:, 0], lo, hi, rot_o, 30, 1e-2, return_evals=True)
        assert_results_close(got, want)
        assert np.array_equal(ev[pos[picks]], ev_o)             # the same Brent path, evaluation for evaluation
    assert not streamed
    # the same batch without the shared-abscissa prefix (every evaluation of every search done by the lane kernel itself):
    # bit-identical rows and evaluation counts
    jx._cabi.lib().jxb_set_prefix_evals(0)
    try:
        mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, **kw)
        keep0, af0, missing0, out0, ev0 = mdl.scan_fetch(rows, cols)
    finally:
        jx._cabi.lib().jxb_set_prefix_evals(1)
    assert np.array_equal(keep, keep0) and np.array_equal(out, out0, equal_nan=True) and np.array_equal(ev, ev0)
    if n <= 24000:
        # the same batch with the compiler-generated divide instead of rcp_fast, and through the streamed scan (rotation
        # slabs under one persistent solve kernel): same arithmetic, bit-identical rows
        jx._cabi.lib().jxb_set_generic_divide(1)
        try:
            mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, **kw)
            keep1, af1, missing1, out1, ev1 = mdl.scan_fetch(rows, cols)
        finally:
            jx._cabi.lib().jxb_set_generic_divide(0)
        assert np.array_equal(out, out1, equal_nan=True) and np.array_equal(ev, ev1)
        jx.set_stream_overlap(True)
        try:
            mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, **kw)
            keep2, af2, missing2, out2, ev2 = mdl.scan_fetch(rows, cols)
            assert mdl.stage_ms()["streamed"] == 1
        finally:
            jx.set_stream_overlap(False)
        assert np.array_equal(keep, keep2) and np.array_equal(out, out2, equal_nan=True) and np.array_equal(ev, ev2)
        # BASELINE.json configs[4]: -fvlmm (fixed lambda = null lambda) on the same full batch
        cols_f = mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, mode="fvlmm", log10_lbd=l10)
        keep_f, _, _, out_f, _ = mdl.scan_fetch(rows, cols_f)
        assert np.array_equal(keep_f, keep) and out_f.shape == (int(keep.sum()), 3)
        want_f, _ = oracle.lmm_assoc_chunk_f32(s_np, xo, yo[:, 0], l10, rot_o, 0, None)
        assert_results_close(out_f[pos[picks]], want_f)
        assert np.isfinite(out_f[:, :2]).all()
        # the lane-per-SNP fixed-lambda kernel (large batches) against the warp-per-SNP one: same ordered sums, same bits
        jx._cabi.lib().jxb_set_fixed_lane_min_rows(1 << 40)
        try:
            mdl.scan_packed_dev(pk.data_ptr(), rows, bps, n, None, mode="fvlmm", log10_lbd=l10)
            _, _, _, out_w, _ = mdl.scan_fetch(rows, cols_f)
        finally:
            jx._cabi.lib().jxb_set_fixed_lane_min_rows(4096)
        assert np.array_equal(out_f, out_w, equal_nan=True)


def test_config2_bed_to_tsv_n5000(jx, oracle, tmp_path):
    """BASELINE.json configs[1] shape: n = 5,000, 3 covariates, -lmm through the file-level entry point
    (jxb_scan_bed_to_tsv) on a 131,072-SNP PLINK file.  Every row's identification / af / miss columns are checked
    against the oracle's counts; 18 sampled rows are checked as text and as numbers at the north-star gates."""
    import sys
    from pathlib import Path
    import torch
    from janusx_b200 import synth
    root = str(Path(__file__).resolve().parents[1])
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench as B
    n, m, q = 5000, 131072, 3
    dev = torch.device("cuda:0")
    s_np, u_t_dev, X_np, y_np = B.build_null_model(torch, n, 4096, q, dev)
    ut = u_t_dev.cpu().numpy()
    del u_t_dev
    xo, yo = oracle.lmm_rotate_x_y_with_ut_f64(ut, X_np, y_np)
    lbd_o, _, _ = oracle.lmm_reml_null_f32(s_np, xo, yo[:, 0], -5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd_o))
    lo, hi = l10 - 2.0, l10 + 2.0
    packed = B.gen_snp_range(torch, n, 0, m, dev).cpu().numpy()
    packed[17, :30] = 0b01010101          # a row with missing calls
    prefix = str(tmp_path / "c2")
    synth.write_plink(prefix, packed, n)
    out_tsv = tmp_path / "c2.lmm.tsv"
    rows = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(out_tsv), s_np, xo, yo[:, 0], ut, 0.02, 0.05, 1.0, low=lo, high=hi)
    keep, af, mr, missing = oracle.count_qc_block(packed, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    assert rows == idx.size > 0.95 * m
    head, body = _tsv_fields(out_tsv)
    assert len(head.split(b"\t")) == 11 and len(body) == rows
    bim = oracle.read_bim(prefix)
    for k in range(0, rows, 97):                      # SNP order, af and miss columns over the whole file
        j = idx[k]
        chrom, snp_id, pos, a0, a1 = bim[j]
        name = snp_id if snp_id not in ("", ".") else f"{chrom}_{pos}"
        rate = np.float32(missing[j]) / np.float32(n)
        assert body[k][:7] == [v.encode() for v in (chrom, str(pos), name, a0, a1, "%.4f" % float(af[j]), "%.4f" % float(rate))], k
    picks = np.concatenate([np.arange(0, 6), np.arange(rows // 2, rows // 2 + 6), np.arange(rows - 6, rows)])
    src = idx[picks]
    g = oracle.decode_centered_block(packed, n, af[src], row_indices=src)
    want = oracle.lmm_reml_chunk_f32(s_np, xo, yo[:, 0], lo, hi, oracle.rotate_block(g, ut), 30, 1e-2)
    for k, j, w in zip(picks, src, want):
        chrom, snp_id, pos, a0, a1 = bim[j]
        rate = float(np.float32(missing[j]) / np.float32(n))
        line = oracle.format_row(chrom, pos, snp_id, a0, a1, float(af[j]), rate, w).rstrip(b"\n").split(b"\t")
        _assert_row_equiv(body[k], line)
    # the same rows as numbers, through the packed entry point of the same model
    mdl = jx._get_model(s_np, xo, yo[:, 0], ut)
    k_d, af_d, miss_d, out_d = mdl.scan_packed(np.ascontiguousarray(packed[src]), n, low=lo, high=hi, maf_thr=0.0, miss_thr=1.0,
                                               het_thr=0.0)
    assert k_d.all() and np.array_equal(af_d.view(np.uint32), af[src].view(np.uint32)) and np.array_equal(miss_d, missing[src])
    assert_results_close(out_d, want)


def test_cli_gwas_lmm_end_to_end(jx, oracle, tmp_path):
    """`python -m janusx_b200.gwas` with the reference's flag names writes the reference TSV schema and naming;
    rows agree with the oracle run on the same null model."""
    from janusx_b200 import gwas, synth
    case = make_problem(n=200, m=300, q=0, seed=88, missing_rate=0.02)
    prefix = str(tmp_path / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    with open(tmp_path / "pheno.tsv", "w") as fh:
        fh.write("id\ttraitA\n")
        for j in range(case.n):
            fh.write(f"S{j}\t{case.y[j]:.10f}\n" if j != 7 else f"S{j}\tNA\n")   # one missing phenotype
    rc = gwas.main(["-bfile", prefix, "-p", str(tmp_path / "pheno.tsv"), "-n", "0", "-lmm", "-lmm2", "-fvlmm",
                    "-k", "1", "-mem", "64MB", "-o", str(tmp_path / "out"), "-prefix", "run"])
    assert rc == 0
    for model, ncols in (("lmm", 11), ("lmm2", 14), ("fvlmm", 11)):
        lines = (tmp_path / "out" / f"run.traitA.{model}.tsv").read_text().splitlines()
        assert lines[0].split("\t")[:5] == ["chrom", "pos", "snp", "allele0", "allele1"]
        assert len(lines[0].split("\t")) == ncols and len(lines) > 100
        assert all(len(l.split("\t")) == ncols for l in lines[1:])
    assert not list((tmp_path / "out").glob("*.tmp"))


def test_cli_lm_switch_without_force_model(jx, tmp_path, capsys):
    """workflow_model_stream.py:930-963: a trait whose null LRT of Va = 0 is not significant is switched to LM by the
    reference.  LM is outside this build, so the trait is skipped with the reference's warning (exit code 3) unless
    -force-model keeps the mixed model."""
    from janusx_b200 import gwas, synth
    case = make_problem(n=240, m=400, q=0, seed=5, missing_rate=0.0)
    prefix = str(tmp_path / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    rng = np.random.default_rng(1)
    with open(tmp_path / "pheno.tsv", "w") as fh:
        fh.write("id\tnoise\n")
        for j in range(case.n):
            fh.write(f"S{j}\t{rng.normal():.10f}\n")      # no genetic component at all
    common = ["-bfile", prefix, "-p", str(tmp_path / "pheno.tsv"), "-lmm", "-k", "1", "-o", str(tmp_path / "out"), "-prefix", "r"]
    rc = gwas.main(common)
    err = capsys.readouterr().err
    assert rc == 3 and "switch to LM for trait noise" in err and not (tmp_path / "out" / "r.noise.lmm.tsv").exists()
    rc = gwas.main(common + ["-force-model"])
    assert rc == 0 and (tmp_path / "out" / "r.noise.lmm.tsv").exists()


def test_graft_entry_smoke():
    """`__graft_entry__.smoke()` (what the driver runs before the bench): both kernel sets against the oracle."""
    import importlib
    import sys
    root = str(__import__("pathlib").Path(__file__).resolve().parents[1])
    if root not in sys.path:
        sys.path.insert(0, root)
    importlib.import_module("__graft_entry__").smoke()
