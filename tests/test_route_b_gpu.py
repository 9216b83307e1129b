"""Route B: BedChunkReader.next_chunk_prepared -> LMM.gwas / LMM2.gwas / FvLMM.gwas -> GwasAssocTsvWriter,
the chain the reference runs when several models share one trait's decoded chunks.  The reader's rows must equal the
row-by-row restatement of src/io/gfreader.rs:3580-3700 bit for bit; the chunked chain must equal the unchunked one."""
from pathlib import Path

import numpy as np
import pytest

from conftest import make_problem, null_model
from test_parity_gpu import assert_results_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jx():
    from janusx_b200 import jxrs
    yield jxrs
    jxrs.clear_model_cache()


def _read_all(reader, chunk, **kw):
    gs, sites, afs, ms, sizes = [], [], [], [], []
    while True:
        out = reader.next_chunk_prepared(chunk, **kw)
        if out is None:
            break
        g, s, af, miss = out
        assert g.dtype == np.float32 and g.shape == (len(s), reader.n_samples) and af.shape == miss.shape == (len(s),)
        gs.append(g); sites.extend(s); afs.append(af); ms.append(miss); sizes.append(len(s))
    return np.concatenate(gs), sites, np.concatenate(afs), np.concatenate(ms), sizes


def test_reader_rows_bit_exact_and_chunk_invariant(jx, oracle, tmp_path):
    from janusx_b200 import synth
    from janusx_b200.gfreader import BedChunkReader
    n_full, m = 211, 700
    packed, _ = synth.draw_genotypes(m, n_full, seed=17, missing_rate=0.04)
    packed[5] = 0b01010101
    packed[6] = 0
    prefix = str(tmp_path / "rb")
    ids = [f"rs{i}" if i % 9 else "." for i in range(m)]
    synth.write_plink(prefix, packed, n_full, snp_ids=ids)
    fam = oracle.read_fam(prefix)
    sub_ids = [fam[i] for i in sorted(np.random.default_rng(2).choice(n_full, size=150, replace=False))]
    sub_idx = np.array([fam.index(s) for s in sub_ids], dtype=np.int64)
    for kw, sidx in ((dict(), None), (dict(sample_ids=sub_ids), sub_idx)):
        for thr in ((0.02, 0.05, 1.0), (0.0, 1.0, 1.0), (0.05, 0.1, 0.45)):
            keep, g_o, af_o, miss_o = oracle.bed_chunk_prepared_rows(packed, n_full, sidx, *thr)
            rd = BedChunkReader(prefix, maf_threshold=thr[0], max_missing_rate=thr[1], het_threshold=thr[2], **kw)
            assert rd.n_snps == m and rd.n_samples == (n_full if sidx is None else 150)
            g, sites, af, miss, sizes = _read_all(rd, 128)
            assert sizes[:-1] == [128] * (len(sizes) - 1) and sum(sizes) == int(keep.sum())   # chunks are filled up
            assert np.array_equal(g.view(np.uint32), g_o.view(np.uint32))
            assert np.array_equal(af.view(np.uint32), af_o.view(np.uint32)) and np.array_equal(miss, miss_o)
            bim = oracle.read_bim(prefix)
            assert [(s.chrom, s.snp, s.pos, s.ref_allele, s.alt_allele) for s in sites] == [bim[i] for i in np.nonzero(keep)[0]]
            g2 = _read_all(BedChunkReader(prefix, maf_threshold=thr[0], max_missing_rate=thr[1], het_threshold=thr[2], **kw), 10_000)[0]
            assert np.array_equal(g2.view(np.uint32), g.view(np.uint32))
    # snp_range / snp_indices / snps_only and error messages
    rd = BedChunkReader(prefix, snp_range=(100, 300))
    assert rd.n_snps == 200 and sum(_read_all(rd, 64)[4]) == 200
    rd = BedChunkReader(prefix, snp_indices=[7, 3, 500])
    assert [s.pos for s in _read_all(rd, 2)[1]] == [oracle.read_bim(prefix)[i][2] for i in (7, 3, 500)]
    with pytest.raises(RuntimeError, match="sample id not found"):
        BedChunkReader(prefix, sample_ids=["nobody"])
    with pytest.raises(RuntimeError, match="invalid snp_range"):
        BedChunkReader(prefix, snp_range=(5, 5))
    with pytest.raises(ValueError, match="chunk_size must be > 0"):
        BedChunkReader(prefix).next_chunk_prepared(0)
    assert jx.BedChunkReader is BedChunkReader


def test_route_b_chain_matches_oracle(jx, oracle, tmp_path):
    """reader -> LMM / LMM2 / FvLMM .gwas -> GwasAssocTsvWriter, chunked, against the oracle on the same rows."""
    from janusx_b200 import assoc, synth
    from janusx_b200.gfreader import BedChunkReader
    case = make_problem(n=260, m=420, q=2, seed=23, missing_rate=0.03)
    prefix = str(tmp_path / "chain")
    synth.write_plink(prefix, case.packed, case.n)
    K = case.u @ np.diag(case.s) @ case.u.T - 1e-6 * np.eye(case.n)       # LMM adds its own 1e-6 ridge
    lmm = assoc.LMM(case.y, case.cov, K)
    keep, g_o, af_o, miss_o = oracle.bed_chunk_prepared_rows(case.packed, case.n, None, 0.02, 0.05, 1.0)
    ut = lmm.Dh
    want = oracle.lmm_reml_chunk_from_snp_f32(lmm.S, lmm.Xcov, lmm.y[:, 0], lmm.bounds[0], lmm.bounds[1], g_o, ut, 30, 1e-2)
    rd = BedChunkReader(prefix, maf_threshold=0.02, max_missing_rate=0.05)
    w = jx.GwasAssocTsvWriter(str(tmp_path / "b.tsv"))
    outs = []
    while True:
        nxt = rd.next_chunk_prepared(100)
        if nxt is None:
            break
        g, sites, af, miss = nxt
        res = lmm.gwas(g)
        outs.append(res)
        w.write_chunk(sites, [s.snp for s in sites], af, miss / np.float32(case.n), res)
    w.close()
    got = np.concatenate(outs)
    assert_results_close(got, want)
    assert w.rows_written == int(keep.sum())
    lines = (tmp_path / "b.tsv").read_bytes().split(b"\n")
    assert len(lines) == w.rows_written + 2 and lines[0].startswith(b"chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta")
    bim = oracle.read_bim(prefix)
    k0 = int(np.nonzero(keep)[0][0])
    assert lines[1] + b"\n" == oracle.format_row(bim[k0][0], bim[k0][2], bim[k0][1], bim[k0][3], bim[k0][4], float(af_o[0]),
                                                 float(np.float32(miss_o[0]) / np.float32(case.n)), got[0])
    # LMM2 and FvLMM objects on the same chunks
    lmm2 = assoc.LMM2(case.y, case.cov, K)
    res2 = lmm2.gwas(g_o[:64])
    want2 = oracle.lmm_reml_lmm2_chunk_from_snp_f32(lmm2.S, lmm2.Xcov, lmm2.y[:, 0], lmm2.bounds[0], lmm2.bounds[1], g_o[:64],
                                                    lmm2.Dh, float(lmm2._lmm2_ml0_exact), 30, 1e-2)
    assert_results_close(res2, want2, cols_p=(2, 5), cols_lambda=(3,))
    fv = assoc.FvLMM(case.y, case.cov, K)
    res3 = fv.gwas(g_o[:64])
    want3 = oracle.lmm_assoc_chunk_from_snp_f32(fv.S, fv.Xcov, fv.y[:, 0], float(np.log10(fv.lbd_null)), g_o[:64], fv.Dh)
    want3 = want3[0] if isinstance(want3, tuple) else want3
    assert_results_close(res3, want3)


def test_prepared_row_statistics_feed_the_bed_scan(jx, oracle, tmp_path):
    """The default CLI route of the reference: prepare_bed_logic_meta_selected (src/io/gfreader.rs:7119-7232) once per
    trait over the trait's samples, then lmm_reml_assoc_bed_to_tsv_f32 with row_indices / row_missing / row_maf /
    row_flip (assoc/workflow.py:8870-8888).  Counts come from the device; the TSV equals the plain thresholded scan."""
    from janusx_b200 import synth
    from conftest import make_problem, null_model
    case = make_problem(n=260, m=700, q=1, seed=77, missing_rate=0.03)
    nm = null_model(oracle, case)
    prefix = str(tmp_path / "p")
    synth.write_plink(prefix, case.packed, case.n)
    row_idx, miss, af, flip, site_keep, n_full, n_snps = jx.prepare_bed_logic_meta_selected(prefix, None, 0.02, 0.05, 1.0)
    keep_o, af_o, mr_o, _ = oracle.count_qc_block(case.packed, case.n, None, 0.02, 0.05, 1.0)
    assert (n_full, n_snps) == (case.n, 700) and np.array_equal(site_keep, keep_o)
    assert np.array_equal(row_idx, np.nonzero(keep_o)[0]) and not flip.any()
    assert np.array_equal(af.view(np.uint32), af_o[keep_o].view(np.uint32))
    assert np.array_equal(miss.view(np.uint32), mr_o[keep_o].view(np.uint32))
    mask, n2, m2 = jx.prepare_bed_logic_keep_mask(prefix, None, 0.02, 0.05, 1.0)
    assert np.array_equal(mask, keep_o) and (n2, m2) == (case.n, 700)
    args = (case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05, 1.0)
    rows = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "a.tsv"), *args, low=nm["low"], high=nm["high"])
    rows_p = jx.lmm_reml_assoc_bed_to_tsv_f32(prefix, str(tmp_path / "b.tsv"), *args, low=nm["low"], high=nm["high"],
                                              row_indices=row_idx, row_flip=flip, row_missing=miss, row_maf=af)
    assert rows == rows_p == int(keep_o.sum())
    assert (tmp_path / "a.tsv").read_bytes() == (tmp_path / "b.tsv").read_bytes()
    # a sample subset: the statistics are those of the subset
    sub = np.arange(5, case.n, 2, dtype=np.int64)
    r2 = jx.prepare_bed_logic_meta_selected(prefix, sub, 0.05, 0.1, 1.0)
    k2, af2, mr2, _ = oracle.count_qc_block(case.packed, case.n, sub, 0.05, 0.1, 1.0)
    assert np.array_equal(r2[4], k2) and np.array_equal(r2[2].view(np.uint32), af2[k2].view(np.uint32))
    with pytest.raises(ValueError, match="sample index out of range"):
        jx.prepare_bed_logic_meta_selected(prefix, np.array([0, case.n]))


def test_bed_chunk_reader_from_meta(jx, oracle, tmp_path):
    """BedChunkReaderFromMeta (src/io/gfreader.rs:7440-7730): rows listed by the shared per-trait metadata, decoded as
    (code - 2*af) with missing -> 0 and no re-centring; chunking and snps_only follow the row list."""
    from janusx_b200 import synth
    case = make_problem(n=150, m=300, q=0, seed=12, missing_rate=0.05)
    prefix = str(tmp_path / "p")
    synth.write_plink(prefix, case.packed, case.n)
    sub = np.arange(3, case.n, 2, dtype=np.int64)
    row_idx, miss, af, flip, _, _, _ = jx.prepare_bed_logic_meta_selected(prefix, sub, 0.05, 0.2, 1.0)
    flip = flip.copy()
    flip[::7] = True                                   # exercised although the reference's own metadata never sets it
    rd = jx.BedChunkReaderFromMeta(prefix, row_idx, flip, miss, af, sample_indices=sub)
    assert rd.n_snps == row_idx.size and rd.n_samples == sub.size
    gs, afs, ms, names = [], [], [], []
    while True:
        out = rd.next_chunk_prepared(64)
        if out is None:
            break
        g, sites, a, mi = out
        gs.append(g); afs.append(a); ms.append(mi); names += [s.snp for s in sites]
    g = np.concatenate(gs)
    assert g.shape == (row_idx.size, sub.size) and names == [f"snp{i}" for i in row_idx]
    # row-by-row restatement
    mc = np.clip(af, 0, 1).astype(np.float32)
    alt_mean = (np.float32(2.0) * np.where(flip, np.float32(1.0) - mc, mc)).astype(np.float32)
    codes = np.stack([(case.packed[:, sub // 4] >> ((sub % 4) * 2)) & 3])[0][row_idx]       # [rows, samples]
    lut = np.stack([np.float32(0.0) - alt_mean, np.zeros_like(alt_mean), np.float32(1.0) - alt_mean,
                    np.float32(2.0) - alt_mean], axis=1)
    want = np.take_along_axis(lut, codes.astype(np.int64), axis=1)
    assert np.array_equal(g.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(np.concatenate(afs), alt_mean * np.float32(0.5)) and np.array_equal(np.concatenate(ms), miss)
    with pytest.raises(ValueError, match="sorted in ascending BED order"):
        jx.BedChunkReaderFromMeta(prefix, row_idx[::-1], flip, miss, af)
    with pytest.raises(ValueError, match="additive coding only"):
        jx.BedChunkReaderFromMeta(prefix, row_idx, flip, miss, af, sample_indices=sub).next_chunk_prepared(8, coding="dom")


def test_raw_next_chunk_minor_allele_recoding(jx, oracle, tmp_path):
    """BedChunkReader.next_chunk (src/io/gfreader.rs:3319-3440; process_snp_row, src/io/gfcore.rs:405-480): raw dosages,
    rows with ALT frequency > 0.5 recoded to the other allele (alleles swapped), missing calls filled with the row mean;
    bit-identical to a row-by-row restatement."""
    from janusx_b200 import synth
    case = make_problem(n=130, m=260, q=0, seed=4, missing_rate=0.04)
    n = case.n
    codes0 = np.stack([(case.packed[:, j // 4] >> ((j % 4) * 2)) & 3 for j in range(4 * case.packed.shape[1])], axis=1)
    swap = np.array([3, 1, 2, 0], dtype=np.uint8)       # hom-ref <-> hom-alt: every third row gets ALT frequency > 0.5
    codes0[::3] = swap[codes0[::3]]
    codes0[:, n:] = 0
    packed = np.zeros_like(case.packed)
    for k in range(4):
        packed |= (codes0[:, k::4].astype(np.uint8) << (2 * k))
    prefix = str(tmp_path / "p")
    synth.write_plink(prefix, packed, case.n)
    rd = jx.BedChunkReader(prefix, maf_threshold=0.03, max_missing_rate=0.2)
    gs, names, alleles = [], [], []
    while True:
        out = rd.next_chunk(50)
        if out is None:
            break
        gs.append(out[0]); names += [s.snp for s in out[1]]; alleles += [(s.ref_allele, s.alt_allele) for s in out[1]]
    g = np.concatenate(gs)
    codes = np.stack([(packed[:, j // 4] >> ((j % 4) * 2)) & 3 for j in range(n)], axis=1)
    want_rows, want_names, flips = [], [], 0
    for i in range(packed.shape[0]):
        raw = np.array([0.0, -9.0, 1.0, 2.0], dtype=np.float32)[codes[i]]
        nm = raw >= 0
        non_missing = int(nm.sum())
        if np.float32(1.0 - non_missing / n) > np.float32(0.2) or non_missing == 0:
            continue
        alt_sum = float(raw[nm].astype(np.float64).sum())
        af = alt_sum / (2.0 * non_missing)
        flipped = af > 0.5
        if flipped:
            raw[nm] = np.float32(2.0) - raw[nm]
            alt_sum = 2.0 * non_missing - alt_sum
        if np.float32(min(af, 1.0 - af)) < np.float32(0.03):
            continue
        raw[~nm] = np.float32(alt_sum / non_missing)
        want_rows.append(raw); want_names.append(f"snp{i}"); flips += flipped
        if flipped:
            assert alleles[len(want_rows) - 1] == ("T", "A")
    assert names == want_names and flips > 20
    assert np.array_equal(g.view(np.uint32), np.stack(want_rows).view(np.uint32))


def test_reader_site_selectors(jx, oracle, tmp_path):
    """bim_range / snp_sites / chr_keys / bp_min / bp_max / ranges (src/io/gfreader.rs:125-215, 583-726, 3246-3281): the
    reader walks exactly the rows the same selection given as snp_indices walks, bit for bit."""
    from janusx_b200 import synth
    from janusx_b200.gfreader import BedChunkReader
    n_full, m = 97, 420
    packed, _ = synth.draw_genotypes(m, n_full, seed=23, missing_rate=0.03)
    prefix = str(tmp_path / "sel")
    synth.write_plink(prefix, packed, n_full)
    # three chromosomes, one written with a "chr" prefix
    lines = open(prefix + ".bim").read().splitlines()
    with open(prefix + ".bim", "w") as fh:
        for i, ln in enumerate(lines):
            tok = ln.split("\t")
            tok[0] = ("1", "chr2", "X")[i // 140]
            tok[3] = str(1000 + 10 * (i % 140))
            fh.write("\t".join(tok) + "\n")

    def rows_of(**kw):
        rd = BedChunkReader(prefix, maf_threshold=0.0, max_missing_rate=1.0, **kw)
        if rd.n_snps == 0:
            assert rd.next_chunk_prepared(64) is None
            return 0, None, []
        g, sites, af, miss, _ = _read_all(rd, 64)
        return rd.n_snps, g, [(s.chrom, s.pos, s.snp) for s in sites]

    cases = [
        (dict(bim_range=("chr2", 1100, 1500)), list(range(150, 191))),
        (dict(bim_range=("2", 1100, 1500)), []),                                   # exact chromosome string
        (dict(snp_sites=[("X", 1000), ("1", 1050), ("chr2", 2390)]), [280, 5, 279]),
        (dict(chr_keys=["2"]), list(range(140, 280))),                             # normalised on both sides
        (dict(chr_keys=["CHRX", "1"], bp_min=2300), list(range(130, 140)) + list(range(410, 420))),
        (dict(ranges=[("chr1", 1000, 1040), ("x", 2380, 2390)], bp_max=2385), [0, 1, 2, 3, 4, 418]),
        (dict(snp_range=(100, 200), chr_keys=["chr2"], bp_max=1100), list(range(140, 151))),
    ]
    for kw, want in cases:
        n_snps, g, sites = rows_of(**kw)
        assert n_snps == len(want), kw
        if not want:
            continue
        n2, g2, sites2 = rows_of(snp_indices=want)
        assert sites == sites2 and np.array_equal(g.view(np.uint32), g2.view(np.uint32)), kw
    with pytest.raises(ValueError, match="mmap_window_mb does not support"):
        BedChunkReader(prefix, mmap_window_mb=64, snp_range=(0, 10))
    assert BedChunkReader(prefix, mmap_window_mb=64, chr_keys=["X"]).n_snps == 140
    with pytest.raises(RuntimeError, match="snp site not found"):
        BedChunkReader(prefix, snp_sites=[("1", 7)])


@pytest.mark.parametrize("coding", ["dom", "rec", "het"])
def test_prepared_chunks_non_additive_codings(jx, oracle, tmp_path, coding):
    """next_chunk_prepared(coding=dom|rec|het) (src/io/gfreader.rs:3161-3186, 3632-3653): rows, af (= coded mean) and missing
    counts against the per-sample restatement, on all samples and on a subset, independent of the chunk size."""
    from janusx_b200 import synth
    from janusx_b200.gfreader import BedChunkReader
    n_full, m = 83, 300
    packed, _ = synth.draw_genotypes(m, n_full, seed=31, missing_rate=0.05)
    prefix = str(tmp_path / "cod")
    synth.write_plink(prefix, packed, n_full)
    fam = oracle.read_fam(prefix)
    sub = sorted(np.random.default_rng(3).choice(n_full, size=60, replace=False).tolist())
    for kw, cols in ((dict(), list(range(n_full))), (dict(sample_ids=[fam[i] for i in sub]), sub)):
        n = len(cols)
        codes = np.stack([(packed[:, j // 4] >> ((j % 4) * 2)) & 3 for j in cols], axis=1)
        rd = BedChunkReader(prefix, maf_threshold=0.02, max_missing_rate=0.2, **kw)
        g, sites, af, miss, _ = _read_all(rd, 77, coding=coding)
        g_big = _read_all(BedChunkReader(prefix, maf_threshold=0.02, max_missing_rate=0.2, **kw), 10_000, coding=coding)[0]
        assert np.array_equal(g.view(np.uint32), g_big.view(np.uint32))
        keep, _, _, _ = oracle.bed_chunk_prepared_rows(packed, n_full, None if not kw else np.asarray(sub, dtype=np.int64),
                                                       0.02, 0.2, 1.0)
        kept = np.nonzero(keep)[0]
        assert g.shape == (kept.size, n) and [s.snp for s in sites] == [f"snp{i}" for i in kept]
        one, two, tol = np.float32(1.0), np.float32(2.0), np.float32(1e-6)
        for out_r, i in enumerate(kept):
            c = codes[i]
            nm = c != 1
            imputed = np.float32(float((c == 2).sum() + 2 * (c == 3).sum()) / float(nm.sum()))
            filled = np.array([0.0, imputed, 1.0, 2.0], dtype=np.float32)[c]
            h1, h2 = np.abs(filled - one) <= tol, np.abs(filled - two) <= tol
            coded = np.where({"dom": h1 | h2, "rec": h2, "het": h1}[coding], one, np.float32(0.0)).astype(np.float32)
            cm = np.float32(float(coded.astype(np.float64).sum()) / n)
            assert np.array_equal(g[out_r].view(np.uint32), (coded - cm).view(np.uint32)), (coding, i)
            assert af[out_r].view(np.uint32) == cm.view(np.uint32) and miss[out_r] == float((c == 1).sum())


def test_reader_without_missing_fill(jx, oracle, tmp_path):
    """fill_missing = false (src/io/gfcore.rs:468-476 skipped): next_chunk keeps -9 at missing calls (also in recoded rows),
    next_chunk_prepared carries the marker through coding, mean and centring."""
    from janusx_b200 import synth
    from janusx_b200.gfreader import BedChunkReader
    n, m = 61, 200
    packed, _ = synth.draw_genotypes(m, n, seed=37, missing_rate=0.08)
    packed[3] = 0b01010101                                   # a row with every call missing
    prefix = str(tmp_path / "nofill")
    synth.write_plink(prefix, packed, n)
    codes = np.stack([(packed[:, j // 4] >> ((j % 4) * 2)) & 3 for j in range(n)], axis=1)
    rd = BedChunkReader(prefix, maf_threshold=0.0, max_missing_rate=1.0, fill_missing=False)
    gs, names = [], []
    while True:
        out = rd.next_chunk(64)
        if out is None:
            break
        gs.append(out[0]); names.extend(s.snp for s in out[1])
    g = np.concatenate(gs)
    assert names == [f"snp{i}" for i in range(m)]
    for i in range(m):
        raw = np.array([0.0, -9.0, 1.0, 2.0], dtype=np.float32)[codes[i]]
        nm = raw >= 0
        if nm.any() and float(raw[nm].sum()) / (2.0 * nm.sum()) > 0.5:
            raw[nm] = np.float32(2.0) - raw[nm]
        assert np.array_equal(g[i].view(np.uint32), raw.view(np.uint32)), i
    for coding in ("add", "rec"):
        rd = BedChunkReader(prefix, maf_threshold=0.0, max_missing_rate=1.0, fill_missing=False)
        gp, sites, af, miss, _ = _read_all(rd, 50, coding=coding)
        assert gp.shape == (m, n)
        for i in range(m):
            raw = np.array([0.0, -9.0, 1.0, 2.0], dtype=np.float32)[codes[i]]
            coded = raw if coding == "add" else np.where(np.abs(raw - np.float32(2.0)) <= np.float32(1e-6), np.float32(1.0),
                                                         np.float32(0.0)).astype(np.float32)
            cm = np.float32(float(coded.astype(np.float64).sum()) / n)
            assert np.array_equal(gp[i].view(np.uint32), (coded - cm).view(np.uint32)), (coding, i)
            assert af[i].view(np.uint32) == (cm * np.float32(0.5) if coding == "add" else cm).view(np.uint32)
            assert miss[i] == float((codes[i] == 1).sum())
