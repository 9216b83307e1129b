"""Row selection of BedChunkReader (host logic, src/io/gfreader.rs:125-215, 583-726, 3246-3281): the four primary
selectors, the site filter applied on top, the reference's error messages."""
import numpy as np
import pytest

from janusx_b200.gfreader import normalize_chr_key, select_snp_rows

CHROMS = ["1", "1", "chr2", "2", "X", "1", " Chr3 ", "3"]
POS = [10, 20, 5, 30, 7, 20, 11, 12]


def test_chr_key_normalisation():
    assert normalize_chr_key("chr1") == "1" and normalize_chr_key(" CHRx ") == "X" and normalize_chr_key("Chr 7") == "7"
    assert normalize_chr_key("mt") == "MT" and normalize_chr_key("") == "" and normalize_chr_key("ch1") == "CH1"


def test_primary_selectors():
    assert select_snp_rows(CHROMS, POS) is None
    assert select_snp_rows(CHROMS, POS, snp_range=(1, 4)).tolist() == [1, 2, 3]
    assert select_snp_rows(CHROMS, POS, snp_indices=[5, 0, 7]).tolist() == [5, 0, 7]          # order kept
    assert select_snp_rows(CHROMS, POS, bim_range=("1", 10, 20)).tolist() == [0, 1, 5]         # closed interval
    assert select_snp_rows(CHROMS, POS, bim_range=("2", 0, 100)).tolist() == [3]               # exact chromosome string
    assert select_snp_rows(CHROMS, POS, bim_range=("9", 0, 100)).tolist() == []                # empty selection is allowed
    assert select_snp_rows(CHROMS, POS, snp_sites=[("1", 20), ("X", 7)]).tolist() == [1, 5, 4]  # key order, duplicates of a key


@pytest.mark.parametrize("kw, msg", [
    (dict(snp_range=(0, 2), snp_indices=[1]), "Provide only one of snp_range, snp_indices, bim_range, or snp_sites"),
    (dict(bim_range=("1", 0, 5), snp_sites=[("1", 10)]), "Provide only one of"),
    (dict(snp_range=(3, 3)), r"invalid snp_range: \(3, 3\)"),
    (dict(snp_range=(0, 9)), r"invalid snp_range: \(0, 9\)"),
    (dict(snp_indices=[]), "snp_indices is empty"),
    (dict(snp_indices=[8]), "snp index out of range: 8"),
    (dict(snp_indices=[2, 2]), "duplicate snp index: 2"),
    (dict(bim_range=("1", 5, 4)), "bim_range start > end"),
    (dict(snp_sites=[]), "snp_sites is empty"),
    (dict(snp_sites=[("1", 11)]), r"snp site not found: \(1, 11\)"),
    (dict(bp_min=9, bp_max=8), "bp_min cannot be greater than bp_max"),
    (dict(ranges=[("1", 9, 8)]), "One range has start > end"),
])
def test_selector_errors(kw, msg):
    with pytest.raises(RuntimeError, match=msg):
        select_snp_rows(CHROMS, POS, **kw)


def test_site_filter_on_top():
    # every given condition must hold; chromosome keys are normalised on both sides
    assert select_snp_rows(CHROMS, POS, chr_keys=["CHR2"]).tolist() == [2, 3]
    assert select_snp_rows(CHROMS, POS, chr_keys=["chr2", "3"], bp_min=6).tolist() == [3, 6, 7]
    assert select_snp_rows(CHROMS, POS, bp_min=10, bp_max=12).tolist() == [0, 6, 7]
    assert select_snp_rows(CHROMS, POS, ranges=[("chr1", 15, 25), ("x", 0, 10)]).tolist() == [1, 4, 5]   # union of ranges
    assert select_snp_rows(CHROMS, POS, ranges=[("1", 0, 100)], chr_keys=["2"]).tolist() == []
    # applied to what the primary selector left, order kept
    assert select_snp_rows(CHROMS, POS, snp_indices=[5, 4, 0], ranges=[("chr1", 15, 25), ("x", 0, 10)]).tolist() == [5, 4]
    assert select_snp_rows(CHROMS, POS, snp_range=(0, 6), chr_keys=["1"], bp_max=10).tolist() == [0]
    # empty key lists / range lists switch their condition off
    assert select_snp_rows(CHROMS, POS, chr_keys=[" "], ranges=[]) is None
    assert select_snp_rows(CHROMS, POS, snp_range=(2, 4), chr_keys=[]).tolist() == [2, 3]


@pytest.mark.parametrize("coding", ["dom", "rec", "het"])
def test_coded_row_lut_equals_row_by_row_restatement(coding):
    """Non-additive codings of next_chunk_prepared (src/io/gfreader.rs:3161-3186, 3632-3653): the four values per row
    against the reference's per-sample loop (map with 1e-6 tolerance, f64 sum, f32 mean, f32 centring)."""
    from janusx_b200.gfreader import coded_row_lut, prepared_row_decisions
    rng = np.random.default_rng(5)
    n = 37
    rows = [rng.choice(4, size=n, p=[0.5, 0.1, 0.3, 0.1]) for _ in range(40)]
    rows.append(np.array([2] * 30 + [1] * 7))          # every call het or missing: imputed == 1.0 exactly
    rows.append(np.array([3] * 30 + [1] * 7))          # every call hom-alt or missing: imputed == 2.0 exactly
    rows.append(np.array([2] * 18 + [3] * 18 + [1]))   # imputed 1.5
    codes = np.stack(rows)
    missing, het, hom = (codes == 1).sum(1), (codes == 2).sum(1), (codes == 3).sum(1)
    keep, imputed = prepared_row_decisions(missing, het, hom, n, 0.0, 1.0, 1.0)
    lut, mean = coded_row_lut(missing, het, hom, imputed, n, coding)
    one, two, tol = np.float32(1.0), np.float32(2.0), np.float32(1e-6)
    for r in range(codes.shape[0]):
        filled = np.array([0.0, imputed[r], 1.0, 2.0], dtype=np.float32)[codes[r]]
        total, coded = 0.0, np.empty(n, dtype=np.float32)
        for j, v in enumerate(filled):
            h1, h2 = abs(np.float32(v - one)) <= tol, abs(np.float32(v - two)) <= tol
            mv = np.float32(1.0 if {"dom": h1 or h2, "rec": h2, "het": h1}[coding] else 0.0)
            coded[j] = mv
            total += float(mv)
        cm = np.float32(total / n)
        want = coded - cm
        assert mean[r].view(np.uint32) == cm.view(np.uint32)
        assert np.array_equal(lut[r][codes[r]].view(np.uint32), want.view(np.uint32)), (coding, r)


@pytest.mark.parametrize("coding", ["add", "dom", "het"])
def test_coded_row_lut_unfilled_rows(coding):
    """fill_missing = false: missing calls keep the decoder's -9 through the coding map, the mean and the centring."""
    from janusx_b200.gfreader import coded_row_lut
    rng = np.random.default_rng(6)
    n = 29
    codes = np.stack([rng.choice(4, size=n, p=[0.45, 0.15, 0.3, 0.1]) for _ in range(30)] + [np.ones(n, dtype=np.int64)])
    missing, het, hom = (codes == 1).sum(1), (codes == 2).sum(1), (codes == 3).sum(1)
    lut, mean = coded_row_lut(missing, het, hom, np.full(codes.shape[0], -9.0, dtype=np.float32), n, coding)
    one, two, tol = np.float32(1.0), np.float32(2.0), np.float32(1e-6)
    for r in range(codes.shape[0]):
        raw = np.array([0.0, -9.0, 1.0, 2.0], dtype=np.float32)[codes[r]]
        if coding == "add":
            coded = raw.copy()
        else:
            h1, h2 = np.abs(raw - one) <= tol, np.abs(raw - two) <= tol
            coded = np.where({"dom": h1 | h2, "het": h1}[coding], one, np.float32(0.0)).astype(np.float32)
        total = 0.0
        for v in coded:
            total += float(v)
        cm = np.float32(total / n)
        assert mean[r].view(np.uint32) == cm.view(np.uint32)
        assert np.array_equal(lut[r][codes[r]].view(np.uint32), (coded - cm).view(np.uint32)), (coding, r)
