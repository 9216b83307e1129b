"""CPU-only checks of the drop-in boundary: the library loads, exports every declared symbol, refuses to
compute without a GPU, and its host-side pieces (TSV formatter, argument validation) match the oracle."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def cabi():
    from janusx_b200 import _cabi
    from janusx_b200 import build as jb
    jb.build()
    return _cabi


def test_header_symbols_all_exported(cabi):
    hdr = (ROOT / "include" / "jxb200.h").read_text()
    declared = set(re.findall(r"\b(jxb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(str(cabi.LIB_PATH))
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(cabi.SYMBOLS), declared ^ set(cabi.SYMBOLS)
    assert b"sm_100a" in cabi.lib().jxb_build_info()


def test_no_cpu_fallback(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from janusx_b200 import jxrs
    with pytest.raises(cabi.JxbError, match="no CPU fallback"):
        jxrs.DeviceModel(np.ones(4), np.ones((4, 1)), np.zeros(4))
    h = C.c_void_p()
    s = np.ones(4)
    rc = cabi.lib().jxb_model_create(0, 4, 1, s.ctypes.data, s.ctypes.data, s.ctypes.data, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in cabi.lib().jxb_last_error()


def test_product_never_imports_oracle():
    for path in (ROOT / "janusx_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"}:
            txt = path.read_text()
            assert "oracle" not in txt.replace("oracle-free", ""), f"{path} mentions the oracle"


def test_format_row_matches_oracle(cabi, oracle):
    rng = np.random.default_rng(0)
    rows = [
        [0.5, 0.25, 6.1e-5], [float("nan"), float("nan"), 1.0], [1.0, 0.0, 0.3], [-3.25, 1e-3, 0.0],
        [1.2345678, 0.5, 1e-320, 0.5], [0.1, 0.2, 0.3, 2.5, -1234.5678, 3e-300], [1e10, 1e-10, 1.0, float("nan"), float("inf"), 1.0],
    ]
    for _ in range(200):
        k = int(rng.choice([3, 4, 6]))
        rows.append(list(rng.normal(size=k) * 10.0 ** rng.integers(-8, 8, size=k)))
    buf = C.create_string_buffer(4096)
    for i, r in enumerate(rows):
        a = np.ascontiguousarray(r, dtype=np.float64)
        a[1] = abs(a[1]) if i % 5 else a[1]
        a[2] = abs(a[2])
        snp = "." if i % 3 == 0 else f"rs{i}"
        af, mr = np.float32(rng.random()), np.float32(rng.random() * 0.1)
        n = cabi.lib().jxb_format_row(buf, 4096, b"7", 1000 + i, snp.encode(), b"A", b"TG", af, mr,
                                      a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0])
        assert buf.raw[:n] == oracle.format_row("7", 1000 + i, snp, "A", "TG", float(af), float(mr), a), (i, r)


def test_tsv_formatter_long_alleles_never_overflow(cabi, oracle):
    """BIM alleles have no length limit (indels, SVs; vcf_cache.cpp copies REF/ALT verbatim) and Rust formats rows into a
    String.  A buffer below the worst case is refused with the size to retry with; a big one holds the full row."""
    a = np.array([0.5, 0.25, 6.1e-5], dtype=np.float64)
    short = oracle.format_row("7", 12, "rs1", "A", "T", 0.25, 0.5, a)
    for alen in (1900, 3000, 5000, 70000):
        allele = ("ACGT" * (alen // 4 + 1))[:alen].encode()
        small = C.create_string_buffer(2048)
        need = cabi.lib().jxb_format_row(small, 2048, b"7", 12, b"rs1", b"A", allele, np.float32(0.25), np.float32(0.5),
                                         a.ctypes.data_as(C.POINTER(C.c_double)), 3)
        assert need > 2048
        big = C.create_string_buffer(need)
        n = cabi.lib().jxb_format_row(big, need, b"7", 12, b"rs1", b"A", allele, np.float32(0.25), np.float32(0.5),
                                      a.ctypes.data_as(C.POINTER(C.c_double)), 3)
        assert n <= need
        assert big.raw[:n] == short.replace(b"\tA\tT\t", b"\tA\t" + allele + b"\t")
    # huge finite values: `{:.4}` of 1e300 is 306 characters per field
    h = np.array([1e300, 1e300, 0.5, 1e300, -1e300, 0.5], dtype=np.float64)
    buf = C.create_string_buffer(8192)
    n = cabi.lib().jxb_format_row(buf, 8192, b"1", 1, b"s", b"A", b"C", np.float32(0.1), np.float32(0.0),
                                  h.ctypes.data_as(C.POINTER(C.c_double)), 6)
    f = buf.raw[:n].split(b"\t")
    assert len(f) == 14 and f[7] == (b"%.4f" % 1e300) and f[-1] == b"5.0000e-1\n"


def test_tsv_writer_blocks_match_oracle_rows(cabi, oracle, tmp_path):
    """GwasAssocTsvWriter (assoc2tsv.rs:765-892): header by schema, verbatim SNP names, allele strings by model."""
    from janusx_b200 import jxrs
    rng = np.random.default_rng(1)
    m = 300
    sites = [jxrs.SiteInfo(str(1 + i % 5), 100 + i, "ACGT"[i % 4], "TGCA"[i % 4] + ("C" if i % 11 == 0 else "")) for i in range(m)]
    snp = ["." if i % 7 == 0 else f"rs{i}" for i in range(m)]
    maf = rng.random(m).astype(np.float32)
    miss = (rng.random(m) * 0.05).astype(np.float32)
    for cols, model in ((3, "add"), (4, "dom"), (6, "rec"), (3, "het")):
        res = rng.normal(size=(m, cols)) * 10.0 ** rng.integers(-6, 6, size=(m, cols))
        res[:, 1:3] = np.abs(res[:, 1:3])
        res[5, :2] = np.nan
        path = tmp_path / f"w{cols}{model}.tsv"
        w = jxrs.GwasAssocTsvWriter(str(path), model)
        assert w.write_chunk(sites[:100], snp[:100], maf[:100], miss[:100], res[:100]) == 100
        assert w.write_chunk(sites[100:], snp[100:], maf[100:], miss[100:], res[100:]) == m - 100
        with pytest.raises(ValueError, match="inconsistent results columns"):
            w.write_chunk(sites[:1], snp[:1], maf[:1], miss[:1], np.zeros((1, 4 if cols != 4 else 3)))
        with pytest.raises(ValueError, match="snp length mismatch"):
            w.write_chunk(sites[:2], snp[:1], maf[:2], miss[:2], res[:2])
        w.close()
        assert w.rows_written == m
        lines = path.read_bytes().split(b"\n")
        assert lines[0].split(b"\t")[:7] == [b"chrom", b"pos", b"snp", b"allele0", b"allele1", b"af", b"miss"]
        assert len(lines[0].split(b"\t")) == 8 + cols and len(lines) == m + 2
        for i in (0, 5, 7, 11, 123, m - 1):
            r, a = sites[i].ref_allele, sites[i].alt_allele
            a0, a1 = {"add": (r, a), "dom": (r + r, r + a + "/" + a + a), "rec": (r + a + "/" + r + r, a + a),
                      "het": (r + r + "/" + a + a, r + a)}[model]
            want = oracle.format_row(sites[i].chrom, sites[i].pos, snp[i] if snp[i] != "." else "\x01", a0, a1,
                                     float(maf[i]), float(miss[i]), res[i]).replace(b"\x01", b".")
            assert lines[1 + i] + b"\n" == want, (model, i)
    with pytest.raises(ValueError, match="genetic_model must be one of"):
        jxrs.GwasAssocTsvWriter(str(tmp_path / "x.tsv"), "mult")


def test_argument_validation_messages():
    from janusx_b200 import jxrs
    n = 8
    s, x, y, ut = np.ones(n), np.ones((n, 2)), np.zeros(n), np.eye(n, dtype=np.float32)
    with pytest.raises(RuntimeError, match="low must be < high"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut, 0.02, 0.05, 1.0, low=1.0, high=1.0)
    with pytest.raises(RuntimeError, match="tol must be positive and finite"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut, 0.02, 0.05, 1.0, tol=0.0)
    with pytest.raises(RuntimeError, match=r"u_t must be \(n, n\) row-major U\^T"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut[:, :4], 0.02, 0.05, 1.0)
    with pytest.raises(RuntimeError, match="model must be one of: add, dom, rec, het"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut, 0.02, 0.05, 1.0, genetic_model="xyz")
    with pytest.raises(RuntimeError, match="prepared row metadata must provide all or none"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut, 0.02, 0.05, 1.0, row_indices=np.arange(3))
    with pytest.raises(RuntimeError, match="sorted in ascending BED order"):
        jxrs.lmm_reml_assoc_bed_to_tsv_f32("p", "o", s, x, y, ut, 0.02, 0.05, 1.0, row_indices=np.array([3, 1]),
                                           row_flip=np.zeros(2, bool), row_missing=np.zeros(2, np.float32),
                                           row_maf=np.zeros(2, np.float32))
    with pytest.raises(RuntimeError, match=r"x rows must equal len\(y\)"):
        jxrs.lmm_rotate_x_y_with_ut_f64(ut, np.ones((n - 1, 2)), y)
    with pytest.raises(RuntimeError, match="invalid log10_lbd"):
        jxrs.fvlmm_assoc_bed_to_tsv_f32("p", "o", s, x, y, float("nan"), ut, 0.02, 0.05, 1.0)


def test_lmm_lm_null_lrt_decision_host_logic():
    """src/stats/gwas_unified.rs:119-175: LM null ML closed form + boundary-mixture LRT."""
    import math
    from janusx_b200 import jxrs
    rng = np.random.default_rng(3)
    n = 300
    x = rng.normal(size=(n, 2))
    y = 1.0 + x @ np.array([0.5, -0.25]) + rng.normal(size=n)
    design = np.concatenate([np.ones((n, 1)), x], axis=1)
    rss = float(np.sum((y - design @ np.linalg.lstsq(design, y, rcond=None)[0]) ** 2))
    lm_ml0 = n * (math.log(n) - 1 - math.log(2 * math.pi)) / 2 - 0.5 * n * math.log(rss)
    sw, stat, p, got = jxrs.gwas_lmm_lm_null_lrt_decision(y, x, lm_ml0 + 0.1)
    assert math.isclose(got, lm_ml0, rel_tol=1e-12) and math.isclose(stat, 0.2, rel_tol=1e-9)
    assert math.isclose(p, 0.5 * math.erfc(math.sqrt(0.1)), rel_tol=1e-12) and sw is True
    sw2, stat2, p2, _ = jxrs.gwas_lmm_lm_null_lrt_decision(y, x, lm_ml0 + 10.0)
    assert sw2 is False and p2 < 1e-5
    sw3, stat3, p3, _ = jxrs.gwas_lmm_lm_null_lrt_decision(y, x, lm_ml0 - 5.0, boundary_mixture=False)
    assert stat3 == 0.0 and p3 == 1.0 and sw3 is True
    with pytest.raises(RuntimeError, match="alpha must be in"):
        jxrs.gwas_lmm_lm_null_lrt_decision(y, x, 0.0, alpha=1.5)
    with pytest.raises(RuntimeError, match="insufficient samples"):
        jxrs.gwas_lmm_lm_null_lrt_decision(y[:3], x[:3], 0.0)


def test_cli_argument_rules_and_vcf_cache(tmp_path, monkeypatch):
    """Host-side CLI logic that needs no GPU: flag validation and the VCF -> PLINK cache (rebuilt only when stale)."""
    from janusx_b200 import gwas, jxrs
    with pytest.raises(SystemExit):
        gwas.parse_args(["-p", "pheno.tsv", "-lmm"])                              # neither -bfile nor -vcf
    with pytest.raises(SystemExit):
        gwas.parse_args(["-bfile", "a", "-vcf", "b.vcf", "-p", "pheno.tsv", "-lmm"])  # both
    with pytest.raises(SystemExit):
        gwas.parse_args(["-bfile", "a", "-p", "pheno.tsv"])                       # no model selected
    a = gwas.parse_args(["-vcf", "x.vcf.gz", "-p", "p.tsv", "-lmm2", "-maf", "0.05", "-snps-only"])
    assert a.vcf == "x.vcf.gz" and a.lmm2 and not a.lmm and a.maf == 0.05 and a.snps_only and a.grm == "1"
    vcf = tmp_path / "toy.vcf"
    vcf.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n1\t10\trs1\tA\tG\t.\t.\t.\tGT\t0/1\t1/1\n")
    calls = []
    real = jxrs.vcf_to_plink
    monkeypatch.setattr(jxrs, "vcf_to_plink", lambda *args: (calls.append(args), real(*args))[1])
    prefix = gwas._vcf_cache(str(vcf), str(tmp_path), False)
    assert prefix.endswith("~toy.snp0") and len(calls) == 1 and Path(prefix + ".bed").read_bytes() == b"\x6c\x1b\x01\x0e"
    assert gwas._vcf_cache(str(vcf), str(tmp_path), False) == prefix and len(calls) == 1      # fresh: reused
    import os
    os.utime(vcf, (os.path.getmtime(prefix + ".bed") + 10,) * 2)
    gwas._vcf_cache(str(vcf), str(tmp_path), False)
    assert len(calls) == 2                                                                     # stale: rebuilt
    assert jxrs.default_device_batch(20000) == 151552 and jxrs.default_device_batch(50000) == 75776


def test_packed_entry_point_validation_and_writer_text_paths(tmp_path):
    """Argument checks of lmm_reml_assoc_packed_f32 (src/stats/lmm.rs:3089-3187: same messages, raised before any device
    work) and the raw-text paths of GwasAssocTsvWriter (append_text / send_block / flush, assoc2tsv.rs:862-885)."""
    from janusx_b200 import jxrs
    n, m = 10, 4
    bps = (n + 3) // 4
    packed = np.zeros((m, bps), np.uint8)
    s, xc, y, ut = np.ones(n), np.ones((n, 1)), np.zeros(n), np.eye(n, dtype=np.float32)
    flip, maf = np.zeros(m, bool), np.zeros(m, np.float32)
    call = lambda **kw: jxrs.lmm_reml_assoc_packed_f32(**{**dict(packed=packed, n_samples=n, row_flip=flip, row_maf=maf, s=s,
                                                                 xcov=xc, y_rot=y, u_t=ut), **kw})
    for kw, msg in ((dict(n_samples=0), "n_samples must be > 0"), (dict(low=1.0, high=1.0), "low must be < high"),
                    (dict(tol=0.0), "tol must be positive and finite"), (dict(packed=np.zeros(5, np.uint8)), "packed must be 2D"),
                    (dict(packed=np.zeros((m, bps + 1), np.uint8)), "packed second dimension mismatch: got 4, expected 3"),
                    (dict(row_maf=maf[:-1]), "row_flip/row_maf length mismatch"),
                    (dict(row_indices=[0, 9]), "row_indices out of range"),
                    (dict(sample_indices=[0, 1]), "sample_indices length mismatch: got 2, expected 10"),
                    (dict(n_samples=12, packed=np.zeros((m, 3), np.uint8)), "must equal n_samples=12 when sample_indices is not provided"),
                    (dict(u_t=np.eye(n - 1, dtype=np.float32)), "u_t must be"), (dict(model="mult"), "model must be one of")):
        with pytest.raises((RuntimeError, ValueError), match=msg):
            call(**kw)
    w = jxrs.GwasAssocTsvWriter(str(tmp_path / "t.tsv"))
    w.append_text("", True, 0)                                  # no-op: nothing opened yet
    with pytest.raises(IOError):
        w.send_block(b"x")                                      # writer not initialised
    w.append_text("1\t5\trs1\tA\tG\t0.1000\t0.0000\t1.0000\t0.5000\t4.0000e0\t4.5500e-2\t5.0000e-2\n", True, 1)
    w.send_block(b"# trailer\n")
    w.flush()
    with pytest.raises(ValueError, match="inconsistent results columns"):
        w.append_text("x\n", False, 1)
    w.close()
    w.close()
    lines = (tmp_path / "t.tsv").read_text().splitlines()
    assert lines[0].endswith("pwald\tplrt") and lines[1].startswith("1\t5\trs1") and lines[2] == "# trailer" and w.rows_written == 1
    cols = jxrs._read_bim_columns.__wrapped__ if hasattr(jxrs._read_bim_columns, "__wrapped__") else jxrs._read_bim_columns
    (tmp_path / "p.bim").write_text("1\trs1\t0\t100\tA\tG\n2\t.\t0\tx\tC\tT\n")
    chrom, pos, snp, a0, a1 = cols(str(tmp_path / "p"), None)
    assert (chrom, pos, snp, a0, a1) == (["1", "2"], [100, 0], ["rs1", "."], ["A", "C"], ["G", "T"])
    assert cols(str(tmp_path / "p"), [1])[2] == ["."]


def test_prepared_meta_row_decisions_host_logic():
    """prepare_bed_logic_meta_selected's row closure (src/io/gfreader.rs:5378-5424) on integer counts: f32 missing rate and
    ALT frequency, f64 het-rate comparison, `non_missing == 0` kept only without a MAF threshold."""
    from janusx_b200.gfreader import _meta_row_decisions
    rng = np.random.default_rng(0)
    n = 137
    miss, het, hom = rng.integers(0, 40, 500), rng.integers(0, 60, 500), rng.integers(0, 37, 500)
    miss[::50], het[::50], hom[::50] = n, 0, 0
    for maf_t, miss_t, het_t in ((0.05, 0.2, 0.4), (0.0, 1.0, 0.0), (0.02, 0.05, 1.0)):
        k, mr, af = _meta_row_decisions(miss, het, hom, n, maf_t, miss_t, het_t)
        for i in range(500):
            nm, alt = n - miss[i], het[i] + 2 * hom[i]
            mr_i = np.float32(n - nm) / np.float32(n)
            af_i = np.float32(alt) / (np.float32(2.0) * np.float32(nm)) if nm > 0 else np.float32(0)
            if mr_i > np.float32(miss_t):
                keep = False
            elif nm == 0:
                keep = np.float32(maf_t) <= 0
            elif np.float32(het_t) > 0 and (het[i] / nm) > float(np.float32(het_t)):
                keep = False
            else:
                keep = min(af_i, np.float32(1) - af_i) >= np.float32(maf_t)
            assert keep == k[i] and mr_i == mr[i] and af_i == af[i], (i, maf_t)


def test_number_formatters_equal_printf():
    """The TSV writer's exact fast path for `{:.4}` / `{:.4e}` / `{:.6e}` (src/io/assoc2tsv.rs:430-517 prints with Rust's
    correctly rounded formatter) must print what the printf route prints: ties and near-ties at the printed precision,
    neighbours of powers of ten, carries into the next decade, f32 values, either sign."""
    import ctypes as C
    from janusx_b200 import _cabi
    lib = _cabi.lib()
    for prec in (4, 6):
        msg = C.create_string_buffer(256)
        bad = lib.jxb_selftest_format(1_500_000, 20260609 + prec, prec, msg, 256)
        assert bad == 0, msg.value.decode()
    # spot values through the row formatter: a tie (0.03125 -> 312.5e-4), a negative that rounds to zero, a carry, a tiny p
    buf = C.create_string_buffer(4096)
    row = (C.c_double * 3)(-0.00004, 9.99996, 3.2e-301)
    n = lib.jxb_format_row(buf, 4096, b"7", -12, b".", b"A", b"G", C.c_float(0.03125), C.c_float(0.5), row, 3)
    want = "7\t-12\t7_-12\tA\tG\t%.4f\t%.4f\t%.4f\t%.4f\t" % (0.03125, 0.5, -0.00004, 9.99996)
    z = (-0.00004 / 9.99996) ** 2
    m, e = ("%.4e" % z).split("e")
    want += f"{m}e{int(e)}\t3.2000e-301\n"
    assert buf.raw[:n].decode() == want
