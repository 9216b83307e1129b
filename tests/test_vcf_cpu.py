"""VCF -> PLINK conversion (host only): the library's converter against a line-by-line Python
restatement of the reference's reader (VcfSnpIter::next_snp_raw, src/io/gfcore.rs:2875-2980;
plink2bits_from_g_f32, src/io/gfreader.rs:2630-2641)."""
import gzip
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
CODE = {"0/0": 0, "0|0": 0, "0/1": 2, "1/0": 2, "0|1": 2, "1|0": 2, "1/1": 3, "1|1": 3}   # else 1 = missing


def restate(vcf_path, snps_only=False):
    """-> (samples, bim rows, packed u8[m, ceil(n/4)])"""
    op = gzip.open if str(vcf_path).endswith(".gz") else open
    samples, bim, rows = None, [], []
    with op(vcf_path, "rt") as fh:
        for line in fh:
            if line.startswith("#CHROM"):
                samples = line.rstrip().split("\t")[9:]
                continue
            if line.startswith("#") or not line.strip():
                continue
            parts = line.rstrip().split("\t")
            if len(parts) < 10 or "GT" not in parts[8].split(":"):
                continue
            ref, alt = parts[3], parts[4]
            if snps_only and not all(len(a.strip()) == 1 and a.strip().upper() in "ACGT" for a in (ref, alt)):
                continue
            try:
                pos = int(parts[1])
            except ValueError:
                pos = 0
            snp = parts[2] if parts[2].strip() and parts[2] != "." else f"{parts[0]}_{parts[1]}"
            bim.append(f"{parts[0]}\t{snp}\t0\t{pos}\t{ref}\t{alt}\n")
            codes = np.ones(len(samples), dtype=np.uint8)
            for j, f in enumerate(parts[9:9 + len(samples)]):
                codes[j] = CODE.get(f.split(":")[0], 1)
            pad = np.zeros((len(samples) + 3) // 4 * 4, dtype=np.uint8)
            pad[: len(samples)] = codes
            q = pad.reshape(-1, 4)
            rows.append(q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6))
    return samples, bim, np.array(rows, dtype=np.uint8)


def check(prefix, vcf, snps_only):
    samples, bim, packed = restate(vcf, snps_only)
    raw = np.fromfile(str(prefix) + ".bed", dtype=np.uint8)
    assert bytes(raw[:3]) == b"\x6c\x1b\x01"
    assert np.array_equal(raw[3:].reshape(len(bim), -1), packed)
    assert open(str(prefix) + ".bim").readlines() == bim
    assert open(str(prefix) + ".fam").readlines() == [f"{s}\t{s}\t0\t0\t0\t-9\n" for s in samples]
    return len(samples), len(bim)


def test_edge_case_vcf_plain_and_gz(tmp_path):
    from janusx_b200 import jxrs
    text = "\n".join([
        "##fileformat=VCFv4.2",
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\ts3\ts4\ts5",
        "1\t100\trs1\tA\tG\t.\t.\t.\tGT\t0/0\t0/1\t1/1\t./.\t1|0",
        "1\t200\t.\tC\tT\t.\t.\t.\tGT:DP\t0|0:12\t1|1:3\t0/1:9\t1/2:4\t1:7",          # multi-allelic / haploid -> missing
        "2\tabc\t\tG\tGA\t.\t.\t.\tGT\t0/0\t0/0\t0/1\t0/1\t1/1",                        # bad POS -> 0, indel
        "2\t300\trs4\tT\tC\t.\t.\t.\tDP\t1\t2\t3\t4\t5",                                # no GT in FORMAT: skipped
        "2\t400\trs5\tT\tC",                                                              # short line: skipped
        "",
        "X\t500\trs6\tt\tc\t.\t.\t.\tGT\t.|.\t0/1\t0/1\t0/0\t1/1",                        # lower-case alleles count as SNP
    ]) + "\n"
    plain = tmp_path / "e.vcf"
    plain.write_text(text)
    gz = tmp_path / "e.vcf.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(text)
    for src, so, want_sites in ((plain, False, 4), (gz, False, 4), (gz, True, 3)):
        prefix = tmp_path / f"out_{src.suffix}_{int(so)}"
        ns, nv = jxrs.vcf_to_plink(str(src), str(prefix), so)
        assert (ns, nv) == (5, want_sites) == check(prefix, src, so)
    bim = open(str(tmp_path / "out_.vcf_0") + ".bim").readlines()
    assert bim[1].split("\t")[1] == "1_200" and bim[2].split("\t")[1] == "2_abc" and bim[2].split("\t")[3] == "0"
    with pytest.raises(RuntimeError, match="No such file"):
        jxrs.vcf_to_plink(str(tmp_path / "missing.vcf"), str(tmp_path / "x"))
    (tmp_path / "nohdr.vcf").write_text("1\t1\t.\tA\tC\t.\t.\t.\tGT\t0/0\n")
    with pytest.raises(RuntimeError, match="No #CHROM header"):
        jxrs.vcf_to_plink(str(tmp_path / "nohdr.vcf"), str(tmp_path / "x"))


def test_mouse_fixture_conversion(tmp_path):
    """BASELINE.json configs[0] input (first 1,500 records of the reference's example VCF)."""
    from janusx_b200 import jxrs
    vcf = GOLDEN / "mouse_hs1940_sub.vcf.gz"
    ns, nv = jxrs.vcf_to_plink(str(vcf), str(tmp_path / "mouse"), False)
    assert (ns, nv) == (1940, 1500) == check(tmp_path / "mouse", vcf, False)
    ns2, nv2 = jxrs.vcf_to_plink(str(vcf), str(tmp_path / "mouse_snp"), True)
    assert ns2 == 1940 and nv2 <= nv and (ns2, nv2) == check(tmp_path / "mouse_snp", vcf, True)
