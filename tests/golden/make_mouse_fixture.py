"""Derive the real-data fixture of BASELINE.json configs[0] from the reference's example files.

Reads   /root/reference/example/mouse_hs1940.vcf.gz  (1,940 heterogeneous-stock mice, 10,300 sites, GT only)
        /root/reference/example/mouse_hs1940.pheno    (6 traits, NA = missing)
Writes  tests/golden/mouse_hs1940_full.vcf.gz         (all 10,300 records: the input of configs[0] as BASELINE.json names it)
        tests/golden/mouse_hs1940_sub.vcf.gz          (header + the first 1,500 records, bytes unchanged; fast CLI test)
        tests/golden/mouse_hs1940_sub.pheno           (sample id + traits test0, test3)
The reference holds no expected association output for this data set, so the fixture supplies real
genotype structure (relatedness, real allele-frequency spectrum) as INPUT; expectations come from the oracle at test
time.  /root/reference does not exist on the GPU box, hence the committed copy.

Run:  python tests/golden/make_mouse_fixture.py
"""
import gzip
from pathlib import Path

SRC = Path("/root/reference/example")
OUT = Path(__file__).resolve().parent
N_RECORDS = 1500


def main():
    import shutil
    shutil.copyfile(SRC / "mouse_hs1940.vcf.gz", OUT / "mouse_hs1940_full.vcf.gz")
    kept, records = [], 0
    with gzip.open(SRC / "mouse_hs1940.vcf.gz", "rt") as fh:
        for line in fh:
            if line.startswith("#"):
                kept.append(line)
                continue
            kept.append(line)
            records += 1
            if records >= N_RECORDS:
                break
    with gzip.GzipFile(OUT / "mouse_hs1940_sub.vcf.gz", "wb", compresslevel=9, mtime=0) as fh:
        fh.write("".join(kept).encode())
    with open(SRC / "mouse_hs1940.pheno") as fh, open(OUT / "mouse_hs1940_sub.pheno", "w") as out:
        header = fh.readline().rstrip("\n").split("\t")
        cols = [header.index("test0"), header.index("test3")]
        out.write("id\t" + "\t".join(header[c] for c in cols) + "\n")
        for line in fh:
            tok = line.rstrip("\n").split("\t")
            out.write(tok[0] + "\t" + "\t".join(tok[c] for c in cols) + "\n")
    print("records", records)


if __name__ == "__main__":
    main()
