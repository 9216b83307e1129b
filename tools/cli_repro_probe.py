"""CLI reproducibility probe: the same job as 1 rank (twice) and as 2 ranks; reports differing TSV lines."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import make_problem  # noqa: E402

from janusx_b200 import synth  # noqa: E402


def run(args):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "janusx_b200.gwas", *args], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if r.returncode:
        print(r.stdout)
        raise SystemExit(r.returncode)
    return r.stdout


def main():
    tmp = Path(tempfile.mkdtemp())
    case = make_problem(n=400, m=3000, q=0, seed=123, missing_rate=0.02)
    prefix = str(tmp / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    with open(tmp / "pheno.tsv", "w") as fh:
        fh.write("id\ttraitA\n")
        for j in range(case.n):
            fh.write(f"S{j}\t{case.y[j]:.10f}\n")
    outs = {}
    for tag, g in (("a1", 1), ("b1", 1), ("c2", 2)):
        out = tmp / tag
        log = run(["-bfile", prefix, "-p", str(tmp / "pheno.tsv"), "-lmm", "-k", "1", "-q", "2", "-force-model", "-gpus", str(g),
                   "-o", str(out), "-prefix", "run"])
        print(tag, [l for l in log.splitlines() if "lambda_null" in l])
        outs[tag] = (out / "run.traitA.lmm.tsv").read_text().splitlines()
    for x, y in (("a1", "b1"), ("a1", "c2")):
        a, b = outs[x], outs[y]
        diff = [i for i, (p, q) in enumerate(zip(a, b)) if p != q]
        print(x, y, "lines", len(a), len(b), "differing", len(diff))
        for i in diff[:4]:
            print("  ", i, a[i])
            print("  ", i, b[i])


if __name__ == "__main__":
    main()
