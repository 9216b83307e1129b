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


def run(args, dump=None):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    if dump:
        env["JXB_DEBUG_DUMP_NULL"] = dump
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
                   "-o", str(out), "-prefix", "run"], dump=str(tmp / f"null_{tag}.npz"))
        print(tag, [l for l in log.splitlines() if "lambda_null" in l])
        outs[tag] = (out / "run.traitA.lmm.tsv").read_text().splitlines()
    for x, y in (("a1", "b1"), ("a1", "c2")):
        za, zb = np.load(tmp / f"null_{x}.npz"), np.load(tmp / f"null_{y}.npz")
        print(x, y, "null model:", {k: (bool(np.array_equal(za[k], zb[k])), float(np.max(np.abs(za[k] - zb[k])))) for k in za.files})
        a, b = outs[x], outs[y]
        diff = [i for i, (p, q) in enumerate(zip(a, b)) if p != q]
        print(x, y, "lines", len(a), len(b), "differing", len(diff))
        for i in diff[:4]:
            print("  ", i, a[i])
            print("  ", i, b[i])


if __name__ == "__main__":
    main()
