"""2-GPU diagnosis: the same small job on GPU 0, on GPU 1 (-gpu 1) and as 2 NCCL ranks; reports differing TSV lines."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

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
        print(r.stdout[-3000:])
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
    for tag, extra in (("g0", ["-gpu", "0"]), ("g1", ["-gpu", "1"]), ("n2", ["-gpus", "2"])):
        out = tmp / tag
        log = run(["-bfile", prefix, "-p", str(tmp / "pheno.tsv"), "-lmm", "-lmm2", "-fvlmm", "-k", "1", "-q", "2", "-force-model",
                   *extra, "-o", str(out), "-prefix", "run"])
        print(tag, [l for l in log.splitlines() if "lambda_null" in l or "SNPs ->" in l])
        outs[tag] = {m: (out / f"run.traitA.{m}.tsv").read_text().splitlines() for m in ("lmm", "lmm2", "fvlmm")
                     if (out / f"run.traitA.{m}.tsv").exists()}
    for x, y in (("g0", "g1"), ("g0", "n2")):
        for m in ("lmm", "lmm2", "fvlmm"):
            a, b = outs[x].get(m, []), outs[y].get(m, [])
            diff = [i for i, (p, q) in enumerate(zip(a, b)) if p != q]
            print(x, y, m, "lines", len(a), len(b), "differing", len(diff), "first", diff[:3], "last", diff[-3:])
            for i in diff[:2] + diff[-1:]:
                print("   ", i, a[i])
                print("   ", i, b[i])


if __name__ == "__main__":
    main()
