"""The north-star job end to end through the shipped CLI: synthetic PLINK files (n samples, m SNPs) -> 
`python -m janusx_b200.gwas -lmm2 -k 1 -q 3 -gpus G` for each G given, TSVs compared byte for byte, sampled rows of the
largest-G output checked against the CPU oracle run on the dumped null model.
usage: python tools/job_1m.py [n] [m] [gpus,comma,separated] [out.json]"""
import hashlib
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench as B  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    gpus = [int(g) for g in (sys.argv[3] if len(sys.argv) > 3 else "8,1").split(",")]
    out_json = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/r2_job.json"
    tmp = Path(os.environ.get("TMPDIR", "/tmp")) / "jxb_job"
    tmp.mkdir(exist_ok=True)
    prefix = str(tmp / "panel")
    dev = torch.device("cuda:0")
    t0 = time.time()
    gv = torch.zeros(n, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(B.SEED + 7)
    with open(prefix + ".bed", "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        for b0 in range(0, m, 65536):
            rows = min(65536, m - b0)
            pk = B.gen_snp_range(torch, n, b0, b0 + rows, dev)
            fh.write(pk.cpu().numpy().tobytes())
            if b0 < 65536:        # a polygenic trait from the first 4,096 SNPs
                sub = pk[:4096]
                codes = torch.stack([(sub >> (2 * k)) & 3 for k in range(4)], dim=2).reshape(sub.shape[0], -1)[:, :n]
                dos = torch.tensor([0, 0, 1, 2], device=dev, dtype=torch.float64)[codes.long()]
                dos -= dos.mean(dim=1, keepdim=True)
                gv += torch.randn(sub.shape[0], generator=gen, device=dev, dtype=torch.float64) @ dos
    y = 100.0 + gv + torch.randn(n, generator=gen, device=dev, dtype=torch.float64) * float(gv.std())
    with open(prefix + ".bim", "w") as fh:
        fh.write("".join(f"1\tsnp{i}\t0\t{i}\tA\tT\n" for i in range(m)))
    with open(prefix + ".fam", "w") as fh:
        fh.write("".join(f"F{j}\tS{j}\t0\t0\t0\t-9\n" for j in range(n)))
    yh = y.cpu().numpy()
    with open(tmp / "pheno.tsv", "w") as fh:
        fh.write("id\ttrait\n" + "".join(f"S{j}\t{yh[j]:.10f}\n" for j in range(n)))
    del gv, y
    torch.cuda.empty_cache()
    res = {"n": n, "m": m, "data_s": round(time.time() - t0, 1), "runs": {}}
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    digests = {}
    for g in gpus:
        out = tmp / f"out{g}"
        e = dict(env)
        if g == gpus[0]:
            e["JXB_DEBUG_DUMP_NULL"] = str(tmp / "null.npz")
        t1 = time.time()
        r = subprocess.run([sys.executable, "-m", "janusx_b200.gwas", "-bfile", prefix, "-p", str(tmp / "pheno.tsv"), "-lmm2",
                            "-k", "1", "-q", "3", "-force-model", "-gpus", str(g), "-o", str(out), "-prefix", "job"], env=e,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.time() - t1
        log = [l for l in r.stdout.splitlines() if l.startswith("[") or l.startswith("done")]
        tsv = out / "job.trait.lmm2.tsv"
        if r.returncode or not tsv.exists():
            res["runs"][str(g)] = {"rc": r.returncode, "tail": r.stdout[-1500:]}
            continue
        h = hashlib.md5()
        with open(tsv, "rb") as fh:
            for blk in iter(lambda: fh.read(1 << 24), b""):
                h.update(blk)
        digests[g] = h.hexdigest()
        res["runs"][str(g)] = {"rc": 0, "wall_s": round(dt, 2), "md5": digests[g], "tsv_mb": round(tsv.stat().st_size / 1e6, 1),
                               "log": log}
    res["byte_identical"] = len(set(digests.values())) == 1 and len(digests) == len(gpus)
    # sampled oracle parity on the first run's output
    try:
        from oracle import oracle as O
        from test_parity_gpu import _assert_row_equiv
        O.build()
        z = np.load(tmp / "null.npz")
        tsv = tmp / f"out{gpus[0]}" / "job.trait.lmm2.tsv"
        lines = tsv.read_bytes().split(b"\n")
        rows = len(lines) - 2
        bps = (n + 3) // 4
        raw = np.memmap(prefix + ".bed", dtype=np.uint8, mode="r")[3:].reshape(m, bps)
        picks = sorted(set(list(range(0, 6)) + list(range(rows // 2, rows // 2 + 6)) + list(range(rows - 6, rows))))
        src = [int(lines[1 + k].split(b"\t")[1]) for k in picks]       # pos column == BED row index
        sub = np.ascontiguousarray(raw[src])
        keep, af, mr, missing = O.count_qc_block(sub, n, None, 0.02, 0.05, 1.0)
        assert keep.all()
        g = O.decode_centered_block(sub, n, af)
        l10 = float(np.log10(float(z["lbd"])))
        want = O.lmm_reml_lmm2_chunk_f32(z["s"], z["xcov"], z["y"], float(z["low"]), float(z["high"]),
                                         O.rotate_block(g, z["u_t"]), float(z["nullml"]), 30, 1e-2, init_reml=l10)
        for k, j, w in zip(picks, src, want):
            rate = float(np.float32(missing[picks.index(k)]) / np.float32(n))
            line = O.format_row("1", j, f"snp{j}", "A", "T", float(af[picks.index(k)]), rate, w).rstrip(b"\n").split(b"\t")
            _assert_row_equiv(lines[1 + k].split(b"\t"), line)
        res["oracle_rows_checked"] = len(picks)
        res["rows_written"] = rows
    except Exception as ex:  # noqa: BLE001
        res["oracle_error"] = repr(ex)[:500]
    Path(out_json).parent.mkdir(exist_ok=True)
    Path(out_json).write_text(json.dumps(res, indent=1))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
