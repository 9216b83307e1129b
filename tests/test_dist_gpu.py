"""Two ranks (gloo rendezvous, both on cuda:0 of the 1-GPU test box) scan contiguous SNP shards of one BED with
the DEVICE path and the ordered concatenation must equal the single-process file byte for byte: exercises
snp_begin/snp_end, write_header, the one-broadcast null model and the rank-order gather on real kernels."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    from janusx_b200 import dist as jd, jxrs
    jd.init_process_group("gloo")
    rank, world, _ = jd.env_rank_world()
    prefix, out, npz = sys.argv[1], sys.argv[2], sys.argv[3]
    Z = np.load(npz)
    n, p, m = int(Z["n"]), int(Z["p"]), int(Z["m"])
    model = None
    if rank == 0:
        model = jd.NullModel(Z["s"], Z["xcov"], Z["y"], Z["ut"], float(Z["low"]), float(Z["high"]), float(Z["lbd"]))
    model = jd.broadcast_null_model(model, n, p)
    dev = jxrs.DeviceModel(model.s, model.xcov, model.y, model.u_t, device=0)
    def scan_range(b, e, part, header):
        return dev.scan_bed_to_tsv(prefix, part, 0.02, 0.05, 1.0, mode="lmm", low=model.low, high=model.high,
                                   batch_rows=97, snp_begin=b, snp_end=e, write_header=header)
    total = jd.scan_bed_sharded(prefix, out, m, scan_range)
    if rank == 0:
        print("TOTAL", total)
""")


def test_two_rank_device_scan_equals_single(tmp_path, oracle):
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import make_problem, null_model
    from janusx_b200 import jxrs, synth
    case = make_problem(n=180, m=401, q=2, seed=33, missing_rate=0.03)
    nm = null_model(oracle, case)
    prefix = str(tmp_path / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    single = tmp_path / "single.tsv"
    dev = jxrs.DeviceModel(case.s, nm["xcov"], nm["y"], nm["ut"])
    rows1 = dev.scan_bed_to_tsv(prefix, str(single), 0.02, 0.05, 1.0, mode="lmm", low=nm["low"], high=nm["high"])
    np.savez(tmp_path / "null.npz", n=case.n, p=nm["xcov"].shape[1], m=401, s=case.s, xcov=nm["xcov"], y=nm["y"],
             ut=nm["ut"], low=nm["low"], high=nm["high"], lbd=nm["lbd"])
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER.format(root=str(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(worker), prefix, str(tmp_path / "sharded.tsv"),
                                       str(tmp_path / "null.npz")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert f"TOTAL {rows1}" in outs[0]
    assert (tmp_path / "sharded.tsv").read_bytes() == single.read_bytes()


def _run_cli(args, timeout=900):
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "janusx_b200.gwas", *args], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout
    return r.stdout


def test_cli_one_job_on_n_ranks_is_byte_identical(tmp_path, oracle):
    """The shipped multi-GPU job (`python -m janusx_b200.gwas ... -gpus N`, src/stats/lmm.rs:2488-2750 scans the whole BED
    in one call): rank 0 GRM + eigh + null fit -> one broadcast -> contiguous SNP shards -> ordered concatenation.
    The TSV must be byte-identical for N = 1, 2, 3 (ranks take distinct GPUs when the box has them and rendezvous over
    NCCL; on a 1-GPU box they share cuda:0 over gloo), and sampled rows match the oracle on the N-rank output."""
    import torch
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import make_problem
    from janusx_b200 import synth
    from test_parity_gpu import _assert_row_equiv, _tsv_fields
    case = make_problem(n=400, m=3000, q=0, seed=123, missing_rate=0.02)
    prefix = str(tmp_path / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    with open(tmp_path / "pheno.tsv", "w") as fh:
        fh.write("id\ttraitA\n")
        for j in range(case.n):
            fh.write(f"S{j}\t{case.y[j]:.10f}\n")
    outs = {}
    ndev = torch.cuda.device_count()
    counts = [1, 2, 3] if ndev < 4 else sorted({1, 2, 4, min(8, ndev)})
    for g in counts:
        out = tmp_path / f"out{g}"
        log = _run_cli(["-bfile", prefix, "-p", str(tmp_path / "pheno.tsv"), "-lmm", "-lmm2", "-fvlmm", "-k", "1", "-q", "2",
                        "-force-model", "-gpus", str(g), "-o", str(out), "-prefix", "run"])
        outs[g] = {m: (out / f"run.traitA.{m}.tsv").read_bytes() for m in ("lmm", "lmm2", "fvlmm")}
        assert not list(out.glob("*.part*")) and not list(out.glob("*.tmp")), log
    for g in counts[1:]:
        for m in ("lmm", "lmm2", "fvlmm"):
            assert outs[g][m].count(b"\n") == outs[1][m].count(b"\n"), (g, m, "row count")
            assert outs[g][m] == outs[1][m], (g, m)
    # every kept SNP exactly once, in BED order
    keep, af, mr, missing = oracle.count_qc_block(case.packed, case.n, None, 0.02, 0.05, 1.0)
    lines = outs[counts[-1]]["lmm"].split(b"\n")
    assert len(lines) == int(keep.sum()) + 2
    assert [l.split(b"\t")[2] for l in lines[1:-1]] == [f"snp{i}".encode() for i in np.nonzero(keep)[0]]
