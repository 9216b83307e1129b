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
