"""world_size=2 gloo test (CPU) of the multi-GPU host logic: one broadcast of the null model, contiguous
SNP-range shards scanned independently, ordered gather == single-process result.  The per-shard compute
is the CPU oracle here (no GPU in this container); on the GPU box the same functions drive the device scan
(tests/test_dist_gpu.py)."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    sys.path.insert(0, {root!r} + "/tests")
    from janusx_b200 import dist as jd, synth
    from oracle import oracle as O
    from conftest import make_problem, null_model
    d = jd.init_process_group("gloo")
    rank, world, _ = jd.env_rank_world()
    prefix, out = sys.argv[1], sys.argv[2]
    n, p, m = 60, 2, 45
    model = None
    if rank == 0:
        case = make_problem(n=n, m=m, q=1, seed=21, missing_rate=0.04)
        nm = null_model(O, case)
        model = jd.NullModel(case.s, nm["xcov"], nm["y"], nm["ut"], nm["low"], nm["high"], nm["lbd"])
    model = jd.broadcast_null_model(model, n, p)
    assert model.u_t.shape == (n, n) and model.u_t.dtype == np.float32

    def scan_range(b, e, part, header):
        # oracle scan of BED rows [b, e): write a shard file
        fam = O.read_fam(prefix); packed = O.read_bed(prefix, len(fam)); sites = O.read_bim(prefix)
        keep, af, mr, missing = O.count_qc_block(packed[b:e], n, None, 0.02, 0.05, 1.0)
        idx = np.nonzero(keep)[0]
        rows = 0
        with open(part, "wb") as fh:
            if header: fh.write(O.HEADERS[3])
            if idx.size:
                g = O.decode_centered_block(packed[b:e], n, af[idx], row_indices=idx)
                res = O.lmm_reml_chunk_from_snp_f32(model.s, model.xcov, model.y, model.low, model.high, g, model.u_t, 30, 1e-2)
                for k, j in enumerate(idx):
                    c, snp, pos, a0, a1 = sites[b + j]
                    fh.write(O.format_row(c, pos, snp, a0, a1, float(af[j]), float(np.float32(missing[j]) / np.float32(n)), res[k]))
                    rows += 1
        return rows

    total = jd.scan_bed_sharded(prefix, out, m, scan_range)
    # array-level ordered gather
    b, e = jd.shard_range(m, rank, world)
    local = np.arange(b, e, dtype=np.float64)[:, None] * np.ones((1, 3))
    allrows = jd.gather_rows_in_order(local)
    if rank == 0:
        assert np.array_equal(allrows[:, 0], np.arange(m))
        print("TOTAL", total)
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_sharded_scan_equals_single(tmp_path, oracle):
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import make_problem, null_model
    from janusx_b200 import synth
    case = make_problem(n=60, m=45, q=1, seed=21, missing_rate=0.04)
    nm = null_model(oracle, case)
    prefix = str(tmp_path / "panel")
    synth.write_plink(prefix, case.packed, case.n)
    single = tmp_path / "single.tsv"
    rows1 = oracle.scan_bed_to_tsv(prefix, str(single), case.s, nm["xcov"], nm["y"], nm["ut"], 0.02, 0.05, 1.0,
                                   low=nm["low"], high=nm["high"])
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER.format(root=str(ROOT)))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(worker), prefix, str(tmp_path / "sharded.tsv")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert f"TOTAL {rows1}" in outs[0]
    assert (tmp_path / "sharded.tsv").read_bytes() == single.read_bytes()
    assert not list(tmp_path.glob("sharded.tsv.part*"))


def test_shard_ranges_cover_in_order():
    from janusx_b200.dist import shard_range
    for m in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(m, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1
