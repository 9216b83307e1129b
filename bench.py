#!/usr/bin/env python
"""bench.py -- SNPs/s of the exact LMM scan (decode -> rotate -> per-SNP REML/ML solve) on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1) prints ONE JSON line on rank 0.

Workload = BASELINE.json configs[2]: n=20,000 samples, -lmm2 (Wald + LRT), 1 trait, 3 covariates, ONE job of
m=1,000,000 SNPs.  `--scaling strong` (default): a step = one pass over the WHOLE job, SNP-sharded over the N ranks in
contiguous ranges (src/stats/lmm.rs:1213-1215 byte offsets), every rank scanning its range in device batches and the
result rows gathered in order on rank 0; the job's data are synthesised from counter-based chunks keyed by the global
SNP index, so every N sees the same SNPs.  Timed region of a step at N>1: NCCL broadcast of U^T + scan + ordered gather
(the eigendecomposition is timed separately: `null_model.eigh_s`).  `--scaling weak` keeps round 1's measurement (every
rank scans its own fixed batches).  `value` is whole-job SNPs/s with the packed genotypes resident in HBM; `e2e` is the
same job through the reference-facing C-ABI call jxb_scan_packed with pinned HOST buffers (H2D of every packed batch
and D2H of every result row inside the timed region).  `--impl reference` times the CPU restatement of the reference
algorithm (oracle port: numpy/OpenBLAS f32 rotation like the reference's cblas_sgemm + OpenMP per-SNP solve) on all
host cores; its setup uses torch only (no janusx_b200 import).
"""
from __future__ import annotations

import os
import sys

# Host threads for the CPU legs: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the reference arm /
# cpu_baseline on one core.  Rank 0 runs those legs, so it gets every core it is allowed to use -- set before numpy /
# OpenBLAS / the OpenMP oracle are loaded.
if int(os.environ.get("RANK", "0")) == 0:
    _cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_cores)

import argparse
import json
import subprocess
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260609  # the reference's benchmark seed (scripts/benchmark.sh:36)
CHUNK = 8192     # synthetic SNPs are generated in chunks of this many rows, keyed by the global chunk index


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default=os.environ.get("JXB_BENCH_SCALING", "strong"), choices=["strong", "weak"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("JXB_BENCH_N", 20000)))
    ap.add_argument("--job-snps", type=int, default=int(os.environ.get("JXB_BENCH_JOB", 0)),
                    help="strong scaling: SNPs of the whole job (default 1,000,000 for n <= 24,000, else 303,104)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("JXB_BENCH_BATCH", 0)),
                    help="SNPs per device batch; default = the library's: 151,552 for n <= 24,000, else 75,776")
    ap.add_argument("--model", default=os.environ.get("JXB_BENCH_MODEL", "lmm2"), choices=["lmm", "lmm2", "fvlmm"])
    ap.add_argument("--grm-snps", type=int, default=int(os.environ.get("JXB_BENCH_GRM_SNPS", 50000)))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("JXB_BENCH_CPU_SAMPLE", 1024)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=int(os.environ.get("JXB_BENCH_E2E_STEPS", 3)),
                    help="steps of the host-buffer leg (it repeats the job; capped so the default run stays in minutes)")
    ap.add_argument("--rotate-variant", type=int, default=int(os.environ.get("JXB_BENCH_ROTATE", 3)),
                    help="3 = hand-written tcgen05 int8-sliced exact rotation (default), 2 = same via cuBLASLt, "
                         "0 = FP64 DMMA GEMM")
    ap.add_argument("--overlap", type=int, default=int(os.environ.get("JXB_BENCH_OVERLAP", 0)),
                    help="1 = streamed scan: rotation slabs under one persistent solve kernel (measured slower: both kernels "
                         "are bound by the shared-memory pipe); 0 = rotate then solve (default)")
    ap.add_argument("--prefix", type=int, default=1,
                    help="1 (default) = the three SNP-independent leading evaluations of every REML search come from per-batch "
                         "tables; 0 = every evaluation in the lane kernel (same results)")
    ap.add_argument("--slab", type=int, default=int(os.environ.get("JXB_BENCH_SLAB", 0)), help="rows per rotation slab")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic inputs on the GPU (distributions of `jx sim`, python/janusx/script/sim.py:49-66, 133-176, 252-276)
# ------------------------------------------------------------------------------------------------------
def gen_packed_batch(torch, n, rows, key, device, want_dosage=False):
    """HWE genotypes for `rows` SNPs keyed by (SEED, key) so every GPU count sees the same data."""
    g = torch.Generator(device=device)
    g.manual_seed(SEED * 1000003 + int(key))
    maf = torch.rand(rows, generator=g, device=device, dtype=torch.float32) * (0.45 - 0.02) + 0.02
    p0 = (1.0 - maf) ** 2
    p1 = p0 + 2.0 * maf * (1.0 - maf)
    npad = (n + 3) // 4 * 4
    packed = torch.zeros((rows, npad // 4), dtype=torch.uint8, device=device)
    dos_all = [] if want_dosage else None
    code_of = torch.tensor([0, 2, 3], dtype=torch.uint8, device=device)  # dosage -> PLINK code 00/10/11
    step = max(1, (1 << 28) // max(npad, 1))
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        u = torch.rand((r1 - r0, npad), generator=g, device=device, dtype=torch.float32)
        dos = (u >= p0[r0:r1, None]).to(torch.uint8) + (u >= p1[r0:r1, None]).to(torch.uint8)
        if npad != n:
            dos[:, n:] = 0
        if want_dosage:
            dos_all.append(dos[:, :n].clone())
        c = code_of[dos.long()].view(r1 - r0, npad // 4, 4)
        packed[r0:r1] = c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)
    return packed, (torch.cat(dos_all) if want_dosage else None)


def gen_snp_range(torch, n, begin, end, device):
    """Packed rows of the global SNP range [begin, end): chunks of CHUNK rows keyed by the global chunk index."""
    out = torch.empty((end - begin, (n + 3) // 4), dtype=torch.uint8, device=device)
    c = begin // CHUNK
    while c * CHUNK < end:
        pk, _ = gen_packed_batch(torch, n, CHUNK, c, device)
        lo, hi = max(begin, c * CHUNK), min(end, (c + 1) * CHUNK)
        out[lo - begin:hi - begin] = pk[lo - c * CHUNK:hi - c * CHUNK]
        c += 1
    return out


def _phenotype_and_design(torch, n, q, gv, device, gt):
    """gt: the generator that already drew the SNP effects (same draw order as round 1's bench, so the null model and
    with it the evaluation counts per SNP stay comparable between rounds)."""
    vg = float(gv.var(unbiased=False))
    y = 100.0 + gv + torch.randn(n, generator=gt, device=device, dtype=torch.float64) * (vg ** 0.5)   # pve 0.5
    gc = torch.Generator(device=device)
    gc.manual_seed(SEED + 2)
    cov = torch.randn((n, q), generator=gc, device=device, dtype=torch.float64)
    X = torch.cat([torch.ones((n, 1), dtype=torch.float64, device=device), cov], dim=1)
    return X, y


def build_null_model(torch, n, grm_snps, q, device, timings=None, use_library=True):
    """Null-model inputs.  use_library: centred VanRaden GRM of `grm_snps` synthetic SNPs on the int8 tensor cores
    (csrc/grm.cu; src/stats/grm.rs:204-608) and K + 1e-6 I decomposed by the library's eigensolver (csrc/eigh.cu;
    workflow_model_stream.py:902).  Otherwise (reference arm) the same GRM by a plain torch f64 matmul and
    torch.linalg.eigh -- no janusx_b200 code.  Phenotype 100 + G beta + e at pve 0.5, q N(0,1) covariates.
    Returns host arrays (s, X design, y) and U^T as an f32 device tensor (pyBLUP/assoc.py:1818)."""
    timings = {} if timings is None else timings
    if n > 46340:
        # one cusolverDnXsyevd call rejects n*n >= 2^31: model two unrelated populations (block-diagonal GRM and U^T)
        h = n // 2
        s1, u1, X1, y1 = build_null_model(torch, h, grm_snps, q, device, timings, use_library)
        s2, u2, X2, y2 = build_null_model(torch, n - h, max(1024, grm_snps // 2), q, device, None, use_library)
        u_t = torch.zeros((n, n), dtype=torch.float32, device=device)
        u_t[:h, :h] = u1
        u_t[h:, h:] = u2
        timings["blocks"] = 2
        return np.concatenate([s1, s2]), u_t, np.concatenate([X1, X2]), np.concatenate([y1, y2])
    gv = torch.zeros(n, dtype=torch.float64, device=device)
    gb = torch.Generator(device=device)
    gb.manual_seed(SEED + 1)
    chunk = 16384
    t_grm = 0.0
    grm = None
    K = None
    varsum = 0.0
    if use_library:
        from janusx_b200 import jxrs
        grm = jxrs.DeviceGrm(n, None, 1, device.index or 0)
    else:
        K = torch.zeros((n, n), dtype=torch.float64, device=device)
    for b, r0 in enumerate(range(0, grm_snps, chunk)):
        rows = min(chunk, grm_snps - r0)
        packed, dos = gen_packed_batch(torch, n, rows, 10_000_000 + b, device, want_dosage=True)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        if use_library:
            grm.update_dev(packed.data_ptr(), rows, packed.shape[1])
        beta = torch.randn(rows, generator=gb, device=device, dtype=torch.float64)
        for c0 in range(0, rows, 4096):
            z = dos[c0:c0 + 4096].to(torch.float64)
            mu = z.mean(dim=1, keepdim=True)
            z -= mu
            if not use_library:
                K += z.T @ z
                p = (mu[:, 0] / 2.0)
                varsum += float((2.0 * p * (1.0 - p)).sum())
            gv += beta[c0:c0 + 4096] @ z
        if use_library:
            t_grm += time.perf_counter() - t0
        del z, dos, packed
    s = torch.empty(n, dtype=torch.float64, device=device)
    u_t = torch.empty((n, n), dtype=torch.float32, device=device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    if use_library:
        grm.eigh_dev(s.data_ptr(), u_t.data_ptr(), 1e-6)
        grm.close()
    else:
        K /= varsum
        K.diagonal().add_(1e-6)
        w, v = torch.linalg.eigh(K)
        s.copy_(w)
        u_t.copy_(v.T)
        del K, w, v
    torch.cuda.synchronize(device)
    timings["eigh_s"] = time.perf_counter() - t0
    timings["grm_s"] = t_grm
    timings["grm_snps"] = grm_snps
    X, y = _phenotype_and_design(torch, n, q, gv, device, gb)
    return s.cpu().numpy(), u_t, X.cpu().numpy(), y.cpu().numpy()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measure_dgemm_peak(torch, device, n):
    """cuBLAS DGEMM for the rotation shape (M=4096, N=K=n): burst (best of 5) TFLOP/s -- the FP64 tensor ceiling."""
    m = 4096
    a = torch.randn((m, n), dtype=torch.float64, device=device)
    b = torch.randn((n, n), dtype=torch.float64, device=device)
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = a @ b.T
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, 2.0 * m * n * n / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b, c
    return best


def measure_int8_peak(torch, device, n):
    """cuBLASLt int8 GEMM (torch._int_mm) for the rotation shape: burst TOP/s -- the int8 tensor ceiling the sliced
    rotation's MMAs run against."""
    m = 8192
    k = (n + 7) // 8 * 8
    try:
        # operand statistics of the rotation itself (tensor-core power, and with it the clock under the 1 kW cap, depends
        # on how many operand bits toggle): A = dosages 0/1/2, B = balanced base-256 digits; column-major B ("TN")
        a = torch.randint(0, 3, (m, k), dtype=torch.int8, device=device)
        b = torch.randint(-128, 128, (k, k), dtype=torch.int8, device=device).T
        best = 0.0
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c = torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            if i:
                best = max(best, 2.0 * m * k * k / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b, c
        return best
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ------------------------------------------------------------------------------------------------------
def cpu_scan(O, packed_rows, n, s, xcov, y, ut_f32, low, high, model, nullml, l10, stages=None):
    """One pass of the reference algorithm on the host: count/QC, decode, f32 GEMM rotation (numpy/OpenBLAS,
    all cores -- the reference's cblas_sgemm stage), per-SNP solve (OpenMP, one SNP per task).  `stages` (a dict)
    accumulates seconds per stage."""
    t = [time.perf_counter()]

    def lap(name):
        t.append(time.perf_counter())
        if stages is not None:
            stages[name] = stages.get(name, 0.0) + (t[-1] - t[-2])

    keep, af, mr, missing = O.count_qc_block(packed_rows, n, None, 0.02, 0.05, 1.0)
    idx = np.nonzero(keep)[0]
    lap("count_qc")
    g = O.decode_centered_block(packed_rows, n, af[idx], row_indices=idx)
    lap("decode")
    rot = g @ ut_f32.T
    lap("rotate")
    if model == "lmm2":
        out = O.lmm_reml_lmm2_chunk_f32(s, xcov, y, low, high, rot, nullml, 30, 1e-2, 0, init_reml=l10)
    elif model == "fvlmm":
        out, _ = O.lmm_assoc_chunk_f32(s, xcov, y, l10, rot, 0, None)
    else:
        out = O.lmm_reml_chunk_f32(s, xcov, y, low, high, rot, 30, 1e-2, 0, None)
    lap("solve")
    return out


def workload_config(args, n, B, q, job, world, impl):
    what = {"lmm": "-lmm (Wald)", "lmm2": "-lmm2 (Wald + LRT)", "fvlmm": "-fvlmm (fixed lambda)"}[args.model]
    if args.scaling == "strong":
        wl = (f"synthetic n={n}, ONE job of m={job:,} SNPs per step, SNP-sharded in contiguous ranges, device batches of "
              f"{B} SNPs, {what}, 1 trait, {q} covariates (BASELINE.json configs[2])")
    else:
        wl = (f"synthetic n={n}, m=1,000,000 job sampled in per-rank batches of {B} SNPs, {what}, 1 trait, {q} covariates "
              "(BASELINE.json configs[2])")
    return {"workload": wl, "n": n, "batch_snps": B, "job_snps": job if args.scaling == "strong" else None,
            "covariates": q, "model": args.model,
            "parallelism": f"snp-shard x{world}" if impl == "b200" else "host threads",
            "rotation": {0: "fp64-dmma", 1: "fp64-cuda-core", 2: "int8-sliced-exact (cuBLASLt)",
                         3: "int8-sliced-exact (tcgen05)"}[args.rotate_variant],
            "overlap": "rotation slabs under one persistent solve kernel" if args.overlap else "rotate then solve",
            "l2": "inputs larger than L2: every batch streams the digit planes of U^T (2.8 GB at n=20k) and its own "
                  "rotated block (12 GB)"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference" and rank != 0:
        return 0

    import torch

    n, B, q = args.n, args.batch, 3
    if B <= 0:
        B = 2 * 75776 if n <= 24000 else 75776          # janusx_b200.jxrs.default_device_batch
        args.batch = B
    job = args.job_snps if args.job_snps > 0 else (1_000_000 if n <= 24000 else 4 * 75776)
    p = q + 1
    config = workload_config(args, n, B, q, job, world, args.impl)

    have_gpu = torch.cuda.is_available()
    if args.impl == "b200" and not have_gpu:
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    device = torch.device(f"cuda:{local_rank}") if have_gpu else torch.device("cpu")
    if have_gpu:
        torch.cuda.set_device(device)
    dist = None
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    # ---- null model: GRM + eigh once on rank 0, NCCL broadcast of U^T (f32), S, X, y ------------------------------
    t_setup = time.time()
    setup_timings = {}
    if args.impl == "reference" and not have_gpu:
        # CPU-only host: small orthogonal basis from numpy (the reference arm still times the same algorithm)
        grm_m = min(args.grm_snps, 4 * n)
        from janusx_b200 import synth
        pk, _ = synth.draw_genotypes(grm_m, n, seed=SEED)
        K = synth.vanraden_grm(pk, n)
        K[np.diag_indices(n)] += 1e-6
        s_np, u = np.linalg.eigh(K)
        u_t_host = np.ascontiguousarray(u.T.astype(np.float32))
        X_np = np.concatenate([np.ones((n, 1)), np.random.default_rng(SEED + 2).normal(size=(n, q))], axis=1)
        y_np = 100.0 + np.random.default_rng(SEED + 1).normal(size=n) * 1.4
        u_t_dev = None
    else:
        if rank == 0:
            s_np, u_t_dev, X_np, y_np = build_null_model(torch, n, args.grm_snps, q, device, setup_timings,
                                                         use_library=(args.impl == "b200"))
        else:
            s_np = np.zeros(n); X_np = np.zeros((n, p)); y_np = np.zeros(n)
            u_t_dev = torch.empty((n, n), dtype=torch.float32, device=device)
        if dist is not None:
            small = torch.as_tensor(np.concatenate([s_np, X_np.reshape(-1), y_np]), device=device)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dist.broadcast(small, 0)
            dist.broadcast(u_t_dev, 0)      # 4*n*n bytes over NVLink, once per job
            torch.cuda.synchronize()
            setup_timings["broadcast_first_s"] = time.perf_counter() - t0   # includes NCCL communicator set-up
            sm = small.cpu().numpy()
            s_np, X_np, y_np = sm[:n].copy(), sm[n:n + n * p].reshape(n, p).copy(), sm[n + n * p:].copy()
        u_t_host = None

    if args.impl == "reference":
        return run_reference(args, torch, config, n, B, q, s_np, X_np, y_np, u_t_dev, u_t_host, device, have_gpu)

    from janusx_b200 import _cabi, jxrs
    lib = _cabi.lib()
    lib.jxb_set_rotate_variant(args.rotate_variant)
    lib.jxb_set_stream_overlap(args.overlap, args.slab)
    lib.jxb_set_prefix_evals(1 if args.prefix else 0)
    # rotate X, y and fit the null on the device (pyBLUP/assoc.py:1818-1876)
    t0 = time.perf_counter()
    mdl = jxrs.DeviceModel(s_np, np.ones((n, p)), np.zeros(n), u_t_dev, device=local_rank, u_t_on_device=True)
    xcov, yrot = mdl.rotate_xy(X_np, y_np)
    mdl.set_xy(xcov, yrot[:, 0])
    lbd, ml0, reml0 = mdl.reml_null(-5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd))
    low, high = l10 - 2.0, l10 + 2.0
    nullml = None
    if args.model == "lmm2":
        _, nullml = mdl.ml_null(low, high, 30, 1e-2, l10)       # src/stats/lmm.rs:2901-2924
    setup_timings["model_upload_and_null_fit_s"] = time.perf_counter() - t0
    dgemm_peak = measure_dgemm_peak(torch, device, n) if rank == 0 else 0.0
    int8_peak = measure_int8_peak(torch, device, n) if rank == 0 else None
    fp64_rates = None
    if rank == 0:
        import ctypes as C
        r3 = (C.c_double * 3)()
        _cabi.check(lib.jxb_fp64_probe(local_rank, r3))
        fp64_rates = [float(v) for v in r3]
    ut_host = u_t_dev.cpu().numpy() if (rank == 0 and not args.no_cpu_baseline) else None
    bcast_buf = u_t_dev if (dist is not None and args.scaling == "strong") else None   # re-broadcast inside every step
    if bcast_buf is None:
        del u_t_dev
    torch.cuda.empty_cache()

    # ---- this rank's SNPs: resident in HBM + a pinned host copy for the e2e leg ------------------------------------
    bps = (n + 3) // 4
    if args.scaling == "strong":
        begin, end = (rank * job) // world, ((rank + 1) * job) // world      # janusx_b200.dist.shard_range
    else:
        n_bufs = min(4, max(1, args.steps))
        begin, end = rank * 64 * CHUNK * 1000, rank * 64 * CHUNK * 1000 + n_bufs * B   # disjoint per-rank SNP ranges
    shard_dev = gen_snp_range(torch, n, begin, end, device)
    shard_rows = end - begin
    shard_host = torch.empty((shard_rows, bps), dtype=torch.uint8, pin_memory=True)
    shard_host.copy_(shard_dev)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    cols = 6 if args.model == "lmm2" else 3
    scan_kw = dict(maf_thr=0.02, miss_thr=0.05, het_thr=1.0, genetic_model="add", mode=args.model, low=low, high=high,
                   max_iter=30, tol=1e-2, init=(l10 if args.model == "lmm2" else None), nullml=nullml, log10_lbd=l10)
    # per-rank result buffers in HBM (rows compacted in SNP order) -- what the ordered gather ships to rank 0
    max_shard = (job + world - 1) // world + 1 if args.scaling == "strong" else shard_rows
    res_out = torch.zeros((max_shard, cols), dtype=torch.float64, device=device)
    res_af = torch.zeros(max_shard, dtype=torch.float32, device=device)
    res_counts = torch.zeros((max_shard, 4), dtype=torch.int32, device=device)
    gather_bufs = None
    if dist is not None and rank == 0 and args.scaling == "strong":
        gather_bufs = ([torch.empty_like(res_out) for _ in range(world)], [torch.empty_like(res_af) for _ in range(world)],
                       [torch.empty_like(res_counts) for _ in range(world)])
    if args.scaling == "strong":
        batches = [(r0, min(shard_rows, r0 + B)) for r0 in range(0, shard_rows, B)]
    stage_acc = {}

    def gather_to_rank0():
        if dist is None or args.scaling != "strong":
            return
        dist.gather(res_out, gather_bufs[0] if rank == 0 else None, dst=0)
        dist.gather(res_af, gather_bufs[1] if rank == 0 else None, dst=0)
        dist.gather(res_counts, gather_bufs[2] if rank == 0 else None, dst=0)

    def step_resident(i, record=False):
        """One step with the packed rows resident in HBM; results stay in HBM (gathered on rank 0's GPU)."""
        kept = 0
        if args.scaling == "strong":
            if bcast_buf is not None:
                dist.broadcast(bcast_buf, 0)
            todo = batches
        else:
            b0 = (i % n_bufs) * B
            todo = [(b0, b0 + B)]
        for (r0, r1) in todo:
            mdl.scan_packed_dev(int(shard_dev.data_ptr()) + r0 * bps, r1 - r0, bps, n, None, **scan_kw)
            nk = mdl.scan_fetch_dev(r1 - r0, cols, int(res_out.data_ptr()) + kept * cols * 8,
                                    int(res_af.data_ptr()) + r0 * 4, int(res_counts.data_ptr()) + r0 * 16)
            kept += nk
            if record:
                for k, v in mdl.stage_ms().items():
                    stage_acc[k] = stage_acc.get(k, 0.0) + v
                stage_acc["batches"] = stage_acc.get("batches", 0) + 1
        gather_to_rank0()
        return kept

    def step_host(i):
        """One step through jxb_scan_packed with pinned host buffers: H2D of every packed batch and D2H of every
        result row inside the call; N>1: the host rows go back up for the NCCL gather and down again on rank 0."""
        kept, h2d, d2h = 0, 0, 0
        if args.scaling == "strong":
            if bcast_buf is not None:
                dist.broadcast(bcast_buf, 0)
            todo = batches
        else:
            b0 = (i % n_bufs) * B
            todo = [(b0, b0 + B)]
        outs = []
        hnp = shard_host.numpy()
        for (r0, r1) in todo:
            keep, af, missing, out = mdl.scan_packed(hnp[r0:r1], n, **scan_kw)
            outs.append(out)
            kept += out.shape[0]
            h2d += (r1 - r0) * bps
            d2h += out.shape[0] * (cols * 8 + 4) + (r1 - r0) * (16 + 4) + 32
        if dist is not None and args.scaling == "strong":
            allo = np.concatenate(outs, axis=0) if outs else np.zeros((0, cols))
            res_out[:allo.shape[0]].copy_(torch.from_numpy(allo), non_blocking=False)
            gather_to_rank0()
            if rank == 0:
                _ = [b.cpu() for b in gather_bufs[0]]
                d2h += sum(int(b.numel()) * 8 for b in gather_bufs[0])
            h2d += allo.size * 8
        return kept, h2d, d2h

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(mdl.stream, device=device)
    lib.jxb_set_timing(1)
    for i in range(args.warmup):
        step_resident(i)
    mdl.sync()
    # timed region: inputs resident (CUDA events on the model's own stream bracket the K steps; barrier + sync on both sides)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kept_total = 0
    e0.record(stream)
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        kept_total += step_resident(i, record=(i == args.steps - 1))
    torch.cuda.synchronize()
    e1.record(stream)
    mdl.sync()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    launches = _cabi.launch_count() - launches0
    kept_step = kept_total // max(args.steps, 1)
    # evaluation counts of the last batch (per-SNP i32, fetched outside the timed region)
    last_rows = (batches[-1][1] - batches[-1][0]) if args.scaling == "strong" else B
    _, _, _, _, evals = mdl.scan_fetch(last_rows, cols)
    mean_evals = float(evals.mean()) if evals.size and args.model != "fvlmm" else 0.0
    cached = 0 if args.model == "fvlmm" else (2 if (args.model == "lmm2" or nullml is not None) else 1)
    # the three leading REML abscissae are the same for every SNP: evaluated from per-batch tables (SNP column only)
    shared = 3 if (args.prefix and args.model != "fvlmm" and evals.size >= 2048) else 0
    shared_evals = float(np.minimum(np.maximum(evals - cached, 0), shared).mean()) if evals.size and shared else 0.0
    exec_evals = float(np.maximum(evals - cached - shared, 0).mean()) if evals.size and args.model != "fvlmm" else 0.0
    kept_last = int(evals.size)

    # e2e: same job through the C-ABI call with pinned HOST buffers (H2D + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    step_host(0)
    barrier()
    t0 = time.perf_counter()
    kept_e2e, h2d_b, d2h_b = 0, 0, 0
    for i in range(e2e_steps):
        k_, h_, d_ = step_host(i)
        kept_e2e += k_; h2d_b += h_; d2h_b += d_
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    tvec = torch.tensor([dev_ms, e2e_s * 1e3, wall_ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = float(tvec[0]), float(tvec[1]), float(tvec[2])
    snps_per_step = job if args.scaling == "strong" else world * B
    value = snps_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = snps_per_step * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        st = {k: v for k, v in stage_acc.items()}
        nb = max(1, int(st.get("batches", 1)))
        streamed = st.get("streamed", 0.0) > 0
        kept_rank = kept_step if args.scaling == "weak" or world == 1 else int(res_counts[:shard_rows, 3].ne(0).sum())
        d = p + 1
        # K3: the solve kernel's own duration on its stream (streamed: it runs concurrently with the rotation slabs)
        solve_s = st.get("solve_kernel", st.get("solve", 0.0)) * 1e-3
        rot_s = st.get("rotate", 0.0) * 1e-3
        flop_per_eval = n * (3 * d * (d + 1) / 2 + 5 * d + 3 + 2)   # + 1 divide + 1 log per sample
        solve_flop = mean_evals * flop_per_eval * kept_rank
        solve_flop_exec = (exec_evals * flop_per_eval + (4.0 * n * (4 * d + 6) if shared_evals > 0 else 0.0)) * kept_rank
        # the solve TUs are compiled without FMA contraction (reference rounding): ceiling = separate DMUL/DADD issue
        # rate measured in this run (jxb_fp64_probe), charged as 1 flop per instruction; FMA-rate x2 given for context
        fp64_issue = min(fp64_rates[1], fp64_rates[2]) if fp64_rates else None
        traffic = None
        tpath = ROOT / "profiles" / "r2_ncu_traffic.json"
        if tpath.exists():
            try:
                traffic = json.loads(tpath.read_text())
            except Exception:
                traffic = None

        def tr(kernel):
            if not traffic or traffic.get("n") != n or traffic.get("model") != args.model:
                return None
            ent = traffic.get(kernel)
            return ent["dram_bytes_per_kept_snp"] * kept_rank if ent else None

        rot_ops_int8 = 10.0 * 2.0 * n * n * kept_rank                # 7 + 3 digit-plane products (17 with missing calls)
        rot_roof = {"kernel": ("rotate_dmma_kernel (FP64 DMMA GEMM)" if args.rotate_variant == 0 else
                               ("i8_rotate_kernel (tcgen05 int8-sliced exact rotation, pass 2 + pass D per slab)" if args.rotate_variant == 3 else
                                "int8-sliced exact rotation (10 cuBLASLt slice GEMMs + recombine_kernel)")),
                    "bound": "tensor",
                    "achieved": (rot_ops_int8 / rot_s / 1e12) if rot_s > 0 else 0.0, "peak": int8_peak,
                    "unit": "int8 TOP/s", "frac": (rot_ops_int8 / rot_s / 1e12 / int8_peak) if rot_s > 0 and int8_peak else None,
                    "traffic": tr("rotation"), "launch_ms": rot_s * 1e3,
                    "fp64_equivalent_tflops": (2.0 * n * n * kept_rank / rot_s / 1e12) if rot_s > 0 else None,
                    "dgemm_peak_tflops": dgemm_peak,
                    "peak_source": "cuBLASLt int8 GEMM (torch._int_mm, M=8192, N=K=n) measured in this run; "
                                   "fp64_equivalent = 2*n*n flop per SNP against the cuBLAS DGEMM measured in this run; "
                                   "MEASURED_PEAKS.json holds neither an int8 nor an FP64 figure"
                                   + ("; streamed: the slabs run concurrently with the solve kernel on the same SMs" if streamed else "")}
        solve_roof = {"kernel": (f"solve_lane_stream_kernel<{p}>" if streamed else
                                 (f"solve_lane_kernel<{p}>" if kept_last >= 32768 else f"solve_warp_kernel<{p}>"))
                                + " (per-SNP REML/ML Brent, FP64 CUDA cores)", "bound": "fp64-cuda-core",
                      "achieved": solve_flop / solve_s / 1e12 if solve_s > 0 else 0.0, "peak": fp64_issue,
                      "unit": "TFLOP/s", "frac": (solve_flop / solve_s / 1e12 / fp64_issue) if solve_s > 0 and fp64_issue else None,
                      "frac_executed": (solve_flop_exec / solve_s / 1e12 / fp64_issue) if solve_s > 0 and fp64_issue else None,
                      "frac_of_fma_rate": (solve_flop / solve_s / 1e12 / (2.0 * fp64_rates[0])) if solve_s > 0 and fp64_rates else None,
                      "traffic": tr("solve"), "launch_ms": solve_s * 1e3,
                      "algorithmic_flop_per_launch": solve_flop,
                      "algorithmic_bytes_per_launch": (exec_evals * 2.0 + 1.0) * kept_rank * n * 4.0,
                      "peak_source": "FP64 DADD/DMUL issue rate measured in this run (jxb_fp64_probe: "
                                     + (", ".join(f"{v:.2f}" for v in fp64_rates) if fp64_rates else "n/a")
                                     + " T lane-instr/s for DFMA, DADD, DMUL). The reference's rounding forbids FMA contraction "
                                       "(-fmad=false), so 1 flop = 1 instruction; the algorithmic count charges 1 flop per divide / "
                                       "log, which cost ~5 DFMA + MUFU and a table log",
                      "traffic_source": "profiles/r2_ncu_traffic.json (ncu dram__bytes_read+write per launch, same command, scaled "
                                        "per kept SNP)" if traffic else None}
        dominant = solve_roof if solve_s >= rot_s else rot_roof
        for k in ("count_qc", "decode", "rotate", "solve", "h2d", "d2h", "solve_kernel"):
            if k in st:
                st[k] = round(st[k], 3)
        line = {
            "metric": "SNPs/sec exact -lmm scan (n=20k)", "value": value, "unit": "SNPs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "SNPs/s", "h2d_bytes_per_step": h2d_b // e2e_steps,
                    "d2h_bytes_per_step": d2h_b // e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": dominant, "roofline_rotation": rot_roof, "roofline_solve": solve_roof,
            "stage_ms_last_step_rank0": st, "wall_ms_per_step": wall_ms / args.steps,
            "solve": {"mean_objective_evals_per_snp": mean_evals, "executed_evals_per_snp": exec_evals,
                      "shared_abscissa_evals_per_snp": shared_evals,
                      "shared_abscissa_note": "first three abscissae of every REML search are SNP-independent (x0, golden step, "
                                              "one of two golden steps): 1/(s+lambda), the covariate sums and sum ln v come from "
                                              "per-batch tables and prefix_eval_kernel evaluates the four candidates in two sweeps, "
                                              "4(4d+6)n flop per SNP; counted at that cost in frac_executed",
                      "note": "the reference's final_beta_se / ml_loglike passes (and LMM2's first ML evaluation) repeat an "
                              "abscissa already evaluated: counted by both sides, executed once here"},
            "decode": {"hbm_gb_per_s": ((shard_rows * bps + kept_rank * n * (8 if args.rotate_variant == 0 else 3))
                                        / (st["decode"] * 1e-3) / 1e9) if st.get("decode", 0) > 0 else None},
            "null_model": dict({"lambda": lbd, "ml0": ml0, "reml0": reml0, "setup_s": setup_s,
                                "front": "GRM: csrc/grm.cu (tcgen05 int8), eigh: csrc/eigh.cu"},
                               **setup_timings),
            "kept_snps_per_step": kept_step if (args.scaling == "weak" or world == 1) else None,
            "kept_snps_rank0": kept_rank, "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            hnp = shard_host.numpy()
            line["cpu_baseline"] = cpu_baseline(args, n, B, s_np, xcov, yrot[:, 0].copy(), ut_host, hnp, low, high, nullml, l10)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(args, n, B, s_np, xcov, y, ut, packed_host, low, high, nullml, l10):
    """Oracle port timed on the box's host cores on a bounded SNP sample of the same workload."""
    from oracle import oracle as O
    O.build()
    rows = min(args.cpu_sample, packed_host.shape[0])
    stages = {}
    t0 = time.perf_counter()
    cpu_scan(O, packed_host[:rows], n, s_np, xcov, y, ut, low, high, args.model, nullml, l10, stages)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "SNPs/s", "cores": O.max_threads(), "kind": "port",
            "stage_s": {k: round(v, 4) for k, v in stages.items()},
            "sample": f"{rows} SNPs of the same job (count/QC + decode + f32 OpenBLAS rotation + OpenMP per-SNP "
                      f"{args.model} solve), {dt:.1f} s"}


def run_reference(args, torch, config, n, B, q, s_np, X_np, y_np, u_t_dev, u_t_host, device, have_gpu):
    from oracle import oracle as O
    O.build()
    ut = u_t_host if u_t_host is not None else u_t_dev.cpu().numpy()
    xcov, yrot = O.lmm_rotate_x_y_with_ut_f64(ut, X_np, y_np)
    y = yrot[:, 0].copy()
    lbd, ml0, reml0 = O.lmm_reml_null_f32(s_np, xcov, y, -5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd))
    low, high = l10 - 2.0, l10 + 2.0
    nullml = None
    if args.model == "lmm2":
        _, nullml = O.lmm_ml_null_brent(s_np, xcov, y, low, high, 30, 1e-2, l10)
    rows = min(args.cpu_sample, B)
    if have_gpu:
        packed = gen_snp_range(torch, n, 0, rows, device).cpu().numpy()      # the first SNPs of the same job
    else:
        from janusx_b200 import synth
        packed, _ = synth.draw_genotypes(rows, n, seed=SEED)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_scan(O, packed[: max(8, rows // 16)], n, s_np, xcov, y, ut, low, high, args.model, nullml, l10)
    stages = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_scan(O, packed, n, s_np, xcov, y, ut, low, high, args.model, nullml, l10, stages)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    cores = O.max_threads()
    cfg = dict(config)
    cfg["parallelism"] = f"{cores} host threads"
    cfg["sample_per_step"] = rows
    line = {"impl": "reference", "metric": "SNPs/sec exact -lmm scan (n=20k)", "value": value, "unit": "SNPs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 rotation / f64 solve",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "SNPs/s", "cores": cores, "kind": "port",
                             "stage_s": {k: round(v, 4) for k, v in stages.items()},
                             "sample": f"{rows} SNPs per step (bounded sample: the first SNPs of the {config.get('job_snps') or B}-SNP job)"},
            "e2e": {"value": value, "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "null_model": {"lambda": lbd, "front": "torch f64 matmul GRM + torch.linalg.eigh (untimed setup; no janusx_b200 code)"},
            "note": "CPU restatement of the reference algorithm (the Rust reference cannot be built in this image)"}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
