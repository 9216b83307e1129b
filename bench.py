#!/usr/bin/env python
"""bench.py -- SNPs/s of the exact LMM scan (decode -> rotate -> per-SNP REML/ML solve) on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1) prints ONE JSON line
on rank 0.  A step = one pass of the hot path over one batch of `--batch` synthetic SNP rows at the
workload BASELINE.json's metric is quoted on: n=20,000 samples, -lmm2 (Wald + LRT), 1 trait, 3 covariates
(configs[2]; m=1,000,000 is the job size -- steps sample batches of it).  `value` is whole-job SNPs/s with
packed genotypes resident in HBM; `e2e` is the same metric through the reference-facing C-ABI call
jxb_scan_packed with pinned HOST buffers (H2D of the packed batch and D2H of the result rows inside the
timed region).  `--impl reference` times the CPU restatement of the reference algorithm (oracle port:
numpy/OpenBLAS f32 rotation like the reference's cblas_sgemm + OpenMP per-SNP solve) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260609  # the reference's benchmark seed (scripts/benchmark.sh:36)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("JXB_BENCH_N", 20000)))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("JXB_BENCH_BATCH", 0)),
                    help="SNPs per step; default = the library's device batch: 151,552 (two waves of 148 SMs x 16 warps x "
                         "32 SNPs of the thread-per-SNP solve) for n <= 24,000, else 75,776")
    ap.add_argument("--model", default=os.environ.get("JXB_BENCH_MODEL", "lmm2"), choices=["lmm", "lmm2", "fvlmm"])
    ap.add_argument("--grm-snps", type=int, default=int(os.environ.get("JXB_BENCH_GRM_SNPS", 50000)))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("JXB_BENCH_CPU_SAMPLE", 1024)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rotate-variant", type=int, default=int(os.environ.get("JXB_BENCH_ROTATE", 3)),
                    help="3 = hand-written tcgen05 int8-sliced exact rotation (default), 2 = same via cuBLASLt, "
                         "0 = FP64 DMMA GEMM")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic inputs on the GPU (distributions of `jx sim`, python/janusx/script/sim.py:49-66, 133-176, 252-276)
# ------------------------------------------------------------------------------------------------------
def gen_packed_batch(torch, n, rows, batch_index, device, want_dosage=False):
    """HWE genotypes for `rows` SNPs keyed by (SEED, global batch index) so every GPU count sees the same data."""
    g = torch.Generator(device=device)
    g.manual_seed(SEED * 1000003 + int(batch_index))
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


def build_null_model(torch, n, grm_snps, q, device, timings=None):
    """Null-model inputs through the library's own front steps: centred VanRaden GRM of `grm_snps` synthetic SNPs on
    the int8 tensor cores (csrc/grm.cu; src/stats/grm.rs:204-608), K + 1e-6 I decomposed in place by cuSOLVER
    (csrc/eigh.cu; workflow_model_stream.py:902).  Phenotype 100 + G beta + e at pve 0.5, q N(0,1)
    covariates.  Returns host arrays (s, X design, y) and U^T as an f32 device tensor (pyBLUP/assoc.py:1818)."""
    from janusx_b200 import jxrs
    timings = {} if timings is None else timings
    if n > 46340:
        # cuSOLVER's Xsyevd rejects n*n >= 2^31: model two unrelated populations (block-diagonal GRM and U^T)
        h = n // 2
        s1, u1, X1, y1 = build_null_model(torch, h, grm_snps, q, device, timings)
        s2, u2, X2, y2 = build_null_model(torch, n - h, max(1024, grm_snps // 2), q, device)
        u_t = torch.zeros((n, n), dtype=torch.float32, device=device)
        u_t[:h, :h] = u1
        u_t[h:, h:] = u2
        timings["blocks"] = 2
        return np.concatenate([s1, s2]), u_t, np.concatenate([X1, X2]), np.concatenate([y1, y2])
    grm = jxrs.DeviceGrm(n, None, 1, device.index or 0)
    gv = torch.zeros(n, dtype=torch.float64, device=device)
    gt = torch.Generator(device=device)
    gt.manual_seed(SEED + 1)
    chunk = 16384
    t_grm = 0.0
    for b, r0 in enumerate(range(0, grm_snps, chunk)):
        rows = min(chunk, grm_snps - r0)
        packed, dos = gen_packed_batch(torch, n, rows, 10_000_000 + b, device, want_dosage=True)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        grm.update_dev(packed.data_ptr(), rows, packed.shape[1])
        t_grm += time.perf_counter() - t0
        beta = torch.randn(rows, generator=gt, device=device, dtype=torch.float64)
        for c0 in range(0, rows, 4096):
            z = dos[c0:c0 + 4096].to(torch.float64)
            z -= z.mean(dim=1, keepdim=True)
            gv += beta[c0:c0 + 4096] @ z
        del z, dos, packed
    s = torch.empty(n, dtype=torch.float64, device=device)
    u_t = torch.empty((n, n), dtype=torch.float32, device=device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    grm.eigh_dev(s.data_ptr(), u_t.data_ptr(), 1e-6)
    timings["eigh_s"] = time.perf_counter() - t0
    timings["grm_s"] = t_grm
    timings["grm_snps"] = grm_snps
    grm.close()
    vg = float(gv.var(unbiased=False))
    ve = vg  # pve 0.5
    y = 100.0 + gv + torch.randn(n, generator=gt, device=device, dtype=torch.float64) * (ve ** 0.5)
    gc = torch.Generator(device=device)
    gc.manual_seed(SEED + 2)
    cov = torch.randn((n, q), generator=gc, device=device, dtype=torch.float64)
    X = torch.cat([torch.ones((n, 1), dtype=torch.float64, device=device), cov], dim=1)
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


def measure_fp64_peak(torch, device, n):
    """cuBLAS DGEMM ceiling for the rotation shape (M=4096, N=K=n): burst (best of 5) TFLOP/s."""
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
    p = q + 1
    config = {"workload": f"synthetic n={n}, m=1,000,000 job sampled in batches of {B} SNPs, -{args.model} "
                          f"(Wald{' + LRT' if args.model == 'lmm2' else ''}), 1 trait, {q} covariates "
                          "(BASELINE.json configs[2])",
              "n": n, "batch_snps": B, "covariates": q, "model": args.model,
              "parallelism": f"snp-shard x{world}" if args.impl == "b200" else "host threads",
              "l2": "inputs larger than L2: every step streams the 8*n*n-byte U^T (3.2 GB at n=20k)"}

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

    # ---- null model: eigh once on rank 0, NCCL broadcast of U^T (f32), S, X_rot, y_rot ----------------
    t_setup = time.time()
    setup_timings = {}
    if args.impl == "reference" and not have_gpu:
        # CPU-only host: small orthogonal basis from numpy (the reference arm still times the same algorithm)
        rng = np.random.default_rng(SEED)
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
            s_np, u_t_dev, X_np, y_np = build_null_model(torch, n, args.grm_snps, q, device, setup_timings)
        else:
            s_np = np.zeros(n); X_np = np.zeros((n, p)); y_np = np.zeros(n)
            u_t_dev = torch.empty((n, n), dtype=torch.float32, device=device)
        if dist is not None:
            small = torch.as_tensor(np.concatenate([s_np, X_np.reshape(-1), y_np]), device=device)
            dist.broadcast(small, 0)
            dist.broadcast(u_t_dev, 0)      # 4*n*n bytes over NVLink, once
            sm = small.cpu().numpy()
            s_np, X_np, y_np = sm[:n].copy(), sm[n:n + n * p].reshape(n, p).copy(), sm[n + n * p:].copy()
        u_t_host = None

    if args.impl == "reference":
        return run_reference(args, torch, config, n, B, q, s_np, X_np, y_np, u_t_dev, u_t_host, device, have_gpu)

    from janusx_b200 import _cabi, jxrs
    lib = _cabi.lib()
    lib.jxb_set_rotate_variant(args.rotate_variant)
    # rotate X, y and fit the null on the device (pyBLUP/assoc.py:1818-1876)
    mdl = jxrs.DeviceModel(s_np, np.ones((n, p)), np.zeros(n), u_t_dev, device=local_rank, u_t_on_device=True)
    xcov, yrot = mdl.rotate_xy(X_np, y_np)
    mdl.set_xy(xcov, yrot[:, 0])
    lbd, ml0, reml0 = mdl.reml_null(-5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd))
    low, high = l10 - 2.0, l10 + 2.0
    nullml = None
    if args.model == "lmm2":
        _, nullml = mdl.ml_null(low, high, 30, 1e-2, l10)       # src/stats/lmm.rs:2901-2924
    fp64_peak = measure_fp64_peak(torch, device, n) if rank == 0 else 0.0
    ut_host = u_t_dev.cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    del u_t_dev
    torch.cuda.empty_cache()

    # ---- synthetic SNP batches: this rank's shard, resident in HBM + a pinned host copy for e2e --------
    n_bufs = min(4, max(1, args.steps))
    bps = (n + 3) // 4
    dev_batches, host_batches = [], []
    for i in range(n_bufs):
        gidx = rank * 1000 + i   # contiguous SNP range per rank: rank r owns batches [r*1000, r*1000+...)
        pk, _ = gen_packed_batch(torch, n, B, gidx, device)
        dev_batches.append(pk)
        hb = torch.empty((B, bps), dtype=torch.uint8, pin_memory=True)
        hb.copy_(pk)
        host_batches.append(hb)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    scan_kw = dict(maf_thr=0.02, miss_thr=0.05, het_thr=1.0, genetic_model="add", mode=args.model, low=low, high=high,
                   max_iter=30, tol=1e-2, init=(l10 if args.model == "lmm2" else None), nullml=nullml, log10_lbd=l10)
    cols = 6 if args.model == "lmm2" else 3

    def step_resident(i):
        pk = dev_batches[i % n_bufs]
        mdl.scan_packed_dev(int(pk.data_ptr()), B, bps, n, None, **scan_kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(mdl.stream, device=device)
    lib.jxb_set_timing(1)
    for i in range(args.warmup):
        step_resident(i)
    mdl.sync()
    # timed region: kernels only, inputs resident (CUDA events on the model's own stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rot_ms, solve_ms, dec_ms, cnt_ms, kept = [], [], [], [], 0
    e0.record(stream)
    for i in range(args.steps):
        step_resident(i)
    e1.record(stream)
    mdl.sync()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = _cabi.launch_count() - launches0
    # per-kernel durations of the LAST timed step (events recorded inside the library on the same stream)
    keep, af, missing, out, evals = mdl.scan_fetch(B, cols)
    st = mdl.stage_ms()
    kept = int(keep.sum())
    mean_evals = float(evals.mean()) if evals.size and args.model != "fvlmm" else 0.0

    # e2e: same metric through the C-ABI call with pinned HOST buffers (H2D + D2H inside the timed region)
    for i in range(min(2, args.warmup)):
        mdl.scan_packed(host_batches[i % n_bufs].numpy(), n, **{k: v for k, v in scan_kw.items()})
    barrier()
    t0 = time.perf_counter()
    kept_e2e = 0
    for i in range(args.steps):
        r = mdl.scan_packed(host_batches[i % n_bufs].numpy(), n, **{k: v for k, v in scan_kw.items()})
        kept_e2e += int(r[0].sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    tvec = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(tvec[0]), float(tvec[1])
    total_snps = world * args.steps * B
    value = total_snps / (dev_ms * 1e-3)
    e2e_value = total_snps / (e2e_ms * 1e-3)

    if rank == 0:
        rot_s = st["rotate"] * 1e-3
        rot_flop = 2.0 * n * n * kept                      # FP64-equivalent work of the rotation
        d = p + 1
        flop_per_eval = n * (3 * d * (d + 1) / 2 + 5 * d + 3 + 2)   # + 1 divide + 1 log per sample
        solve_flop = mean_evals * flop_per_eval * kept
        solve_s = st["solve"] * 1e-3
        fp64_core_peak = 36.0      # TFLOP/s: 18.0 T DFMA/s measured with tools/fp64_probe.cu (profiles/r1_fp64_probe.txt)
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture of this
        # exact default workload (profiles/r1_ncu_final_metrics.csv); null for any other configuration
        # (solve: profiles/r1_ncu_lane_final.csv at 151,460 kept SNPs; rotation: r1_ncu_final_metrics.csv at 75,734;
        # bytes scale with the kept SNPs of the step)
        default_cfg = (n == 20000 and B % 75776 == 0 and args.model == "lmm2" and args.rotate_variant == 3 and q == 3)
        solve_traffic = (920233095680 + 210615296) * (kept / 151460.0) if default_cfg else None
        rot_traffic = (76375571200 + 12103506432 + 171910203648 + 6066589696) * (kept / 75734.0) if default_cfg else None
        rot_roof = {"kernel": ("rotate_dmma_kernel (FP64 DMMA GEMM)" if args.rotate_variant == 0 else
                               ("i8_rotate_kernel (tcgen05 int8-sliced exact rotation, 2 passes)" if args.rotate_variant == 3 else
                                "int8-sliced exact rotation (10 cuBLASLt slice GEMMs + recombine_kernel)")),
                    "bound": "tensor", "achieved": rot_flop / rot_s / 1e12 if rot_s > 0 else 0.0, "peak": fp64_peak,
                    "unit": "TFLOP/s (FP64-equivalent)", "frac": (rot_flop / rot_s / 1e12 / fp64_peak) if rot_s > 0 and fp64_peak > 0 else None,
                    "traffic": rot_traffic, "launch_ms": st["rotate"],
                    "peak_source": "cuBLAS DGEMM (M=4096,N=K=n) measured in this run; MEASURED_PEAKS.json holds no FP64 "
                                   "figure. frac > 1 for the int8-sliced variant: same f64-accurate result from 10 exact "
                                   "int8 slice GEMMs (" + f"{10 * rot_flop / rot_s / 1e12:.0f}" + " int8 TOP/s of 4500 nominal)"}
        solve_roof = {"kernel": (f"solve_lane_kernel<{p}>" if (args.rotate_variant >= 2 and kept >= 32768) else f"solve_warp_kernel<{p}>") + " (per-SNP REML/ML Brent, FP64 CUDA cores)", "bound": "fp64-cuda-core",
                      "achieved": solve_flop / solve_s / 1e12 if solve_s > 0 else 0.0, "peak": fp64_core_peak,
                      "unit": "TFLOP/s", "frac": (solve_flop / solve_s / 1e12 / fp64_core_peak) if solve_s > 0 else None,
                      "traffic": solve_traffic, "launch_ms": st["solve"],
                      "algorithmic_flop_per_launch": solve_flop,
                      "algorithmic_bytes_per_launch": (mean_evals * 2.0 + 1.0) * kept * n * 4.0,
                      "traffic_source": "profiles/r1_ncu_lane_final.csv (ncu, same command); algorithmic bytes = the f32 "
                                        "rotated block streamed twice per objective evaluation (sums pass + residual pass) "
                                        "plus once for the validity check",
                      "peak_source": "2 x 18.0 T DFMA/s measured on this pool's B200 (tools/fp64_probe.cu); the algorithmic "
                                     "count charges 1 flop per divide / log, which cost 9.5 / 51 DFMA-equivalents"}
        dominant = solve_roof if st["solve"] >= st["rotate"] else rot_roof
        line = {
            "metric": "SNPs/sec exact -lmm scan (n=20k)", "value": value, "unit": "SNPs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config, rotation={0: "fp64-dmma", 1: "fp64-cuda-core", 2: "int8-sliced-exact (cuBLASLt)", 3: "int8-sliced-exact (tcgen05)"}[args.rotate_variant]),
            "e2e": {"value": e2e_value, "unit": "SNPs/s", "h2d_bytes_per_step": B * bps,
                    "d2h_bytes_per_step": kept_e2e // max(args.steps, 1) * (cols * 8 + 4) + B * (16 + 4) + 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": dominant, "roofline_rotation": rot_roof, "roofline_solve": solve_roof,
            "stage_ms_last_step": st,
            "solve": {"mean_objective_evals_per_snp": mean_evals},
            "decode": {"hbm_gb_per_s": ((B * bps + kept * n * (8 if args.rotate_variant == 0 else 3)) / (st["decode"] * 1e-3) / 1e9) if st["decode"] > 0 else None},
            "null_model": dict({"lambda": lbd, "ml0": ml0, "reml0": reml0, "setup_s": setup_s,
                                "front": "GRM: csrc/grm.cu (tcgen05 int8), eigh: csrc/eigh.cu (cuSOLVER Xsyevd)"},
                               **setup_timings),
            "kept_snps_per_step": kept, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, n, B, s_np, xcov, yrot[:, 0].copy(), ut_host,
                                                host_batches[0].numpy(), low, high, nullml, l10)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _ut_host_from_model(torch, n, u_t_dev, u_t_host):
    if u_t_host is not None:
        return u_t_host
    return u_t_dev.cpu().numpy()


def cpu_baseline(args, n, B, s_np, xcov, y, ut, packed_host, low, high, nullml, l10):
    """Oracle port timed on the box's host cores on a bounded SNP sample of the same workload."""
    from oracle import oracle as O
    O.build()
    rows = min(args.cpu_sample, B)
    stages = {}
    t0 = time.perf_counter()
    cpu_scan(O, packed_host[:rows], n, s_np, xcov, y, ut, low, high, args.model, nullml, l10, stages)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "SNPs/s", "cores": O.max_threads(), "kind": "port",
            "stage_s": {k: round(v, 4) for k, v in stages.items()},
            "sample": f"{rows} SNPs of the same batch (count/QC + decode + f32 OpenBLAS rotation + OpenMP per-SNP "
                      f"{args.model} solve), {dt:.1f} s"}


def run_reference(args, torch, config, n, B, q, s_np, X_np, y_np, u_t_dev, u_t_host, device, have_gpu):
    from oracle import oracle as O
    O.build()
    ut = _ut_host_from_model(torch, n, u_t_dev, u_t_host)
    xcov, yrot = O.lmm_rotate_x_y_with_ut_f64(ut, X_np, y_np)
    y = yrot[:, 0].copy()
    lbd, ml0, reml0 = O.lmm_reml_null_f32(s_np, xcov, y, -5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd))
    low, high = l10 - 2.0, l10 + 2.0
    nullml = None
    if args.model == "lmm2":
        _, nullml = O.lmm_ml_null_brent(s_np, xcov, y, low, high, 30, 1e-2, l10)
    rows = min(args.cpu_sample, B)
    from janusx_b200 import synth
    if have_gpu:
        pk, _ = gen_packed_batch(torch, n, rows, 0, device)
        packed = pk.cpu().numpy()
    else:
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
    line = {"impl": "reference", "metric": "SNPs/sec exact -lmm scan (n=20k)", "value": value, "unit": "SNPs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 rotation / f64 solve",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "SNPs/s", "cores": cores, "kind": "port",
                             "stage_s": {k: round(v, 4) for k, v in stages.items()},
                             "sample": f"{rows} SNPs per step (bounded sample of the {B}-SNP batch)"},
            "e2e": {"value": value, "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference algorithm (the Rust reference cannot be built in this image)"}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
