"""GPU probe: cuBLAS DGEMM ceiling and the rotation / solve / decode kernels timed in isolation.
Writes gpurun_out/probe.json.  (Development tool; numbers quoted in DESIGN.md come from bench.py.)"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from janusx_b200 import _cabi, jxrs  # noqa: E402


def dgemm_peak(n, m=4096, reps=5):
    a = torch.randn((m, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    best = 0
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b.T; e1.record(); torch.cuda.synchronize()
        if i:
            best = max(best, 2.0 * m * n * n / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    ns = [int(x) for x in os.environ.get("PROBE_NS", "2000,5000,20000").split(",")]
    rows = int(os.environ.get("PROBE_ROWS", 8192))
    lib = _cabi.lib()
    lib.jxb_set_timing(1)
    for n in ns:
        rec = {"dgemm_tflops": dgemm_peak(n)}
        rng = np.random.default_rng(n)
        # orthogonal-ish U^T is irrelevant for timing; correctness is checked against torch f64 matmul
        ut = torch.randn((n, n), dtype=torch.float32, device="cuda") / np.sqrt(n)
        s = np.abs(rng.normal(size=n)) + 0.5
        X = np.concatenate([np.ones((n, 1)), rng.normal(size=(n, 3))], axis=1)
        y = rng.normal(size=n)
        mdl = jxrs.DeviceModel(s, X, y, ut, device=0, u_t_on_device=True)
        bps = (n + 3) // 4
        packed = torch.randint(0, 256, (rows, bps), dtype=torch.uint8, device="cuda")
        # forbid the missing code so QC keeps every row: map code 01 -> 00
        lo = packed & 0x55
        hi = (packed >> 1) & 0x55
        packed = (packed & ~(lo & ~hi)).contiguous()
        kw = dict(maf_thr=0.0, miss_thr=1.0, het_thr=1.0, mode="lmm2", low=-2.0, high=2.0, init=0.0, nullml=-1e4)
        for variant in [int(v) for v in os.environ.get('PROBE_VARIANTS', '0,3,2,1').split(',')]:
            if variant == 1 and n > 6000:
                continue
            lib.jxb_set_rotate_variant(variant)
            for it in range(3):
                mdl.scan_packed_dev(int(packed.data_ptr()), rows, bps, n, None, **kw)
                mdl.sync()
            keep, af, missing, res, ev = mdl.scan_fetch(rows, 6)
            st = mdl.stage_ms()
            kept = int(keep.sum())
            rec[f"variant{variant}"] = {"stage_ms": st, "kept": kept,
                                        "rotate_tflops": 2.0 * n * n * kept / (st["rotate"] * 1e-3) / 1e12,
                                        "snps_per_s": kept / (sum(st[k] for k in ("count_qc", "decode", "rotate", "solve")) * 1e-3),
                                        "mean_evals": float(ev.mean())}
        lib.jxb_set_rotate_variant(3)
        # rotation correctness at this size against torch f64 (first 256 rows)
        g = torch.randn((256, n), dtype=torch.float32, device="cuda")
        want = (g.double() @ ut.double().T).float().cpu().numpy()
        got = mdl.rotate_block(g.cpu().numpy(), variant=0)
        rec["rot_max_abs_err_vs_torch_f64"] = float(np.abs(got - want).max())
        rec["rot_frac_diff"] = float((got != want).mean())
        mdl.close()
        out[f"n{n}"] = rec
        print(n, json.dumps(rec), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/probe.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
