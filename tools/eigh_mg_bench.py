"""Eigendecomposition timings: cusolverDnXsyevd on one GPU vs cusolverMgSyevd on 1..k GPUs (the Amdahl term of the
multi-GPU job), and the n > 46,340 case.  usage: python tools/eigh_mg_bench.py [n] [devices,comma] [out.json]"""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from janusx_b200 import _cabi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    devs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2").split(",")]    # 0 = cusolverDn
    out_json = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/r2_eigh.json"
    lib = _cabi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    rows = 4096
    a = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for _ in range(4):                                   # low-rank-plus-ridge SPD matrix with a GRM-like spectrum
        z = torch.randn((rows, n), generator=g, device=dev, dtype=torch.float64)
        a += z.T @ z
        del z
    a /= float(4 * rows)
    a.diagonal().add_(1e-3)
    res = {"n": n, "runs": []}
    ref_w = None
    for k in devs:
        if k > torch.cuda.device_count():
            continue
        work = a.clone()
        w = torch.empty(n, dtype=torch.float64, device=dev)
        u32 = torch.empty((n, n), dtype=torch.float32, device=dev)
        if k == 0:
            if n > 46340:
                del work, w, u32
                continue
            os.environ.pop("JXB_EIGH_FORCE_MG", None)
            lib.jxb_set_eigh_devices(0)
        else:
            os.environ["JXB_EIGH_FORCE_MG"] = "1"
            lib.jxb_set_eigh_devices(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.jxb_eigh_dev(0, n, int(work.data_ptr()), C.c_double(0.0), int(w.data_ptr()), int(u32.data_ptr()), None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        entry = {"devices": k, "solver": "cusolverDnXsyevd" if k == 0 else "cusolverMgSyevd", "seconds": round(dt, 3), "rc": rc}
        if rc:
            entry["error"] = lib.jxb_last_error().decode()
        else:
            wh = w.cpu().numpy()
            if ref_w is None:
                ref_w = wh
            entry["max_rel_eval_diff_vs_first"] = float(np.max(np.abs(wh - ref_w)) / np.max(np.abs(ref_w)))
            # residual of a few eigenpairs: || A u - w u || / |w|
            idx = [0, n // 2, n - 1]
            u = work[idx].to(torch.float64)           # rows of U^T
            r = (a @ u.T) - u.T * w[idx]
            entry["max_resid"] = float((r.norm(dim=0) / w[idx].abs()).max())
        res["runs"].append(entry)
        print(entry, flush=True)
        del work, w, u32
        torch.cuda.empty_cache()
    Path(out_json).parent.mkdir(exist_ok=True)
    Path(out_json).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
