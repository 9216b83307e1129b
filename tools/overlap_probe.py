"""Does the tensor-pipe rotation of one batch overlap the FP64 solve of another on the same SMs?

Two independent device models ("lanes", own stream + workspace) scan batches concurrently from two host threads;
throughput is compared with one lane alone.  Round-1 finding (profiles/README.md, "K2 || K3 overlap"): no gain --
with the shipped kernels the two cannot share an SM (registers + shared memory), and with both kernels slimmed to
co-reside (rotation 128 registers / 2 stages, solve 152 registers) a per-CTA trace showed real co-residency but only
one solve CTA beside a rotation CTA and the two lanes' rotations running back to back; 548-582 ms per batch against
523-542 ms for one lane of the same build and 479 ms for the shipped build.
"""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import bench as B
    from janusx_b200 import jxrs
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 56832
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    dev = torch.device("cuda:0")
    q = 3
    s_np, u_t_dev, X_np, y_np = B.build_null_model(torch, n, 20000, q, dev)
    mdls = []
    for _ in range(lanes):
        m = jxrs.DeviceModel(s_np, np.ones((n, q + 1)), np.zeros(n), u_t_dev, device=0, u_t_on_device=True)
        if not mdls:
            xcov, yrot = m.rotate_xy(X_np, y_np)
        m.set_xy(xcov, yrot[:, 0])
        mdls.append(m)
    lbd, ml0, reml0 = mdls[0].reml_null(-5.0, 5.0, 50, 1e-3)
    l10 = float(np.log10(lbd))
    low, high = l10 - 2.0, l10 + 2.0
    _, nullml = mdls[0].ml_null(low, high, 30, 1e-2, l10)
    del u_t_dev
    torch.cuda.empty_cache()
    bps = (n + 3) // 4
    bufs = [B.gen_packed_batch(torch, n, batch, i, dev)[0] for i in range(2 * lanes)]
    torch.cuda.synchronize()
    kw = dict(maf_thr=0.02, miss_thr=0.05, het_thr=1.0, genetic_model="add", mode="lmm2", low=low, high=high,
              max_iter=30, tol=1e-2, init=l10, nullml=nullml, log10_lbd=l10)

    def run(mdl, lane, count):
        for i in range(count):
            pk = bufs[(2 * lane + i) % len(bufs)]
            mdl.scan_packed_dev(pk.data_ptr(), batch, bps, n, None, **kw)
        mdl.sync()

    out = {"n": n, "batch": batch, "steps_per_lane": steps, "lanes": lanes}
    for m in mdls:          # warm-up (allocations, tensor maps, slices)
        run(m, 0, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(mdls[0], 0, steps)
    out["one_lane_ms_per_batch"] = (time.perf_counter() - t0) * 1e3 / steps
    ths = [threading.Thread(target=run, args=(m, i, steps)) for i, m in enumerate(mdls)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    out["multi_lane_ms_per_batch"] = (time.perf_counter() - t0) * 1e3 / (steps * lanes)
    out["gain"] = out["one_lane_ms_per_batch"] / out["multi_lane_ms_per_batch"]
    out["snps_per_s_multi"] = batch / out["multi_lane_ms_per_batch"] * 1e3
    # results must not depend on the lane
    k0 = mdls[0].scan_fetch(batch, 6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
