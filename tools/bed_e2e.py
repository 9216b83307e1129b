"""End-to-end file-level scan: synthetic PLINK files -> jxrs.lmm_reml_assoc_bed_to_tsv_f32 -> TSV (development tool).
usage: python tools/bed_e2e.py [n] [m] [model]"""
import json, os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench as B
from janusx_b200 import jxrs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
model = sys.argv[3] if len(sys.argv) > 3 else "lmm"
dev = torch.device("cuda:0")
tmp = Path(os.environ.get("TMPDIR", "/tmp")) / "jxb_e2e"
tmp.mkdir(exist_ok=True)
prefix = str(tmp / "panel")
t0 = time.time()
with open(prefix + ".bed", "wb") as fh:
    fh.write(bytes([0x6C, 0x1B, 0x01]))
    for b0 in range(0, m, 65536):
        rows = min(65536, m - b0)
        pk, _ = B.gen_packed_batch(torch, n, rows, b0 // 65536, dev)
        fh.write(pk.cpu().numpy().tobytes())
with open(prefix + ".bim", "w") as fh:
    fh.write("".join(f"1\tsnp{i}\t0\t{i}\tA\tT\n" for i in range(m)))
with open(prefix + ".fam", "w") as fh:
    fh.write("".join(f"F{j}\tS{j}\t0\t0\t0\t-9\n" for j in range(n)))
s_np, u_t_dev, X_np, y_np = B.build_null_model(torch, n, min(20000, m), 3, dev)
mdl = jxrs.DeviceModel(s_np, np.ones((n, 4)), np.zeros(n), u_t_dev, device=0, u_t_on_device=True)
xcov, yrot = mdl.rotate_xy(X_np, y_np)
mdl.set_xy(xcov, yrot[:, 0])
lbd, ml0, reml0 = mdl.reml_null(-5.0, 5.0, 50, 1e-3)
l10 = float(np.log10(lbd))
print(f"setup {time.time() - t0:.1f} s, lambda_null={lbd:.4f}", flush=True)
out = str(tmp / "out.tsv")
for rep in range(2):
    t1 = time.time()
    rows = mdl.scan_bed_to_tsv(prefix, out, 0.02, 0.05, 1.0, mode=model, low=l10 - 2, high=l10 + 2, init=l10,
                               batch_rows=jxrs.default_device_batch(n))
    dt = time.time() - t1
    print(json.dumps({"n": n, "m": m, "model": model, "rows": rows, "seconds": dt, "snps_per_s": m / dt,
                      "tsv_mb": os.path.getsize(out) / 1e6}), flush=True)
print(open(out).read(400))
