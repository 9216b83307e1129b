"""`jx gwas`-compatible command line for the exact-LMM path only (-lmm / -lmm2 / -fvlmm on -bfile / -vcf input).

Flag names, defaults and the output naming follow python/janusx/assoc/workflow.py:6599-7047
(`{out}/{prefix}.{trait}.{model}.tsv`, workflow_model_stream.py:992).  The scan itself is ONE call into the
B200 library per trait and rank (DeviceModel.scan_bed_to_tsv -> jxb_scan_bed_to_tsv), exactly where the reference makes
its one Rust call (workflow_model_stream.py:1449-1488).  Everything the reference CLI does outside this path (HMP/TXT
input, FarmCPU, LM, plots, run history) is out of scope; unsupported flags fail loudly.

  python -m janusx_b200.gwas -bfile panel -p pheno.tsv -n 0 -lmm -k 1 -q 3 -o out -prefix run1
  python -m janusx_b200.gwas -bfile panel -p pheno.tsv -lmm2 -gpus 8 -o out          # one job on 8 GPUs
  python -m torch.distributed.run --nproc-per-node 8 -m janusx_b200.gwas -bfile ...   # same, launched by torchrun

Multi-GPU (the reference shards SNPs over rayon workers, src/stats/reml.rs:88-99): one process per GPU.
Rank 0 builds the GRM (-k 1: int8 tensor-core kernel, csrc/grm.cu), decomposes it (csrc/eigh.cu) and fits the null
model; ONE broadcast ships (S, Xcov, y_rot, bounds, lambda, nullml) and U^T (4*n*n bytes, NCCL over NVLink) to every
rank; rank g scans the contiguous SNP range [g*m/G, (g+1)*m/G) of the BED into its own part file; rank 0 concatenates
the parts in rank order = BED order.  No data-path collective; per-SNP results do not depend on G (no warm start), so
the TSV is byte-identical for every GPU count.
"""
from __future__ import annotations

import argparse
import math
import os
import socket
import subprocess
import sys
import time
from typing import List, Optional

# The eigensolver (cusolverDnXsyevd) has a host stage whose summation order follows the host thread count, and
# torchrun exports OMP_NUM_THREADS=1 to its ranks: the same GRM would decompose to different last bits under
# `-gpus 1` and `-gpus 8`.  The rank that decomposes (rank 0) therefore always runs with the cores of its affinity
# mask, set before any threaded library is loaded, so every launch mode on one box yields the same null model and
# with it a byte-identical TSV.
if int(os.environ.get("RANK", "0")) == 0:
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))

import numpy as np


def parse_args(argv: Optional[List[str]] = None) -> argparse.Namespace:
    ap = argparse.ArgumentParser(prog="jx gwas (janusx_b200)", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    g = ap.add_argument_group("Genotype Arguments")
    g.add_argument("-bfile", "--bfile", default=None, help="PLINK prefix (.bed/.bim/.fam)")
    g.add_argument("-vcf", "--vcf", default=None, help="VCF / VCF.gz (GT); converted once to a PLINK cache next to the output")
    p = ap.add_argument_group("Phenotype Arguments")
    p.add_argument("-p", "--pheno", required=True, help="phenotype table: first column sample IDs, header row")
    p.add_argument("-n", "--ncol", action="append", default=None, help="zero-based trait column(s); default all")
    m = ap.add_argument_group("Model Arguments")
    m.add_argument("-lmm", "--lmm", action="store_true", default=False)
    m.add_argument("-lmm2", "--lmm2", action="store_true", default=False)
    m.add_argument("-fvlmm", "--fvlmm", action="store_true", default=False)
    o = ap.add_argument_group("Optional Arguments")
    o.add_argument("-k", "--grm", default="1", help="'1' = centred VanRaden GRM from the BED, or a .npy/.txt matrix")
    o.add_argument("-q", "--qcov", default="0", help="number of leading GRM PCs used as covariates")
    o.add_argument("-c", "--cov", default=None, help="covariate table (first column sample IDs)")
    o.add_argument("-maf", "--maf", type=float, default=0.02)
    o.add_argument("-geno", "--geno", type=float, default=0.05)
    o.add_argument("-het", "--het", type=float, default=1.0)
    o.add_argument("-model", "--model", default="add", choices=["add", "dom", "rec", "het"])
    o.add_argument("-snps-only", "--snps-only", action="store_true", default=False)
    o.add_argument("-force-model", "--force-model", action="store_true", default=False,
                   help="keep the mixed model even when the null LRT of Va = 0 is not significant "
                        "(workflow.py:6867; without it the reference switches such traits to LM, which this build does not hold)")
    o.add_argument("-t", "--thread", type=int, default=0, help="accepted for compatibility (CPU threads)")
    o.add_argument("-mem", "--mem", default=None,
                   help="host memory for the BED staging window, e.g. 4096, 4096MB, 4G (the reference's windowed mmap budget)")
    o.add_argument("-o", "--out", default=".")
    o.add_argument("-prefix", "--prefix", default=None)
    o.add_argument("-gpu", "--gpu", type=int, default=0, help="CUDA device index (single-GPU run)")
    o.add_argument("-gpus", "--gpus", type=int, default=1,
                   help="GPUs of this box to shard the SNPs over (spawns one process per GPU); under torchrun the world size wins")
    args = ap.parse_args(argv)
    if not (args.lmm or args.lmm2 or args.fvlmm):
        ap.error("select at least one of -lmm, -lmm2, -fvlmm (other models are outside this build's scope)")
    if (args.bfile is None) == (args.vcf is None):
        ap.error("give exactly one of -bfile, -vcf")
    args.mem_mb = _parse_mem_mb(args.mem, ap)
    return args


def _parse_mem_mb(text, ap) -> Optional[int]:
    """`-mem` -> MiB: a bare number is MiB (workflow.py help), suffixes K/M/G[B] are accepted."""
    if text is None:
        return None
    t = str(text).strip().upper().rstrip("B")
    mult = 1.0
    if t and t[-1] in "KMG":
        mult = {"K": 1.0 / 1024.0, "M": 1.0, "G": 1024.0}[t[-1]]
        t = t[:-1]
    try:
        v = float(t) * mult
    except ValueError:
        ap.error(f"-mem: cannot parse '{text}'")
    if not (v > 0):
        ap.error("-mem must be positive")
    return max(1, int(v))


def _vcf_cache(vcf: str, out_dir: str, snps_only: bool) -> str:
    """VCF -> PLINK cache (assoc/workflow.py:2431-2477: rebuilt when missing or older than the source)."""
    from . import jxrs
    base = os.path.basename(vcf)
    for ext in (".vcf.gz", ".vcf"):
        if base.lower().endswith(ext):
            base = base[: -len(ext)]
            break
    prefix = os.path.join(out_dir, f"~{base}.snp{1 if snps_only else 0}")
    targets = [prefix + e for e in (".bed", ".bim", ".fam")]
    fresh = all(os.path.isfile(t) for t in targets) and min(os.path.getmtime(t) for t in targets) >= os.path.getmtime(vcf)
    if not fresh:
        t0 = time.time()
        ns, nv = jxrs.vcf_to_plink(vcf, prefix, snps_only)
        print(f"[vcf] {vcf}: {ns} samples x {nv} sites -> {prefix}.bed ({time.time() - t0:.2f} s)", file=sys.stderr)
    return prefix


def _read_table(path: str):
    """-> (ids, column names, f64 matrix with NaN for missing)."""
    import pandas as pd
    df = pd.read_csv(path, sep=None, engine="python")
    ids = df.iloc[:, 0].astype(str).tolist()
    vals = df.iloc[:, 1:].apply(pd.to_numeric, errors="coerce")
    return ids, [str(c) for c in vals.columns], vals.to_numpy(dtype=np.float64)


def _read_fam(prefix: str) -> List[str]:
    with open(prefix + ".fam") as fh:
        return [line.split()[1] for line in fh if line.strip()]


def _bed_snp_count(prefix: str, n_full: int) -> int:
    size = os.path.getsize(prefix + ".bed")
    bps = (n_full + 3) // 4
    if size < 3 or (size - 3) % bps:
        raise SystemExit(f"{prefix}.bed: payload is not a multiple of {bps} bytes per SNP")
    return (size - 3) // bps


def _grm_from_bed(prefix: str, n_full: int, device: int, maf: float, geno: float, het: float) -> np.ndarray:
    """Centred VanRaden GRM (src/stats/grm.rs:204-608) over the SNPs passing the scan's QC thresholds, accumulated on
    the device by the int8 tensor-core kernel (csrc/grm.cu) from the memory-mapped BED."""
    from . import jxrs
    bps = (n_full + 3) // 4
    raw = np.memmap(prefix + ".bed", dtype=np.uint8, mode="r")
    if raw.shape[0] < 3 or bytes(raw[:3]) != b"\x6c\x1b\x01":
        raise SystemExit(f"{prefix}.bed: not a SNP-major PLINK BED file")
    packed = raw[3:3 + (raw.shape[0] - 3) // bps * bps].reshape(-1, bps)
    g = jxrs.DeviceGrm(n_full, None, 1, device)
    try:
        for r0 in range(0, packed.shape[0], 65536):
            g.update(np.ascontiguousarray(packed[r0:r0 + 65536]), None, qc=(maf, geno, het))
        if g.rows_used == 0:
            raise SystemExit("no SNP passed the QC thresholds: cannot build the GRM")
        k, _ = g.finish()
    finally:
        g.close()
    return k


def _spawn_ranks(n: int, argv: List[str]) -> int:
    """`-gpus N` outside torchrun: re-run this command as N ranks (one per GPU) on this box."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "janusx_b200.gwas", *argv]
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    return subprocess.call(cmd, env=env)


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    args = parse_args(argv)
    from . import dist as D
    rank, world, local = D.env_rank_world()
    if world == 1 and args.gpus > 1:
        return _spawn_ranks(args.gpus, argv)
    from . import _cabi, assoc, jxrs

    ndev = _cabi.require_gpu()
    device = args.gpu if world == 1 else local % ndev
    dist = None
    if world > 1:
        # NCCL needs one GPU per rank; ranks sharing a GPU (tests on a 1-GPU box) rendezvous over gloo instead
        dist = D.init_process_group("nccl" if ndev >= world else "gloo", device_index=device)
    log = (lambda *a: print(*a, file=sys.stderr)) if rank == 0 else (lambda *a: None)

    t0 = time.time()
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        if args.vcf is not None:
            args.bfile = _vcf_cache(args.vcf, args.out, args.snps_only)
    if world > 1:
        box = [args.bfile]
        dist.broadcast_object_list(box, src=0)      # the cache prefix (and the barrier behind the conversion)
        args.bfile = box[0]
    elif args.vcf is not None and args.bfile is None:
        args.bfile = _vcf_cache(args.vcf, args.out, args.snps_only)
    fam = _read_fam(args.bfile)
    n_snps = _bed_snp_count(args.bfile, len(fam))
    ids_p, traits, Y = _read_table(args.pheno)
    cols = list(range(len(traits))) if not args.ncol else [int(c) for tok in args.ncol for c in str(tok).split(",")]
    prefix = args.prefix or os.path.basename(args.vcf or args.bfile).replace(".vcf.gz", "").replace(".vcf", "")
    outprefix = os.path.join(args.out, prefix)

    K_full = None
    if rank == 0:
        if args.grm == "1":
            K_full = _grm_from_bed(args.bfile, len(fam), device, args.maf, args.geno, args.het)
        elif args.grm.endswith(".npy"):
            K_full = np.load(args.grm)
        else:
            K_full = np.loadtxt(args.grm)
        if K_full.shape != (len(fam), len(fam)):
            raise SystemExit(f"GRM shape {K_full.shape} does not match {len(fam)} FAM samples")
    cov_ids, cov_mat = None, None
    if args.cov:
        cov_ids, _, cov_mat = _read_table(args.cov)
    pos_fam = {sid: i for i, sid in enumerate(fam)}
    models = [m for m, on in (("lmm", args.lmm), ("lmm2", args.lmm2), ("fvlmm", args.fvlmm)) if on]
    skipped = 0

    for c in cols:
        trait = traits[c]
        y_all = Y[:, c]
        ok = [i for i, sid in enumerate(ids_p) if sid in pos_fam and np.isfinite(y_all[i])]
        if cov_mat is not None:
            pos_cov = {sid: i for i, sid in enumerate(cov_ids)}
            ok = [i for i in ok if ids_p[i] in pos_cov and np.all(np.isfinite(cov_mat[pos_cov[ids_p[i]]]))]
        # keep FAM order (the reference sorts kept samples by genotype order)
        ok.sort(key=lambda i: pos_fam[ids_p[i]])
        sample_ids = [ids_p[i] for i in ok]
        n = len(ok)
        nq = int(args.qcov) if str(args.qcov).isdigit() else 0
        p_cols = 1 + nq + (cov_mat.shape[1] if cov_mat is not None else 0)

        # ---- null model on rank 0 (eigh on K + 1e-6 I, rotation, REML null fit: pyBLUP/assoc.py:1595-1876) ----------
        nm, go = None, True
        if rank == 0:
            fidx = np.array([pos_fam[s] for s in sample_ids], dtype=np.int64)
            y = y_all[ok]
            K = K_full[np.ix_(fidx, fidx)]
            X_parts = []
            if cov_mat is not None:
                X_parts.append(np.stack([cov_mat[pos_cov[s]] for s in sample_ids]))
            t_null = time.time()
            if nq > 0:
                # the PCs and the null model take the same decomposition: K + 1e-6 I is the matrix LMM() itself decomposes
                # (pyBLUP/assoc.py:1626-1629), so it is decomposed once and handed over (LMM.from_spectral, assoc.py:1726)
                evals, evecs = assoc._eigh(K + 1e-6 * np.eye(n), device)
                evd_s = time.time() - t_null
                X_parts.append(evecs[:, ::-1][:, :nq] * np.sqrt(np.maximum(evals[::-1][:nq], 0.0)))
                X_cov = np.concatenate(X_parts, axis=1)
                base = assoc.LMM.from_spectral(y, X_cov, evals, evecs, evd_secs=evd_s, device=device)
            else:
                X_cov = np.concatenate(X_parts, axis=1) if X_parts else None
                base = assoc.LMM(y, X_cov, K, device=device)
            l10 = float(np.log10(base.lbd_null))
            log(f"[{trait}] n={n} covariates={base.Xcov.shape[1]} lambda_null={base.lbd_null:.6g} "
                f"pve={base.pve:.4f} bounds=({base.bounds[0]:.3f},{base.bounds[1]:.3f}) null model {time.time() - t_null:.2f} s")
            # mixed model -> LM switch (workflow_model_stream.py:930-963; src/stats/gwas_unified.rs:119-175)
            if not args.force_model:
                sw, stat, pv, _ = jxrs.gwas_lmm_lm_null_lrt_decision(y, X_cov if X_cov is not None else np.zeros((n, 0)),
                                                                      base.ML0, 0.05, True)
                if sw:
                    log(f"Warning: switch to LM for trait {trait}: null LRT stat={stat:.4g}, p={pv:.4g} (>=0.05). The "
                        f"reference continues with LM here; this build holds the mixed-model path only -- trait skipped, "
                        f"rerun with -force-model to scan it with the mixed model.")
                    go = False
            nullml = float("nan")
            if go and "lmm2" in models:                      # src/stats/lmm.rs:2901-2924, seeded like the CLI
                _, nullml = base.device_model.ml_null(float(base.bounds[0]), float(base.bounds[1]), 30, 1e-2, l10)
            nm = D.NullModel(s=base.S, xcov=base.Xcov, y=base.y[:, 0].copy(), u_t=base.Dh, low=float(base.bounds[0]),
                             high=float(base.bounds[1]), lbd_null=float(base.lbd_null), nullml=nullml)
            if os.environ.get("JXB_DEBUG_DUMP_NULL"):      # development aid: the null model this run scanned with
                np.savez(os.environ["JXB_DEBUG_DUMP_NULL"], s=nm.s, xcov=nm.xcov, y=nm.y, u_t=nm.u_t, low=nm.low, high=nm.high,
                         lbd=nm.lbd_null, nullml=nm.nullml, **({"K": K} if n <= 5000 else {}))
        if world > 1:
            flag = [go]
            dist.broadcast_object_list(flag, src=0)
            go = flag[0]
        if not go:
            skipped += 1
            continue
        t_b = time.time()
        if world > 1:
            nm = D.broadcast_null_model(nm, n, p_cols, src=0)
            mdl = jxrs.DeviceModel(nm.s, nm.xcov, nm.y, nm.u_t, device=device, u_t_on_device=hasattr(nm.u_t, "data_ptr"))
            log(f"[{trait}] null model broadcast to {world} ranks + upload {time.time() - t_b:.2f} s")
        else:
            mdl = base.device_model
        l10 = float(np.log10(nm.lbd_null))
        sids = None if sample_ids == fam else sample_ids
        for model in models:
            out_tsv = f"{outprefix}.{trait}.{model}.tsv" if args.model == "add" else f"{outprefix}.{trait}.{args.model}.{model}.tsv"
            tmp = out_tsv + ".tmp"
            t1 = time.time()
            kw = dict(genetic_model=args.model, snps_only=args.snps_only, sample_ids=sids, mode=model, low=nm.low, high=nm.high,
                      max_iter=30, tol=1e-2, batch_rows=jxrs.default_device_batch(n), mmap_window_mb=args.mem_mb)
            if model == "lmm":
                kw.update(init=l10)
            elif model == "lmm2":
                kw.update(init=l10, nullml=nm.nullml)
            else:
                kw.update(log10_lbd=l10)

            def scan_range(b, e, part, header):
                if e <= b:
                    open(part, "wb").close() if not header else _write_header_only(part, mdl, model, kw)
                    return 0
                return mdl.scan_bed_to_tsv(args.bfile, part, args.maf, args.geno, args.het, snp_begin=b, snp_end=e,
                                           write_header=header, **kw)

            rows = D.scan_bed_sharded(args.bfile, tmp, n_snps, scan_range)
            if rank == 0:
                os.replace(tmp, out_tsv)   # atomic rename like workflow.py:833-846
            log(f"[{trait}] {model}: {rows} SNPs -> {out_tsv} ({time.time() - t1:.2f} s on {world} GPU{'s' if world > 1 else ''})")
        if world > 1:
            mdl.close()
    log(f"done in {time.time() - t0:.2f} s")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 3 if skipped else 0


def _write_header_only(path, mdl, model, kw):
    from ._cabi import lib
    cols = 6 if model == "lmm2" else (4 if kw.get("nullml") is not None else 3)
    with open(path, "wb") as fh:
        fh.write(lib().jxb_tsv_header(cols))


if __name__ == "__main__":
    sys.exit(main())
