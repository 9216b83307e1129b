"""`jx gwas`-compatible command line for the exact-LMM path only (-lmm / -lmm2 / -fvlmm on -bfile / -vcf input).

Flag names, defaults and the output naming follow python/janusx/assoc/workflow.py:6599-7047
(`{out}/{prefix}.{trait}.{model}.tsv`, workflow_model_stream.py:992).  The scan itself is ONE call into the
B200 library per trait (jxrs.*_bed_to_tsv_f32), exactly where the reference makes its one Rust call
(workflow_model_stream.py:1449-1488).  Everything the reference CLI does outside this path (HMP/TXT input,
FarmCPU, plots, run history, -mem budgeting, LM switch) is out of scope; unsupported flags fail loudly.

  python -m janusx_b200.gwas -bfile panel -p pheno.tsv -n 0 -lmm -k 1 -q 3 -o out -prefix run1

The GRM (-k 1) is built on the device by the int8 tensor-core kernel (csrc/grm.cu) and decomposed
by the library's cuSOLVER entry point (csrc/eigh.cu, row N2); no torch is involved.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from typing import List, Optional

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
    o.add_argument("-t", "--thread", type=int, default=0, help="accepted for compatibility (CPU threads)")
    o.add_argument("-mem", "--mem", default=None, help="accepted for compatibility (host memory budget)")
    o.add_argument("-o", "--out", default=".")
    o.add_argument("-prefix", "--prefix", default=None)
    o.add_argument("-gpu", "--gpu", type=int, default=0, help="CUDA device index")
    args = ap.parse_args(argv)
    if not (args.lmm or args.lmm2 or args.fvlmm):
        ap.error("select at least one of -lmm, -lmm2, -fvlmm (other models are outside this build's scope)")
    if (args.bfile is None) == (args.vcf is None):
        ap.error("give exactly one of -bfile, -vcf")
    return args


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


def main(argv: Optional[List[str]] = None) -> int:
    args = parse_args(argv)
    from . import assoc, jxrs

    t0 = time.time()
    os.makedirs(args.out, exist_ok=True)
    if args.vcf is not None:
        args.bfile = _vcf_cache(args.vcf, args.out, args.snps_only)
    fam = _read_fam(args.bfile)
    ids_p, traits, Y = _read_table(args.pheno)
    cols = list(range(len(traits))) if not args.ncol else [int(c) for tok in args.ncol for c in str(tok).split(",")]
    prefix = args.prefix or os.path.basename(args.vcf or args.bfile).replace(".vcf.gz", "").replace(".vcf", "")
    outprefix = os.path.join(args.out, prefix)

    if args.grm == "1":
        K_full = _grm_from_bed(args.bfile, len(fam), args.gpu, args.maf, args.geno, args.het)
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
        fidx = np.array([pos_fam[s] for s in sample_ids], dtype=np.int64)
        y = y_all[ok]
        K = K_full[np.ix_(fidx, fidx)]
        X_parts = []
        if cov_mat is not None:
            X_parts.append(np.stack([cov_mat[pos_cov[s]] for s in sample_ids]))
        nq = int(args.qcov) if str(args.qcov).isdigit() else 0
        # null model (eigh on K + 1e-6 I, rotation, REML null fit: pyBLUP/assoc.py:1595-1876) -- device-backed
        base = assoc.LMM(y, None, K, device=args.gpu) if (nq == 0 and not X_parts) else None
        if base is None:
            if nq > 0:
                evals, evecs = assoc._eigh(K + 1e-6 * np.eye(len(y)), args.gpu)
                X_parts.append(evecs[:, ::-1][:, :nq] * np.sqrt(np.maximum(evals[::-1][:nq], 0.0)))
            base = assoc.LMM(y, np.concatenate(X_parts, axis=1), K, device=args.gpu)
        l10 = float(np.log10(base.lbd_null))
        print(f"[{trait}] n={len(y)} covariates={base.Xcov.shape[1]} lambda_null={base.lbd_null:.6g} "
              f"pve={base.pve:.4f} bounds=({base.bounds[0]:.3f},{base.bounds[1]:.3f})", file=sys.stderr)
        common = (base.S, base.Xcov, base.y[:, 0], base.Dh, args.maf, args.geno, args.het)
        kw = dict(genetic_model=args.model, snps_only=args.snps_only,
                  sample_ids=(None if sample_ids == fam else sample_ids))
        for model in models:
            out_tsv = f"{outprefix}.{trait}.{model}.tsv" if args.model == "add" else f"{outprefix}.{trait}.{args.model}.{model}.tsv"
            tmp = out_tsv + ".tmp"
            t1 = time.time()
            if model == "lmm":
                rows = jxrs.lmm_reml_assoc_bed_to_tsv_f32(args.bfile, tmp, *common, low=base.bounds[0],
                                                          high=base.bounds[1], max_iter=30, tol=1e-2,
                                                          init_log10_lbd=l10, **kw)
            elif model == "lmm2":
                rows = jxrs.lmm_reml_lmm2_assoc_bed_to_tsv_f32(args.bfile, tmp, *common, low=base.bounds[0],
                                                               high=base.bounds[1], max_iter=30, tol=1e-2,
                                                               init_log10_lbd_reml=l10, init_log10_lbd_ml=l10, **kw)
            else:
                rows, _, _ = jxrs.fvlmm_assoc_bed_to_tsv_f32(args.bfile, tmp, base.S, base.Xcov, base.y[:, 0], l10,
                                                            base.Dh, args.maf, args.geno, args.het, **kw)
            os.replace(tmp, out_tsv)   # atomic rename like workflow.py:833-846
            print(f"[{trait}] {model}: {rows} SNPs -> {out_tsv} ({time.time() - t1:.2f} s)", file=sys.stderr)
    print(f"done in {time.time() - t0:.2f} s", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
