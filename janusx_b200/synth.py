"""Synthetic GWAS inputs with the reference simulator's distributions.

Restates `jx sim` (python/janusx/script/sim.py:49-66 HWE draw, :133-176 chunk loop, :252-276 trait):
per-SNP maf ~ U(maf_low, maf_high); genotype from one uniform per call (0 if u<(1-maf)^2,
1 if u<(1-maf)^2+2maf(1-maf), else 2); sites chrom "1", pos i, alleles A/T; phenotype
100 + G beta + e at the requested pve.  Adds a missing-call knob (code 01) the reference
simulator lacks, to exercise imputation.  PLINK packing follows src/math/bedmath.rs:20-27.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

# dosage / missing -> PLINK 2-bit code (00 hom-ref, 10 het, 11 hom-alt, 01 missing)
_CODE_OF_DOSAGE = np.array([0b00, 0b10, 0b11, 0b01], dtype=np.uint8)


def pack_dosages(g: np.ndarray) -> np.ndarray:
    """g: int8[m, n] in {0,1,2} or 3/-1 for missing -> packed u8[m, ceil(n/4)] (SNP-major)."""
    g = np.asarray(g)
    m, n = g.shape
    gi = np.where((g < 0) | (g > 2), 3, g).astype(np.uint8)
    codes = _CODE_OF_DOSAGE[gi]
    pad = (-n) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros((m, pad), dtype=np.uint8)], axis=1)
    c4 = codes.reshape(m, -1, 4)
    return (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)


def unpack_codes(packed: np.ndarray, n: int) -> np.ndarray:
    """packed u8[m, bps] -> codes u8[m, n] (raw 2-bit values)."""
    packed = np.asarray(packed, dtype=np.uint8)
    sh = np.array([0, 2, 4, 6], dtype=np.uint8)
    c = (packed[:, :, None] >> sh[None, None, :]) & 3
    return c.reshape(packed.shape[0], -1)[:, :n]


def draw_genotypes(m: int, n: int, seed: int = 20260609, maf_low: float = 0.02, maf_high: float = 0.45,
                   missing_rate: float = 0.0, chunk: int = 4096):
    """-> (packed u8[m, ceil(n/4)], mafs f32[m]).  Unrelated individuals, HWE."""
    rng = np.random.default_rng(seed)
    bps = (n + 3) // 4
    packed = np.empty((m, bps), dtype=np.uint8)
    mafs = np.empty(m, dtype=np.float32)
    done = 0
    while done < m:
        k = min(chunk, m - done)
        mf = rng.uniform(maf_low, maf_high, size=k).astype(np.float32)
        u = rng.random((k, n), dtype=np.float32)
        p0 = (1.0 - mf) ** 2
        p1 = p0 + 2.0 * mf * (1.0 - mf)
        g = (u >= p0[:, None]).astype(np.int8) + (u >= p1[:, None]).astype(np.int8)
        if missing_rate > 0.0:
            miss = rng.random((k, n), dtype=np.float32) < missing_rate
            g[miss] = 3
        packed[done:done + k] = pack_dosages(g)
        mafs[done:done + k] = mf
        done += k
    return packed, mafs


def dosage_matrix(packed: np.ndarray, n: int) -> np.ndarray:
    """f64[m, n] dosages with NaN for missing (host helper for GRM / phenotype simulation)."""
    codes = unpack_codes(packed, n)
    lut = np.array([0.0, np.nan, 1.0, 2.0])
    return lut[codes]


def vanraden_grm(packed: np.ndarray, n: int) -> np.ndarray:
    """Centred VanRaden GRM K = ZZ^T / sum 2p(1-p) (src/stats/grm.rs:343-356); missing -> mean."""
    g = dosage_matrix(packed, n)
    mu = np.nanmean(g, axis=1)
    g = np.where(np.isnan(g), mu[:, None], g)
    p = mu / 2.0
    z = g - mu[:, None]
    denom = float(np.sum(2.0 * p * (1.0 - p)))
    return (z.T @ z) / max(denom, 1e-12)


@dataclass
class SynthCase:
    packed: np.ndarray          # u8[m, ceil(n/4)]
    n: int
    y: np.ndarray               # f64[n]
    cov: np.ndarray             # f64[n, q] (without intercept)
    s: np.ndarray               # f64[n] eigenvalues of K + 1e-6 I (ascending)
    u: np.ndarray               # f64[n, n] eigenvectors in columns


def make_case(n: int, m: int, q: int = 3, seed: int = 20260609, pve: float = 0.5, missing_rate: float = 0.0,
              grm_snps: Optional[int] = None) -> SynthCase:
    packed, _ = draw_genotypes(m, n, seed=seed, missing_rate=missing_rate)
    k_rows = packed if grm_snps is None else packed[: min(m, grm_snps)]
    K = vanraden_grm(k_rows, n)
    K[np.diag_indices(n)] += 1e-6  # workflow_model_stream.py:902
    s, u = np.linalg.eigh(K)
    rng_t = np.random.default_rng(seed + 1)
    g = dosage_matrix(k_rows, n)
    mu = np.nanmean(g, axis=1)
    g = np.where(np.isnan(g), mu[:, None], g)
    beta = rng_t.normal(0.0, 1.0, size=g.shape[0])
    gv = beta @ (g - mu[:, None])
    vg = float(np.var(gv))
    ve = vg * (1.0 - pve) / max(pve, 1e-12) if vg > 0 else 1.0
    y = 100.0 + gv + rng_t.normal(0.0, np.sqrt(ve), size=n)
    rng_c = np.random.default_rng(seed + 2)
    cov = rng_c.normal(0.0, 1.0, size=(n, q))
    return SynthCase(packed=packed, n=n, y=y, cov=cov, s=s, u=u)


def write_plink(prefix: str, packed: np.ndarray, n: int, snp_ids=None, chrom: str = "1") -> None:
    """Write prefix.bed/.bim/.fam (SNP-major BED, magic 6C 1B 01)."""
    m = packed.shape[0]
    with open(f"{prefix}.bed", "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        fh.write(np.ascontiguousarray(packed, dtype=np.uint8).tobytes())
    with open(f"{prefix}.bim", "w") as fh:
        for i in range(m):
            sid = f"snp{i}" if snp_ids is None else snp_ids[i]
            fh.write(f"{chrom}\t{sid}\t0\t{i}\tA\tT\n")
    with open(f"{prefix}.fam", "w") as fh:
        for j in range(n):
            fh.write(f"F{j}\tS{j}\t0\t0\t0\t-9\n")
