"""BedChunkReader: the prepared-chunk source of the reference's route "B" (several models on one trait share decoded
chunks: SURVEY 3.2), device-backed.

Mirrors `BedChunkReader` of src/io/gfreader.rs:3138-3760 for the exact-LMM path: constructor arguments, `n_samples`,
`n_snps`, `sample_ids` and `next_chunk_prepared(chunk_size, coding, snps_only) -> (f32[m, n] centred, sites, af, miss)`.
The per-row QC of this route is NOT the unified scan's f32 arithmetic: it runs in f64 on the genotype counts
(process_snp_row_with_precomputed_counts_impl, src/io/gfcore.rs:405-480).  The counts are exact integers, so they come
from the device (jxb_decode_packed), the handful of f64 expressions per row is evaluated here exactly as the reference
writes them, and the decode / impute / centre of the kept rows runs on the device (jxb_decode_packed_prepared).
Row selection (snp_range / snp_indices / bim_range / snp_sites, then the chr_keys / bp_min / bp_max / ranges site filter)
is host logic and follows src/io/gfreader.rs:125-215, 583-726, 3256-3281 (`select_snp_rows`).  `mmap_window_mb` is
accepted with the reference's argument check; the payload is a demand-paged numpy memmap either way.
Non-additive codings of next_chunk_prepared (dom / rec / het, src/io/gfreader.rs:3161-3186): four values per row computed
here exactly as the reference writes them, per-row LUT decode on the device; fill_missing = false (missing calls keep the
decoder's -9 marker, src/io/gfcore.rs:468-476) goes the same way.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._cabi import check, lib, ptr, require_gpu
from .jxrs import DeviceModel, SiteInfo


class ChunkSite(SiteInfo):
    """chrom / pos / snp / ref_allele / alt_allele (the Py-exposed SiteInfo, src/io/gfreader.rs:33-47)."""
    __slots__ = ("snp",)

    def __init__(self, chrom, pos, snp, ref_allele, alt_allele):
        super().__init__(chrom, pos, ref_allele, alt_allele)
        self.snp = str(snp)


def normalize_chr_key(chrom: str) -> str:
    """Chromosome key of the site filter (src/io/gfreader.rs:227-234): trimmed, a leading "chr" (any case) dropped, upper case."""
    c = str(chrom).strip()
    if c[:3].lower() == "chr":
        c = c[3:]
    return c.strip().upper()


def select_snp_rows(chroms: Sequence[str], positions: Sequence[int], snp_range=None, snp_indices=None, bim_range=None,
                    snp_sites=None, chr_keys=None, bp_min=None, bp_max=None, ranges=None) -> Optional[np.ndarray]:
    """Source rows a BedChunkReader walks, in order; None = every row of the BIM.

    First one of the four primary selectors (src/io/gfreader.rs:125-215: at most one; snp_range half open; snp_indices as
    given, in range, no duplicates; bim_range by exact chromosome string and closed position interval; snp_sites in the order
    of the keys, every key must exist, a key listed twice in the BIM contributes all its rows), then the site filter on what is
    left (src/io/gfreader.rs:583-726, 3259-3281: normalised chromosome keys, every given condition must hold, `ranges` is
    a union).  Errors carry the reference's messages (RuntimeError there as here)."""
    m = len(chroms)
    if sum(v is not None for v in (snp_range, snp_indices, bim_range, snp_sites)) > 1:
        raise RuntimeError("Provide only one of snp_range, snp_indices, bim_range, or snp_sites")
    pos = np.asarray(positions, dtype=np.int64)
    rows: Optional[np.ndarray] = None
    if snp_range is not None:
        start, end = int(snp_range[0]), int(snp_range[1])
        if start >= end or end > m:
            raise RuntimeError(f"invalid snp_range: ({start}, {end})")
        rows = np.arange(start, end, dtype=np.int64)
    elif snp_indices is not None:
        rows = np.asarray(list(snp_indices), dtype=np.int64)
        if rows.size == 0:
            raise RuntimeError("snp_indices is empty")
        seen = set()
        for i in rows.tolist():
            if i < 0 or i >= m:
                raise RuntimeError(f"snp index out of range: {i}")
            if i in seen:
                raise RuntimeError(f"duplicate snp index: {i}")
            seen.add(i)
    elif bim_range is not None:
        chrom, start, end = str(bim_range[0]), int(bim_range[1]), int(bim_range[2])
        if start > end:
            raise RuntimeError("bim_range start > end")
        same = np.fromiter((c == chrom for c in chroms), dtype=bool, count=m)
        rows = np.nonzero(same & (pos >= start) & (pos <= end))[0].astype(np.int64)
    elif snp_sites is not None:
        keys = [(str(c), int(p)) for c, p in snp_sites]
        if not keys:
            raise RuntimeError("snp_sites is empty")
        where = {}
        for i, (c, p) in enumerate(zip(chroms, pos.tolist())):
            where.setdefault((c, p), []).append(i)
        picked: List[int] = []
        for c, p in keys:
            if (c, p) not in where:
                raise RuntimeError(f"snp site not found: ({c}, {p})")
            picked.extend(where[(c, p)])
        if not picked:
            raise RuntimeError("no SNPs matched from snp_sites")
        rows = np.asarray(picked, dtype=np.int64)

    # site filter
    if bp_min is not None and bp_max is not None and int(bp_min) > int(bp_max):
        raise RuntimeError("bp_min cannot be greater than bp_max")
    chr_set = None
    if chr_keys is not None:
        chr_set = {k for k in (normalize_chr_key(c) for c in chr_keys) if k}
        chr_set = chr_set or None
    rng = None
    if ranges is not None and len(ranges) > 0:
        rng = []
        for c, a, b in ranges:
            if int(a) > int(b):
                raise RuntimeError("One range has start > end")
            rng.append((normalize_chr_key(c), int(a), int(b)))
    if chr_set is None and bp_min is None and bp_max is None and rng is None:
        return rows
    cand = np.arange(m, dtype=np.int64) if rows is None else rows
    norm = {}
    keep = np.ones(cand.shape[0], dtype=bool)
    for j, i in enumerate(cand.tolist()):
        c = chroms[i]
        k = norm.get(c)
        if k is None:
            k = norm[c] = normalize_chr_key(c)
        p = int(pos[i])
        ok = True
        if chr_set is not None and k not in chr_set:
            ok = False
        if ok and bp_min is not None and p < int(bp_min):
            ok = False
        if ok and bp_max is not None and p > int(bp_max):
            ok = False
        if ok and rng is not None and not any(k == rc and rs <= p <= re for rc, rs, re in rng):
            ok = False
        keep[j] = ok
    return cand[keep]


def _simple_allele(a: str) -> bool:
    a = a.strip()
    return len(a) == 1 and a.upper() in "ACGT"


def prepared_row_decisions(missing, het, hom_alt, n: int, maf_thr: float, miss_thr: float, het_thr: float):
    """The f64 QC of route B on integer genotype counts (src/io/gfcore.rs:405-480, preserve_alt_orientation = true,
    fill_missing = true) -> (keep bool[m], imputed f32[m]).  Thresholds are f32 like the reference's fields."""
    missing = np.asarray(missing, dtype=np.int64)
    het = np.asarray(het, dtype=np.int64)
    hom = np.asarray(hom_alt, dtype=np.int64)
    maf_t, miss_t, het_t = np.float32(maf_thr), np.float32(miss_thr), np.float32(het_thr)
    non_missing = n - missing
    alt_sum = (het + 2 * hom).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        missing_rate = (1.0 - non_missing.astype(np.float64) / float(n)).astype(np.float32)
        keep = ~(missing_rate > miss_t)
        empty = non_missing == 0
        keep &= ~(empty & (maf_t > 0))
        if het_t < np.float32(1.0):                                   # apply_het_filter = het < 1.0
            het_rate = het.astype(np.float64) / non_missing.astype(np.float64)
            keep &= ~(~empty & (het_rate > np.float64(het_t)))
        alt_freq = alt_sum / (2.0 * non_missing.astype(np.float64))
        maf = np.minimum(alt_freq, 1.0 - alt_freq).astype(np.float32)
        keep &= ~(~empty & (maf < maf_t))
        imputed = (alt_sum / non_missing.astype(np.float64)).astype(np.float32)
    imputed = np.where(empty, np.float32(0.0), imputed).astype(np.float32)   # all-missing rows are filled with 0
    return keep, imputed


def raw_row_decisions(missing, het, hom_alt, n: int, maf_thr: float, miss_thr: float, het_thr: float):
    """process_snp_row (src/io/gfcore.rs:405-480 with preserve_alt_orientation = false, fill_missing = true) on integer
    counts -> (keep bool[m], flip bool[m], lut f32[m, 4]): rows whose ALT frequency exceeds 0.5 are recoded to the other
    allele (g -> 2 - g, alleles swapped by the caller); missing calls take the mean dosage of the (recoded) row."""
    missing = np.asarray(missing, dtype=np.int64)
    het = np.asarray(het, dtype=np.int64)
    hom = np.asarray(hom_alt, dtype=np.int64)
    maf_t, miss_t, het_t = np.float32(maf_thr), np.float32(miss_thr), np.float32(het_thr)
    non_missing = n - missing
    raw_alt = (het + 2 * hom).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        missing_rate = (1.0 - non_missing.astype(np.float64) / float(n)).astype(np.float32)
        keep = ~(missing_rate > miss_t)
        empty = non_missing == 0
        keep &= ~(empty & (maf_t > 0))
        if het_t < np.float32(1.0):
            keep &= ~(~empty & ((het.astype(np.float64) / non_missing.astype(np.float64)) > np.float64(het_t)))
        alt_freq = raw_alt / (2.0 * non_missing.astype(np.float64))
        flip = ~empty & (alt_freq > 0.5)
        alt_sum = np.where(flip, 2.0 * non_missing.astype(np.float64) - raw_alt, raw_alt)
        maf = np.minimum(alt_freq, 1.0 - alt_freq).astype(np.float32)
        keep &= ~(~empty & (maf < maf_t))
        imputed = np.where(empty, 0.0, alt_sum / non_missing.astype(np.float64)).astype(np.float32)
    lut = np.empty((missing.shape[0], 4), dtype=np.float32)
    lut[:, 0] = np.where(flip, np.float32(2.0), np.float32(0.0))
    lut[:, 1] = imputed
    lut[:, 2] = np.float32(1.0)
    lut[:, 3] = np.where(flip, np.float32(0.0), np.float32(2.0))
    return keep, flip, lut


def coded_row_lut(missing, het, hom_alt, imputed, n: int, coding: str):
    """Codings of next_chunk_prepared (src/io/gfreader.rs:3161-3186, 3632-3653) for rows whose missing calls hold
    `imputed` (the row mean, or the raw marker -9 when fill_missing = false): value map with the reference's 1e-6 tolerance in
    f32 ("add" = identity), exact sum of the coded row, mean as f32(sum / n), centring in f32 -> (lut f32[m, 4] indexed
    by the PLINK 2-bit code {hom-ref, missing, het, hom-alt}, coded_mean f32[m]).  The additive coding of filled rows does not
    come through here (its sum involves the imputed value; jxb_decode_packed_prepared handles it)."""
    missing = np.asarray(missing, dtype=np.int64)
    het = np.asarray(het, dtype=np.int64)
    hom = np.asarray(hom_alt, dtype=np.int64)
    imputed = np.asarray(imputed, dtype=np.float32)
    m = missing.shape[0]
    vals = np.empty((m, 4), dtype=np.float32)
    vals[:, 0] = np.float32(0.0); vals[:, 1] = imputed; vals[:, 2] = np.float32(1.0); vals[:, 3] = np.float32(2.0)
    tol = np.float32(1e-6)
    near1 = np.abs(vals - np.float32(1.0)) <= tol
    near2 = np.abs(vals - np.float32(2.0)) <= tol
    if coding == "add":
        coded = vals                                                      # integers (or the -9 marker): the sum stays exact
    else:
        hit = {"dom": near1 | near2, "rec": near2, "het": near1}[coding]
        coded = np.where(hit, np.float32(1.0), np.float32(0.0)).astype(np.float32)
    cnt = np.stack([n - missing - het - hom, missing, het, hom], axis=1).astype(np.float64)
    total = (cnt * coded.astype(np.float64)).sum(axis=1)                 # 0/1 values times counts: an exact integer
    coded_mean = (total / float(n)).astype(np.float32)
    return np.ascontiguousarray(coded - coded_mean[:, None], dtype=np.float32), coded_mean


class BedChunkReader:
    def __init__(self, prefix, maf_threshold=None, max_missing_rate=None, fill_missing=None, snp_range=None,
                 snp_indices=None, bim_range=None, snp_sites=None, sample_ids=None, sample_indices=None,
                 mmap_window_mb=None, chr_keys=None, bp_min=None, bp_max=None, ranges=None, model=None,
                 het_threshold=None, device: int = 0):
        require_gpu()
        self.maf = 0.0 if maf_threshold is None else float(maf_threshold)
        self.miss = 1.0 if max_missing_rate is None else float(max_missing_rate)
        self._fill = True if fill_missing is None else bool(fill_missing)   # false: missing calls keep the decoder's -9
        model_key = (model or "add").lower()
        if model_key not in ("add", "dom", "rec", "het"):
            raise ValueError("model must be one of: add, dom, rec, het")
        self.het = 1.0 if het_threshold is None else float(het_threshold)
        if not (0.0 <= self.het <= 1.0):
            raise ValueError("het_threshold must be within [0, 1.0]")
        if mmap_window_mb is not None and (snp_range is not None or snp_indices is not None or bim_range is not None):
            raise ValueError("mmap_window_mb does not support snp_range/snp_indices/bim_range")   # gfreader.rs:3246-3252
        self._mmap_window_mb = mmap_window_mb
        self.prefix = str(prefix)
        with open(self.prefix + ".fam") as fh:
            fam = [line.split()[1] for line in fh if line.strip()]
        if not fam:
            raise RuntimeError("no samples in PLINK FAM")
        self._n_full = len(fam)
        # build_sample_selection, src/io/gfreader.rs:75-123
        if sample_ids is not None and sample_indices is not None:
            raise RuntimeError("Provide only one of sample_ids or sample_indices")
        if sample_ids is not None:
            ids = [str(s) for s in sample_ids]
            if not ids:
                raise RuntimeError("sample_ids is empty")
            pos = {sid: i for i, sid in enumerate(fam)}
            idx, seen = [], set()
            for sid in ids:
                if sid not in pos:
                    raise RuntimeError(f"sample id not found: {sid}")
                if pos[sid] in seen:
                    raise RuntimeError(f"duplicate sample id: {sid}")
                seen.add(pos[sid])
                idx.append(pos[sid])
        elif sample_indices is not None:
            idx = [int(i) for i in sample_indices]
            if not idx:
                raise RuntimeError("sample_indices is empty")
            seen = set()
            for i in idx:
                if i < 0 or i >= len(fam):
                    raise RuntimeError(f"sample index out of range: {i}")
                if i in seen:
                    raise RuntimeError(f"duplicate sample index: {i}")
                seen.add(i)
            ids = [fam[i] for i in idx]
        else:
            idx, ids = list(range(len(fam))), list(fam)
        self._sidx = np.asarray(idx, dtype=np.int64)
        self._ids = ids
        self._identity = len(idx) == len(fam) and idx == list(range(len(fam)))
        # sites + packed payload
        self._sites: List[ChunkSite] = []
        with open(self.prefix + ".bim") as fh:
            for ln, line in enumerate(fh, 1):
                tok = line.split()
                if len(tok) < 6:
                    raise RuntimeError(f"Malformed BIM line at {self.prefix}.bim:{ln}: {line.rstrip()}")
                try:
                    p = int(tok[3])
                    p = p if -(1 << 31) <= p < (1 << 31) else 0
                except ValueError:
                    p = 0
                self._sites.append(ChunkSite(tok[0], p, tok[1], tok[4], tok[5]))
        bps = (self._n_full + 3) // 4
        raw = np.memmap(self.prefix + ".bed", dtype=np.uint8, mode="r")
        if raw.shape[0] < 3 or bytes(raw[:3]) != b"\x6c\x1b\x01":
            raise RuntimeError("only SNP-major BED supported")
        if (raw.shape[0] - 3) % bps != 0 or (raw.shape[0] - 3) // bps != len(self._sites):
            raise RuntimeError("BED payload does not match the FAM/BIM dimensions")
        self._packed = raw[3:].reshape(len(self._sites), bps)
        # build_snp_indices + the site filter, src/io/gfreader.rs:125-215, 3256-3281
        self._snp_indices: Optional[np.ndarray] = select_snp_rows(
            [st.chrom for st in self._sites], [st.pos for st in self._sites], snp_range=snp_range, snp_indices=snp_indices,
            bim_range=bim_range, snp_sites=snp_sites, chr_keys=chr_keys, bp_min=bp_min, bp_max=bp_max, ranges=ranges)
        self._cursor = 0
        n = len(idx)
        self._dev = DeviceModel(np.ones(n), np.ones((n, 1)), np.zeros(n), device=device)   # decode workspace only

    @property
    def n_samples(self) -> int:
        return len(self._ids)

    @property
    def n_snps(self) -> int:
        return len(self._sites) if self._snp_indices is None else int(self._snp_indices.shape[0])

    @property
    def sample_ids(self) -> List[str]:
        return list(self._ids)

    def _source_rows(self, count: int) -> np.ndarray:
        if self._snp_indices is None:
            rows = np.arange(self._cursor, min(self._cursor + count, len(self._sites)), dtype=np.int64)
        else:
            rows = self._snp_indices[self._cursor:self._cursor + count]
        self._cursor += int(rows.shape[0])
        return rows

    def _decode_lut(self, packed: np.ndarray, lut: np.ndarray) -> np.ndarray:
        g = np.empty((packed.shape[0], self.n_samples), dtype=np.float32)
        lut = np.ascontiguousarray(lut, dtype=np.float32)
        check(lib().jxb_decode_packed_lut(self._dev.handle, ptr(packed), packed.shape[1], packed.shape[0], self._n_full,
                                          None if self._identity else ptr(self._sidx), ptr(lut), ptr(g)))
        return g

    def next_chunk(self, chunk_size: int):
        """src/io/gfreader.rs:3319-3440 -> None at the end, else (dosage f32[m, n], sites): rows that pass the reader's
        thresholds, recoded to the minor allele where the ALT frequency exceeds 0.5 (alleles swapped in the returned site),
        missing calls filled with the row mean.  Counts and decode on the device, the f64 row decisions on the host."""
        if chunk_size == 0:
            raise ValueError("chunk_size must be > 0")
        n = self.n_samples
        if n == 0:
            return None
        sidx = None if self._identity else self._sidx
        blocks, sites, m = [], [], 0
        while m < chunk_size and self._cursor < self.n_snps:
            rows = self._source_rows(chunk_size - m)
            packed = np.ascontiguousarray(self._packed[rows])
            counts, _, _, _ = self._dev.decode_packed(packed, self._n_full, sidx, 0.0, 1.0, 0.0, want_g=False)
            keep, flip, lut = raw_row_decisions(counts[:, 0], counts[:, 1], counts[:, 2], n, self.maf, self.miss, self.het)
            if not self._fill:
                lut[:, 1] = np.float32(-9.0)                              # gfcore.rs:468-476 skipped; 2 - g spares g < 0
            k = np.nonzero(keep)[0]
            if k.size == 0:
                continue
            blocks.append(self._decode_lut(np.ascontiguousarray(packed[k]), lut[k]))
            for i in k:
                st = self._sites[int(rows[i])]
                sites.append(ChunkSite(st.chrom, st.pos, st.snp, st.alt_allele, st.ref_allele) if flip[i] else st)
            m += int(k.size)
        if m == 0:
            return None
        return np.concatenate(blocks), sites

    def next_chunk_prepared(self, chunk_size: int, coding: Optional[str] = None, snps_only: bool = False):
        """-> None at the end, else (geno_centered f32[m, n], sites, af f32[m], miss f32[m]); m <= chunk_size rows that
        passed the QC, in BED order (the reference keeps reading until the chunk is full)."""
        if chunk_size == 0:
            raise ValueError("chunk_size must be > 0")
        coding_key = (coding or "add").strip().lower()
        if coding_key not in ("add", "dom", "rec", "het"):
            raise ValueError("coding must be one of: add, dom, rec, het")
        if self._mmap_window_mb is not None and self._snp_indices is not None:   # gfreader.rs:3687-3692
            raise RuntimeError("windowed mmap mode does not support explicit snp index selection")
        n = self.n_samples
        sidx = None if self._identity else self._sidx
        blocks, sites, afs, misses, m = [], [], [], [], 0
        while m < chunk_size and self._cursor < self.n_snps:
            rows = self._source_rows(chunk_size - m)
            packed = np.ascontiguousarray(self._packed[rows])
            counts, _, _, _ = self._dev.decode_packed(packed, self._n_full, sidx, 0.0, 1.0, 0.0, want_g=False)
            keep, imputed = prepared_row_decisions(counts[:, 0], counts[:, 1], counts[:, 2], n, self.maf, self.miss, self.het)
            if snps_only:
                keep &= np.array([_simple_allele(self._sites[int(r)].ref_allele) and _simple_allele(self._sites[int(r)].alt_allele)
                                  for r in rows], dtype=bool)
            nk = int(keep.sum())
            if nk == 0:
                continue
            if coding_key != "add" or not self._fill:
                # dom / rec / het, or unfilled rows: the row takes four values {0, imputed | -9, 1, 2}; the coding map, the exact
                # f64 sum of the coded row and the f32 centring give four output values per row -> per-row LUT decode on the device
                k = np.nonzero(keep)[0]
                fillv = imputed[k] if self._fill else np.full(k.shape[0], -9.0, dtype=np.float32)
                lut, coded_mean = coded_row_lut(counts[k, 0], counts[k, 1], counts[k, 2], fillv, n, coding_key)
                blocks.append(self._decode_lut(np.ascontiguousarray(packed[k]), lut))
                sites.extend(self._sites[int(rows[i])] for i in k)
                afs.append(coded_mean * np.float32(0.5) if coding_key == "add" else coded_mean)
                misses.append(counts[k, 0].astype(np.float32))
                m += nk
                continue
            g = np.empty((nk, n), dtype=np.float32)
            got = C.c_size_t()
            keep_u8 = np.ascontiguousarray(keep, dtype=np.uint8)
            af_half = np.ascontiguousarray(imputed * np.float32(0.5), dtype=np.float32)   # f32(2 * af) == imputed, exactly
            check(lib().jxb_decode_packed_prepared(self._dev.handle, ptr(packed), packed.shape[1], packed.shape[0],
                                                   self._n_full, ptr(sidx), ptr(keep_u8), ptr(af_half), 0, ptr(g),
                                                   C.byref(got)))
            if got.value != nk:
                raise RuntimeError("internal error: device kept-row count differs from the host decisions")
            k = np.nonzero(keep)[0]
            miss_cnt = counts[k, 0].astype(np.int64)
            # coded_mean = (sum of the filled row in f64 / n) as f32; the sum is exact (gfreader.rs:3667-3680)
            total = (counts[k, 1] + 2 * counts[k, 2]).astype(np.float64) + miss_cnt.astype(np.float64) * imputed[k].astype(np.float64)
            coded_mean = (total / float(n)).astype(np.float32)
            blocks.append(g)
            sites.extend(self._sites[int(rows[i])] for i in k)
            afs.append(coded_mean * np.float32(0.5))
            misses.append(miss_cnt.astype(np.float32))
            m += nk
        if m == 0:
            return None
        return np.concatenate(blocks), sites, np.concatenate(afs), np.concatenate(misses)


# ------------------------------------------------------------------------------------------------------
# prepared row statistics of one trait's samples (second half)
# ------------------------------------------------------------------------------------------------------
def _meta_row_decisions(missing, het, hom_alt, n: int, maf_thr, miss_thr, het_thr):
    """prepare_bed_logic_meta_owned... row closure (src/io/gfreader.rs:5378-5424) on integer counts:
    missing_rate f32 = missing / n (packed_row_stats_from_counts, :1911-1929); alt_freq f32 = alt_sum / (2 f32 * non_missing);
    het rate compared in f64; maf test min(af, 1 - af) >= maf_thr in f32.  -> (keep, missing_rate f32, alt_freq f32)."""
    missing = np.asarray(missing, dtype=np.int64)
    het = np.asarray(het, dtype=np.int64)
    hom = np.asarray(hom_alt, dtype=np.int64)
    maf_t, miss_t, het_t = np.float32(maf_thr), np.float32(miss_thr), np.float32(het_thr)
    non_missing = np.maximum(n - missing, 0)
    alt_sum = het + 2 * hom
    with np.errstate(divide="ignore", invalid="ignore"):
        miss_rate = ((n - non_missing).astype(np.float32) / np.float32(n)) if n > 0 else np.zeros(missing.shape, np.float32)
        alt_freq = alt_sum.astype(np.float32) / (np.float32(2.0) * non_missing.astype(np.float32))
        alt_freq = np.where(non_missing > 0, alt_freq, np.float32(0.0)).astype(np.float32)
        empty = non_missing == 0
        keep = ~(miss_rate > miss_t)
        keep &= ~(empty & ~(maf_t <= np.float32(0.0)))
        if het_t > np.float32(0.0):                                   # apply_het_filter = het_threshold > 0
            het_rate = het.astype(np.float64) / non_missing.astype(np.float64)
            keep &= ~(~empty & (het_rate > np.float64(het_t)))
        keep &= empty | (np.minimum(alt_freq, np.float32(1.0) - alt_freq) >= maf_t)
    return keep, miss_rate.astype(np.float32), alt_freq


def _bed_counts_selected(prefix: str, sample_indices, device: int, batch_rows: int = 1 << 16):
    """Integer genotype counts (missing, het, hom_alt) of every BED row over the selected samples, on the device
    (count_packed_row_counts[_selected_with_excluded], src/io/gfreader.rs:1378-1528)."""
    import os
    prefix = str(prefix)
    for ext in (".bed", ".bim", ".fam"):
        if prefix.lower().endswith(ext):
            prefix = prefix[: -len(ext)]
    with open(prefix + ".fam") as fh:
        n_full = sum(1 for line in fh if line.strip())
    if n_full == 0:
        raise RuntimeError("no samples found in PLINK input")
    sidx = None
    if sample_indices is not None:
        sidx = np.ascontiguousarray(np.asarray(sample_indices, dtype=np.int64).reshape(-1))
        bad = (sidx < 0) | (sidx >= n_full)
        if bad.any():
            raise ValueError(f"sample index out of range: {int(sidx[bad][0])} for n_samples={n_full}")
    n = n_full if sidx is None else int(sidx.shape[0])
    bps = (n_full + 3) // 4
    size = os.path.getsize(prefix + ".bed")
    raw = np.memmap(prefix + ".bed", dtype=np.uint8, mode="r")
    if size < 3 or bytes(raw[:3]) != b"\x6c\x1b\x01":
        raise RuntimeError("only SNP-major BED supported")
    if (size - 3) % bps:
        raise RuntimeError(f"BED payload length {size - 3} not a multiple of {bps}")
    m = (size - 3) // bps
    packed = raw[3:].reshape(m, bps)
    counts = np.zeros((m, 3), dtype=np.int64)
    mdl = DeviceModel(np.ones(max(n, 1)), np.ones((max(n, 1), 1)), np.zeros(max(n, 1)), None, device=device)
    try:
        for r0 in range(0, m, batch_rows):
            c, _, _, _ = mdl.decode_packed(np.ascontiguousarray(packed[r0:r0 + batch_rows]), n_full, sample_idx=sidx,
                                           maf_thr=0.0, miss_thr=1.0, het_thr=0.0, want_g=False)
            counts[r0:r0 + c.shape[0]] = c[:, :3]
    finally:
        mdl.close()
    return prefix, n_full, n, m, counts


def prepare_bed_logic_meta_selected(prefix, sample_indices=None, maf_threshold=0.0, max_missing_rate=1.0, het_threshold=1.0,
                                    snps_only=False, mmap_window_mb=None, threads=1, device: int = 0):
    """src/io/gfreader.rs:7119-7232 -> (row_idx i64[k], miss f32[k], af f32[k], row_flip bool[k], site_keep bool[m],
    n_samples, n_snps_total): the per-trait prepared row statistics the reference computes once per trait and hands to
    lmm_reml_assoc_bed_to_tsv_f32 as row_indices / row_missing / row_maf / row_flip (assoc/workflow.py:8870-8888).
    `af` is the ALT frequency over the selected samples (never folded); row_flip is all False (gfreader.rs:5468-5471)."""
    if not (0.0 <= maf_threshold <= 0.5):
        raise ValueError("maf_threshold must be within [0, 0.5]")
    if not (0.0 <= max_missing_rate <= 1.0):
        raise ValueError("max_missing_rate must be within [0, 1.0]")
    if not (0.0 <= het_threshold <= 1.0):
        raise ValueError("het_threshold must be within [0, 1.0]")
    require_gpu()
    prefix, n_full, n, m, counts = _bed_counts_selected(prefix, sample_indices, device)
    keep, miss_rate, alt_freq = _meta_row_decisions(counts[:, 0], counts[:, 1], counts[:, 2], n, maf_threshold,
                                                    max_missing_rate, het_threshold)
    if snps_only:
        with open(prefix + ".bim") as fh:
            simple = np.fromiter((_simple_allele(t[4]) and _simple_allele(t[5]) for t in (ln.split() for ln in fh)),
                                 dtype=bool, count=m)
        keep &= simple
    if not keep.any():
        raise RuntimeError("No SNPs left after packed BED filtering. Please relax thresholds.")
    row_idx = np.nonzero(keep)[0].astype(np.int64)
    return (row_idx, np.ascontiguousarray(miss_rate[keep]), np.ascontiguousarray(alt_freq[keep]),
            np.zeros(row_idx.shape[0], dtype=bool), keep, n_full, m)


def prepare_bed_logic_keep_mask(prefix, sample_indices=None, maf_threshold=0.0, max_missing_rate=1.0, het_threshold=1.0,
                                snps_only=False, mmap_window_mb=None, threads=1, device: int = 0):
    """src/io/gfreader.rs:7234-7300 -> (site_keep bool[m], n_samples, n_snps_total)."""
    out = prepare_bed_logic_meta_selected(prefix, sample_indices, maf_threshold, max_missing_rate, het_threshold, snps_only,
                                          mmap_window_mb, threads, device)
    return out[4], out[5], out[6]


class BedChunkReaderFromMeta(BedChunkReader):
    """src/io/gfreader.rs:7440-7730: the chunk source of the "meta-shared" route -- several models scan one trait with
    ONE prepared row list (prepare_bed_logic_meta_selected) instead of re-deriving QC per model
    (workflow_model_stream.py:400-415).  Rows are decoded with the shared metadata: mean = 2 * ALT frequency
    (row_maf, flipped to 1 - maf where row_flip), missing calls -> 0, no re-centring; af = mean / 2 and the caller's
    row_missing are returned per row.  Additive coding only, like the reference."""

    def __init__(self, prefix, row_indices, row_flip, row_missing, row_maf, sample_ids=None, sample_indices=None,
                 mmap_window_mb=None, device: int = 0):
        super().__init__(prefix, sample_ids=sample_ids, sample_indices=sample_indices, device=device)
        idx = np.ascontiguousarray(np.asarray(row_indices, dtype=np.int64).reshape(-1))
        flip = np.asarray(row_flip, dtype=bool).reshape(-1)
        miss = np.ascontiguousarray(np.asarray(row_missing, dtype=np.float32).reshape(-1))
        maf = np.asarray(row_maf, dtype=np.float32).reshape(-1)
        m = idx.shape[0]
        if flip.shape[0] != m or miss.shape[0] != m or maf.shape[0] != m:
            raise ValueError(f"metadata length mismatch: row_indices={m}, row_flip={flip.shape[0]}, "
                             f"row_missing={miss.shape[0]}, row_maf={maf.shape[0]}")
        n_total = len(self._sites)
        if m and (idx.min() < 0 or idx.max() >= n_total):
            bad = int(idx[(idx < 0) | (idx >= n_total)][0])
            raise ValueError(f"row index out of range: {bad} for n_snps={n_total}")
        if m > 1 and np.any(np.diff(idx) < 0):
            raise ValueError("row_indices must be sorted in ascending BED order")
        self._rows = idx
        self._row_missing = miss
        # gfreader.rs:7541-7550: alt mean = 2 * clamp(flip ? 1 - clamp(maf) : clamp(maf)) in f32
        mc = np.clip(maf, np.float32(0.0), np.float32(1.0)).astype(np.float32)
        alt = np.where(flip, np.float32(1.0) - mc, mc).astype(np.float32)
        self._row_alt_mean = (np.float32(2.0) * np.clip(alt, np.float32(0.0), np.float32(1.0))).astype(np.float32)
        self._pos = 0

    @property
    def n_snps(self) -> int:
        return int(self._rows.shape[0])

    def next_chunk(self, *_a, **_k):
        raise AttributeError("BedChunkReaderFromMeta has no next_chunk (the reference class only serves next_chunk_prepared)")

    def next_chunk_prepared(self, chunk_size, coding=None, snps_only=False):
        if int(chunk_size) <= 0:
            raise ValueError("chunk_size must be > 0")
        if (coding or "add").strip().lower() != "add":
            raise ValueError("BedChunkReaderFromMeta currently supports additive coding only")
        if self.n_samples == 0:
            return None
        meta_idx = []
        while len(meta_idx) < int(chunk_size) and self._pos < self._rows.shape[0]:
            k = self._pos
            self._pos += 1
            site = self._sites[int(self._rows[k])]
            if snps_only and not (_simple_allele(site.ref_allele) and _simple_allele(site.alt_allele)):
                continue
            meta_idx.append(k)
        if not meta_idx:
            return None
        meta_idx = np.asarray(meta_idx, dtype=np.int64)
        src = self._rows[meta_idx]
        packed = np.ascontiguousarray(self._packed[src])
        mean = np.ascontiguousarray(self._row_alt_mean[meta_idx])
        # bedmath.rs:1161-1230 with inv_sd = 1, flip = false: [(0 - mean), 0 (missing), (1 - mean), (2 - mean)] in f32
        lut = np.stack([np.float32(0.0) - mean, np.zeros_like(mean), np.float32(1.0) - mean, np.float32(2.0) - mean], axis=1)
        g = self._decode_lut(packed, lut)
        sites = [self._sites[int(j)] for j in src]
        return g, sites, (mean * np.float32(0.5)).astype(np.float32), np.ascontiguousarray(self._row_missing[meta_idx])
