"""ctypes front-end of the CPU oracle (oracle/jx_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package (janusx_b200/) must never import it.

PARITY UNPINNED: the reference (Rust cdylib) cannot be built here and holds no expected
beta/se/p/lambda values for this path; see oracle/jx_oracle.h and DESIGN.md.

Function names mirror the reference's PyO3 surface (src/lib.rs:911-941) so that parity tests
read `oracle.lmm_reml_chunk_from_snp_f32(...)` next to `jxrs.lmm_reml_chunk_from_snp_f32(...)`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libjx_oracle.so"
_lib = None

MODEL_CODES = {"add": 0, "dom": 1, "rec": 2, "het": 3}


def _host_key() -> str:
    """-march=native binds the library to the CPU it was built on: a copy that travelled to another host is rebuilt."""
    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name") and not model:
                    model = line.split(":", 1)[1].strip()
                elif line.startswith("flags") and not flags:
                    flags = line.split(":", 1)[1].strip()
                if model and flags:
                    break
    except OSError:
        pass
    import hashlib
    return hashlib.sha1((model + "|" + flags + "|" + (_HERE / "Makefile").read_text()).encode()).hexdigest()


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc -O3 -march=native -ffp-contract=off)."""
    src_m = max((_HERE / "jx_oracle.c").stat().st_mtime, (_HERE / "jx_oracle.h").stat().st_mtime)
    stamp = _HERE / "_build" / "host.key"
    key = _host_key()
    stale = not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m or not stamp.exists() or stamp.read_text() != key
    if force or stale:
        subprocess.run(["make", "-B", "-C", str(_HERE)], check=True, capture_output=True)
        stamp.write_text(key)
    return _LIB_PATH


def _p(a: Optional[np.ndarray], ty):
    if a is None:
        return C.cast(None, C.POINTER(ty))
    return a.ctypes.data_as(C.POINTER(ty))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.jxo_reml_loglike.restype = C.c_double
        _lib.jxo_ml_loglike.restype = C.c_double
        _lib.jxo_normal_sf.restype = C.c_double
        _lib.jxo_normal_sf.argtypes = [C.c_double]
        _lib.jxo_chi2_sf_df1.restype = C.c_double
        _lib.jxo_chi2_sf_df1.argtypes = [C.c_double]
        _lib.jxo_format_row.restype = C.c_size_t
        _lib.jxo_fmt_exp.restype = C.c_size_t
        _lib.jxo_fmt_fixed.restype = C.c_size_t
        _lib.jxo_qc_row.restype = C.c_int
        _lib.jxo_fixed_cache_prepare.restype = C.c_int
        _lib.jxo_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().jxo_max_threads())


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _idx(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


# ------------------------------------------------------------------------------------------
# A2/A3/A4: counts, QC, decode
# ------------------------------------------------------------------------------------------
def count_qc_block(packed: np.ndarray, n_full: int, sample_idx=None, maf_thr=0.02, miss_thr=0.05,
                   het_thr=1.0):
    """packed: u8[rows, ceil(n_full/4)] -> (keep bool[rows], af f32, miss_rate f32, missing i64)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    rows, bps = packed.shape
    sidx = _idx(sample_idx)
    n_sel = 0 if sidx is None else sidx.shape[0]
    keep = np.zeros(rows, dtype=np.uint8)
    af = np.zeros(rows, dtype=np.float32)
    mr = np.zeros(rows, dtype=np.float32)
    missing = np.zeros(rows, dtype=np.int64)
    lib().jxo_count_qc_block(_p(packed, C.c_uint8), C.c_size_t(bps), C.c_size_t(rows), C.c_size_t(n_full),
                             _p(sidx, C.c_int64), C.c_size_t(n_sel),
                             C.c_float(maf_thr), C.c_float(miss_thr), C.c_float(het_thr),
                             _p(keep, C.c_uint8), _p(af, C.c_float), _p(mr, C.c_float), _p(missing, C.c_int64))
    return keep.astype(bool), af, mr, missing


def decode_centered_block(packed: np.ndarray, n_full: int, maf: np.ndarray, sample_idx=None,
                          row_indices=None, flip=None, model: str = "add") -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    bps = packed.shape[1]
    ridx = _idx(row_indices)
    rows = packed.shape[0] if ridx is None else ridx.shape[0]
    sidx = _idx(sample_idx)
    n = n_full if sidx is None else sidx.shape[0]
    maf = _f32(maf)
    assert maf.shape[0] == rows
    fl = None if flip is None else np.ascontiguousarray(np.asarray(flip, dtype=np.uint8))
    out = np.zeros((rows, n), dtype=np.float32)
    lib().jxo_decode_centered_block(_p(packed, C.c_uint8), C.c_size_t(bps), _p(ridx, C.c_int64),
                                    C.c_size_t(rows), C.c_size_t(n_full), _p(sidx, C.c_int64), C.c_size_t(n),
                                    _p(fl, C.c_uint8), _p(maf, C.c_float), C.c_int(MODEL_CODES[model]),
                                    _p(out, C.c_float))
    return out


# ------------------------------------------------------------------------------------------
# A5/A6: rotations
# ------------------------------------------------------------------------------------------
def rotate_block(g: np.ndarray, u_t: np.ndarray, mode: int = 0) -> np.ndarray:
    """mode 0 = ROT_F64_F32STORE (parity oracle); mode 1 = sequential f32 accumulate."""
    g = _f32(g)
    u_t = _f32(u_t)
    rows, n = g.shape
    assert u_t.shape == (n, n)
    out = np.zeros((rows, n), dtype=np.float32)
    lib().jxo_rotate_block(_p(g, C.c_float), C.c_size_t(rows), C.c_size_t(n), _p(u_t, C.c_float),
                           _p(out, C.c_float), C.c_int(mode))
    return out


def lmm_rotate_x_y_with_ut_f64(u_t, x, y, threads: int = 0):
    u_t = _f32(u_t)
    x = _f64(x)
    y = _f64(y).reshape(-1)
    n = y.shape[0]
    if n == 0:
        raise RuntimeError("y must not be empty")
    if x.ndim != 2 or x.shape[0] != n:
        raise RuntimeError(f"x rows must equal len(y): rows={x.shape[0]}, len(y)={n}")
    if u_t.shape != (n, n):
        raise RuntimeError("u_t must be shape (n, n) and row-major U^T")
    q = x.shape[1]
    xr = np.zeros((n, q), dtype=np.float64)
    yr = np.zeros((n, 1), dtype=np.float64)
    lib().jxo_rotate_xy(_p(u_t, C.c_float), C.c_size_t(n), _p(x, C.c_double), C.c_size_t(q),
                        _p(y, C.c_double), _p(xr, C.c_double), _p(yr, C.c_double))
    return xr, yr


# ------------------------------------------------------------------------------------------
# A11/A12/A13 scalars
# ------------------------------------------------------------------------------------------
def _model_args(s, xcov, y):
    s = _f64(s).reshape(-1)
    xcov = _f64(xcov)
    y = _f64(y).reshape(-1)
    n = y.shape[0]
    if xcov.ndim != 2 or xcov.shape[0] != n:
        raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
    if s.shape[0] != n:
        raise RuntimeError("len(S) must equal len(y_rot)")
    return s, xcov, y, n, xcov.shape[1]


def reml_loglike(log10_lbd, s, xcov, y, snp=None) -> float:
    s, xcov, y, n, p = _model_args(s, xcov, y)
    snp_a = None if snp is None else _f64(snp)
    return float(lib().jxo_reml_loglike(C.c_double(log10_lbd), _p(s, C.c_double), _p(xcov, C.c_double),
                                        _p(y, C.c_double), _p(snp_a, C.c_double), C.c_size_t(n), C.c_size_t(p)))


def ml_loglike(log10_lbd, s, xcov, y, snp=None) -> float:
    s, xcov, y, n, p = _model_args(s, xcov, y)
    snp_a = None if snp is None else _f64(snp)
    return float(lib().jxo_ml_loglike(C.c_double(log10_lbd), _p(s, C.c_double), _p(xcov, C.c_double),
                                      _p(y, C.c_double), _p(snp_a, C.c_double), C.c_size_t(n), C.c_size_t(p)))


def ml_loglike_null_f32(s, xcov, y_rot, log10_lbd) -> float:
    return ml_loglike(log10_lbd, s, xcov, y_rot, None)


def final_beta_se(log10_lbd, s, xcov, y, snp) -> Tuple[float, float, float]:
    s, xcov, y, n, p = _model_args(s, xcov, y)
    snp_a = _f64(snp)
    out = np.zeros(3)
    lib().jxo_final_beta_se(C.c_double(log10_lbd), _p(s, C.c_double), _p(xcov, C.c_double), _p(y, C.c_double),
                            _p(snp_a, C.c_double), C.c_size_t(n), C.c_size_t(p), _p(out, C.c_double))
    return float(out[0]), float(out[1]), float(out[2])


_COST = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def brent_minimize(f, low, high, tol, max_iter, init_x=None):
    """src/math/brent.rs -- returns (x, f(x), n_eval) for a Python callable."""
    cb = _COST(lambda x, _ctx: float(f(x)))
    bx = C.c_double()
    bf = C.c_double()
    ne = C.c_size_t()
    lib().jxo_brent(cb, None, C.c_double(low), C.c_double(high), C.c_double(tol), C.c_size_t(max_iter),
                    C.c_int(0 if init_x is None else 1), C.c_double(0.0 if init_x is None else init_x),
                    C.byref(bx), C.byref(bf), C.byref(ne))
    return bx.value, bf.value, ne.value


def normal_sf(z: float) -> float:
    return float(lib().jxo_normal_sf(z))


def chi2_sf_df1(stat: float) -> float:
    return float(lib().jxo_chi2_sf_df1(stat))


def lmm_reml_null_f32(s, xcov, y_rot, low, high, max_iter=50, tol=1e-2):
    s, xcov, y, n, p = _model_args(s, xcov, y_rot)
    if low >= high:
        raise RuntimeError("low must be < high")
    out = np.zeros(3)
    lib().jxo_reml_null(_p(s, C.c_double), _p(xcov, C.c_double), _p(y, C.c_double), C.c_size_t(n), C.c_size_t(p),
                        C.c_double(low), C.c_double(high), C.c_size_t(max_iter), C.c_double(tol), _p(out, C.c_double))
    return float(out[0]), float(out[1]), float(out[2])


def lmm_ml_null_brent(s, xcov, y_rot, low, high, max_iter=30, tol=1e-2, init=None):
    """src/stats/lmm.rs:2901-2924 -> (log10_lbd_ml, ml0)."""
    s, xcov, y, n, p = _model_args(s, xcov, y_rot)
    out = np.zeros(2)
    lib().jxo_ml_null(_p(s, C.c_double), _p(xcov, C.c_double), _p(y, C.c_double), C.c_size_t(n), C.c_size_t(p),
                      C.c_double(low), C.c_double(high), C.c_size_t(max_iter), C.c_double(tol),
                      C.c_int(0 if init is None else 1), C.c_double(0.0 if init is None else init),
                      _p(out, C.c_double))
    return float(out[0]), float(out[1])


# ------------------------------------------------------------------------------------------
# A9/A10: per-SNP scans on rotated blocks; *_from_snp variants rotate first (mode 0 by default)
# ------------------------------------------------------------------------------------------
def lmm_reml_chunk_f32(s, xcov, y_rot, low, high, g_rot_chunk, max_iter=50, tol=1e-2, threads=0,
                       nullml=None, init_log10_lbd=None, return_evals=False):
    s, xcov, y, n, p = _model_args(s, xcov, y_rot)
    g = _f32(g_rot_chunk)
    if g.ndim != 2 or g.shape[1] != n:
        raise RuntimeError("g_rot_chunk must be (m_chunk, n)")
    if low >= high:
        raise RuntimeError("low must be < high")
    m = g.shape[0]
    cols = 4 if nullml is not None else 3
    out = np.zeros((m, cols), dtype=np.float64)
    ne = np.zeros(m, dtype=np.int32)
    lib().jxo_lmm_reml_block(_p(g, C.c_float), C.c_size_t(m), C.c_size_t(n), _p(s, C.c_double),
                             _p(xcov, C.c_double), _p(y, C.c_double), C.c_size_t(p),
                             C.c_double(low), C.c_double(high), C.c_double(tol), C.c_size_t(max_iter),
                             C.c_int(0 if init_log10_lbd is None else 1),
                             C.c_double(0.0 if init_log10_lbd is None else init_log10_lbd),
                             C.c_int(0 if nullml is None else 1), C.c_double(0.0 if nullml is None else nullml),
                             _p(out, C.c_double), _p(ne, C.c_int32), C.c_int(threads))
    return (out, ne) if return_evals else out


def lmm_reml_lmm2_chunk_f32(s, xcov, y_rot, low, high, g_rot_chunk, nullml, max_iter=50, tol=1e-2,
                            threads=0, init_reml=None, init_ml=None, return_evals=False):
    s, xcov, y, n, p = _model_args(s, xcov, y_rot)
    g = _f32(g_rot_chunk)
    if g.ndim != 2 or g.shape[1] != n:
        raise RuntimeError("g_rot_chunk must be (m_chunk, n)")
    m = g.shape[0]
    out = np.zeros((m, 6), dtype=np.float64)
    ne = np.zeros(m, dtype=np.int32)
    lib().jxo_lmm2_block(_p(g, C.c_float), C.c_size_t(m), C.c_size_t(n), _p(s, C.c_double),
                         _p(xcov, C.c_double), _p(y, C.c_double), C.c_size_t(p),
                         C.c_double(low), C.c_double(high), C.c_double(tol), C.c_size_t(max_iter),
                         C.c_int(0 if init_reml is None else 1), C.c_double(0.0 if init_reml is None else init_reml),
                         C.c_int(0 if init_ml is None else 1), C.c_double(0.0 if init_ml is None else init_ml),
                         C.c_double(nullml), _p(out, C.c_double), _p(ne, C.c_int32), C.c_int(threads))
    return (out, ne) if return_evals else out


def lmm_reml_chunk_from_snp_f32(s, xcov, y_rot, low, high, snp_chunk, u_t, max_iter=50, tol=1e-2, threads=0,
                                nullml=None, rotate_block_rows=256, rot_mode=0):
    g_rot = rotate_block(snp_chunk, u_t, mode=rot_mode)
    return lmm_reml_chunk_f32(s, xcov, y_rot, low, high, g_rot, max_iter, tol, threads, nullml)


def lmm_reml_lmm2_chunk_from_snp_f32(s, xcov, y_rot, low, high, snp_chunk, u_t, nullml, max_iter=50, tol=1e-2,
                                     threads=0, rotate_block_rows=256, rot_mode=0):
    g_rot = rotate_block(snp_chunk, u_t, mode=rot_mode)
    return lmm_reml_lmm2_chunk_f32(s, xcov, y_rot, low, high, g_rot, nullml, max_iter, tol, threads)


# ------------------------------------------------------------------------------------------
# A14: fixed lambda
# ------------------------------------------------------------------------------------------
class _FixedCache(C.Structure):
    _fields_ = [("n", C.c_size_t), ("p", C.c_size_t), ("w", C.POINTER(C.c_float)),
                ("py_tilde", C.POINTER(C.c_float)), ("wx_tilde", C.POINTER(C.c_float)),
                ("a_chol", C.POINTER(C.c_double)), ("ypy", C.c_double), ("log_det_v", C.c_double),
                ("df", C.c_int)]


_FIXED_ERR = {-1: "non-positive s[i]+lbd", -2: "X'WX not SPD", -3: "df <= 0", -4: "too many covariates"}


def lmm_assoc_chunk_f32(s, xcov, y_rot, log10_lbd, g_rot_chunk, threads=0, nullml=None):
    """Fixed-lambda scan of a rotated block (src/stats/fvlmm.rs:1484-1563, 1691-1805)."""
    s, xcov, y, n, p = _model_args(s, xcov, y_rot)
    g = _f32(g_rot_chunk)
    if g.ndim != 2 or g.shape[1] != n:
        raise RuntimeError("g_rot_chunk must be (m_chunk, n)")
    if n <= p + 1:
        raise RuntimeError("n must be > p_cov+1")
    cache = _FixedCache()
    rc = lib().jxo_fixed_cache_prepare(_p(s, C.c_double), _p(xcov, C.c_double), _p(y, C.c_double),
                                       C.c_size_t(n), C.c_size_t(p), C.c_double(10.0 ** log10_lbd), C.byref(cache))
    try:
        if rc != 0:
            raise RuntimeError(_FIXED_ERR.get(rc, f"fixed cache error {rc}"))
        m = g.shape[0]
        cols = 4 if nullml is not None else 3
        out = np.zeros((m, cols), dtype=np.float64)
        lib().jxo_fixed_lambda_block(_p(g, C.c_float), C.c_size_t(m), C.byref(cache),
                                     C.c_int(0 if nullml is None else 1),
                                     C.c_double(0.0 if nullml is None else nullml), _p(out, C.c_double),
                                     C.c_int(threads))
        meta = {"ypy": cache.ypy, "log_det_v": cache.log_det_v, "df": cache.df}
    finally:
        lib().jxo_fixed_cache_free(C.byref(cache))
    return out, meta


def lmm_assoc_chunk_from_snp_f32(s, xcov, y_rot, log10_lbd, snp_chunk, u_t, threads=0, nullml=None, rot_mode=0):
    g_rot = rotate_block(snp_chunk, u_t, mode=rot_mode)
    return lmm_assoc_chunk_f32(s, xcov, y_rot, log10_lbd, g_rot, threads, nullml)[0]


# ------------------------------------------------------------------------------------------
# A16: TSV formatting
# ------------------------------------------------------------------------------------------
HEADERS = {
    3: b"chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\n",
    4: b"chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\tplrt\n",
    6: b"chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\tlambda\tml\tplrt\n",
}


def fmt_exp(v: float, prec: int) -> str:
    buf = C.create_string_buffer(96)
    lib().jxo_fmt_exp(buf, C.c_size_t(96), C.c_double(v), C.c_int(prec))
    return buf.value.decode()


def fmt_fixed(v: float, prec: int) -> str:
    buf = C.create_string_buffer(96)
    lib().jxo_fmt_fixed(buf, C.c_size_t(96), C.c_double(v), C.c_int(prec))
    return buf.value.decode()


def format_row(chrom: str, pos: int, snp: str, a0: str, a1: str, af: float, miss_rate: float,
               row: Sequence[float]) -> bytes:
    r = _f64(row)
    buf = C.create_string_buffer(2048)
    n = lib().jxo_format_row(buf, C.c_size_t(2048), chrom.encode(), C.c_int64(pos), snp.encode(), a0.encode(),
                             a1.encode(), C.c_float(af), C.c_float(miss_rate), _p(r, C.c_double),
                             C.c_int(r.shape[0]))
    return buf.raw[:n]


# ------------------------------------------------------------------------------------------
# Route B: BedChunkReader.next_chunk_prepared rows (additive coding) -- numpy restatement
# ------------------------------------------------------------------------------------------
def bed_chunk_prepared_rows(packed, n_samples, sample_idx=None, maf_thr=0.0, miss_thr=1.0, het_thr=1.0):
    """src/io/gfreader.rs:3580-3700 + process_snp_row_with_precomputed_counts_impl (src/io/gfcore.rs:405-480,
    preserve_alt_orientation=true, fill_missing=true), row by row exactly as written there: decode to f32 dosages
    (-9 = missing), f64 QC, impute with (alt_sum / non_missing) as f32, centre with (f64 sum / n) as f32.
    -> (keep bool[m], g f32[kept, n], af f32[kept], miss f32[kept])."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    idx = np.arange(n_samples) if sample_idx is None else np.asarray(sample_idx, dtype=np.int64)
    n = idx.shape[0]
    lut = np.array([0.0, -9.0, 1.0, 2.0], dtype=np.float32)
    maf_t, miss_t, het_t = np.float32(maf_thr), np.float32(miss_thr), np.float32(het_thr)
    apply_het = het_t < np.float32(1.0)
    keep = np.zeros(packed.shape[0], dtype=bool)
    rows, afs, misses = [], [], []
    for r in range(packed.shape[0]):
        codes = (packed[r, idx >> 2] >> (2 * (idx & 3)).astype(np.uint8)) & 3
        row = lut[codes].copy()
        obs = row >= 0.0
        non_missing = int(obs.sum())
        alt_sum = 0.0
        for v in row[obs]:                       # scan_snp_row_counts: sequential f64 sum
            alt_sum += float(v)
        het_count = int((np.abs(row[obs] - np.float32(1.0)) < np.float32(1e-6)).sum())
        missing_rate = np.float32(1.0 - non_missing / float(n))
        missing_count = n - non_missing
        if missing_rate > miss_t:
            continue
        if non_missing == 0:
            if maf_t > 0:
                continue
            row[:] = 0.0
        else:
            if apply_het and het_count / float(non_missing) > float(het_t):
                continue
            alt_freq = alt_sum / (2.0 * non_missing)
            if np.float32(min(alt_freq, 1.0 - alt_freq)) < maf_t:
                continue
            row[~obs] = np.float32(alt_sum / non_missing)
        total = 0.0
        for v in row:
            total += float(v)
        coded_mean = np.float32(total / float(n))
        keep[r] = True
        rows.append((row - coded_mean).astype(np.float32))
        afs.append(np.float32(coded_mean * np.float32(0.5)))
        misses.append(np.float32(missing_count))
    g = np.array(rows, dtype=np.float32).reshape(len(rows), n)
    return keep, g, np.array(afs, dtype=np.float32), np.array(misses, dtype=np.float32)


# ------------------------------------------------------------------------------------------
# N1: centred additive GRM -- numpy restatement, f64 contraction
# ------------------------------------------------------------------------------------------
def grm_packed_f64(packed, n_samples, row_flip, row_maf, sample_indices=None, method=1, mu_grid_bits=None):
    """src/stats/grm.rs:204-608 with decode_additive_grm_block_f32 (src/decode/decode.rs:803-845, subset path
    src/math/bedmath.rs:1359-1440): z = LUT[code], LUT = [0-mu, 0 (missing), 1-mu, 2-mu] in f32, mu = 2*clamp(p,0,1);
    K = Z Z^T / sum_s f64(f32(2p(1-p))), symmetric.  The reference contracts with an f32 SYRK (order-dependent);
    this restatement contracts in f64.  `mu_grid_bits=b` rounds mu to a 2^-b grid first (the device kernel's
    arithmetic), in which case z is exact in f64."""
    assert method == 1, "only the centred additive GRM is restated"
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    m, bps = packed.shape
    assert bps == (n_samples + 3) // 4
    assert not np.any(np.asarray(row_flip, dtype=bool))
    idx = np.arange(n_samples) if sample_indices is None else np.asarray(sample_indices, dtype=np.int64)
    p = np.clip(np.asarray(row_maf, dtype=np.float32), np.float32(0), np.float32(1))
    p = np.where(np.isnan(p), np.float32(0), p).astype(np.float32)
    mean_g = (np.float32(2.0) * p).astype(np.float32)
    var = (np.float32(2.0) * p * (np.float32(1.0) - p)).astype(np.float32)
    varsum = float(np.sum(np.where(np.isfinite(var) & (var > 0), var.astype(np.float64), 0.0)))
    if not (np.isfinite(varsum) and varsum > 0.0):
        raise RuntimeError("invalid centered GRM denominator: sum(2p(1-p)) <= 0")
    n = idx.shape[0]
    K = np.zeros((n, n), dtype=np.float64)
    for r0 in range(0, m, 4096):
        blk = packed[r0:r0 + 4096]
        codes = (blk[:, idx >> 2] >> (2 * (idx & 3)).astype(np.uint8)) & 3
        mg = mean_g[r0:r0 + 4096]
        if mu_grid_bits is None:
            lut = np.stack([np.float32(0) - mg, np.zeros_like(mg), np.float32(1) - mg, np.float32(2) - mg], axis=1)
            lut = lut.astype(np.float32).astype(np.float64)
        else:
            mu = np.rint(mg.astype(np.float64) * 2.0 ** mu_grid_bits) / 2.0 ** mu_grid_bits
            lut = np.stack([0.0 - mu, np.zeros_like(mu), 1.0 - mu, 2.0 - mu], axis=1)
        z = np.take_along_axis(lut, codes.astype(np.int64), axis=1)      # [rows, n]
        K += z.T @ z
    return K / varsum, varsum


# ------------------------------------------------------------------------------------------
# A1/A15: PLINK readers + the BED -> TSV scan, composed from the pieces above
# ------------------------------------------------------------------------------------------
def read_fam(prefix: str):
    """src/io/gfcore.rs:307-324 -- IID = 2nd whitespace column."""
    ids = []
    with open(f"{prefix}.fam") as fh:
        for line in fh:
            tok = line.split()
            if len(tok) < 2:
                raise RuntimeError(f"Malformed FAM line: {line.rstrip()}")
            ids.append(tok[1])
    return ids


def read_bim(prefix: str):
    """src/io/gfcore.rs:1426-1478 -- (chrom, snp, pos(i32 or 0), a1, a2)."""
    sites = []
    with open(f"{prefix}.bim") as fh:
        for ln, line in enumerate(fh, 1):
            tok = line.split()
            if len(tok) < 6:
                raise RuntimeError(f"Malformed BIM line at {prefix}.bim:{ln}: {line.rstrip()}")
            try:
                pos = int(tok[3])
                if not (-2**31 <= pos < 2**31):
                    pos = 0
            except ValueError:
                pos = 0
            sites.append((tok[0], tok[1], pos, tok[4], tok[5]))
    return sites


def read_bed(prefix: str, n_full: int) -> np.ndarray:
    raw = np.fromfile(f"{prefix}.bed", dtype=np.uint8)
    if raw.shape[0] < 3 or raw[0] != 0x6C or raw[1] != 0x1B or raw[2] != 0x01:
        raise RuntimeError("only SNP-major BED supported")
    bps = (n_full + 3) // 4
    if (raw.shape[0] - 3) % bps != 0:
        raise RuntimeError(f"BED payload length {raw.shape[0] - 3} not a multiple of {bps}")
    return raw[3:].reshape(-1, bps)


def _is_simple_snp_allele(a: str) -> bool:
    t = a.strip().upper()
    return len(t) == 1 and t in "ACGT"


def scan_bed_to_tsv(bed_prefix, out_tsv, s, xcov, y_rot, u_t, maf_thr, miss_thr, het_thr, genetic_model="add",
                    snps_only=False, sample_ids=None, low=-5.0, high=5.0, max_iter=30, tol=1e-2, threads=0,
                    nullml=None, init_log10_lbd=None, rotate_block_rows=512, model="lmm",
                    init_log10_lbd_ml=None, log10_lbd=None, rot_mode=0) -> int:
    """src/stats/lmm.rs:975-1477 with warm start disabled (JX_LMM_UNIFIED_NO_WARM_START=1).

    model: "lmm" (3|4 cols), "lmm2" (6 cols; null ML fitted when nullml is None), "fvlmm" (fixed lambda).
    """
    if low >= high:
        raise RuntimeError("low must be < high")
    fam = read_fam(bed_prefix)
    n_full = len(fam)
    if sample_ids is not None:
        pos = {sid: i for i, sid in enumerate(fam)}
        try:
            sidx = np.array([pos[sid] for sid in sample_ids], dtype=np.int64)
        except KeyError as ex:
            raise RuntimeError(f"sample '{ex.args[0]}' not found in PLINK FAM")
    else:
        sidx = None
    n = n_full if sidx is None else sidx.shape[0]
    s_, xcov_, y_, n_model, p = _model_args(s, xcov, y_rot)
    if n != n_model:
        raise RuntimeError(f"sample_ids length {n} != expected sample count {n_model}")
    identity = sidx is None or (n == n_full and np.array_equal(sidx, np.arange(n_full)))
    sidx_eff = None if identity else sidx
    packed = read_bed(bed_prefix, n_full)
    sites = read_bim(bed_prefix)
    m = packed.shape[0]
    if len(sites) != m:
        raise RuntimeError("BIM row count does not match BED")
    if init_log10_lbd is not None and np.isfinite(init_log10_lbd):
        init_log10_lbd = min(max(init_log10_lbd, low), high)
    else:
        init_log10_lbd = None
    if model == "lmm2":
        out_cols = 6
        if nullml is None:
            init = init_log10_lbd_ml if init_log10_lbd_ml is not None else init_log10_lbd
            _, nullml = lmm_ml_null_brent(s_, xcov_, y_, low, high, max_iter, tol, init)
            if not np.isfinite(nullml):
                raise RuntimeError("failed to optimize null ML for LMM2 unified scan")
    else:
        out_cols = 4 if nullml is not None else 3
    rows_written = 0
    with open(out_tsv, "wb") as fh:
        fh.write(HEADERS[out_cols])
        step = max(1, int(rotate_block_rows))
        for c0 in range(0, m, step):
            blk = packed[c0:c0 + step]
            keep, af, mr, missing = count_qc_block(blk, n_full, sidx_eff, maf_thr, miss_thr, het_thr)
            if snps_only:
                for j in range(blk.shape[0]):
                    st = sites[c0 + j]
                    if keep[j] and not (_is_simple_snp_allele(st[3]) and _is_simple_snp_allele(st[4])):
                        keep[j] = False
            idx = np.nonzero(keep)[0]
            if idx.size == 0:
                continue
            g = decode_centered_block(blk, n_full, af[idx], sidx_eff, row_indices=idx, model=genetic_model)
            g_rot = rotate_block(g, u_t, mode=rot_mode)
            if model == "lmm":
                # NO_WARM_START: seed_with_init_guess=false -> midpoint start (lmm.rs:134-140)
                res = lmm_reml_chunk_f32(s_, xcov_, y_, low, high, g_rot, max_iter, tol, threads, nullml)
            elif model == "lmm2":
                res = lmm_reml_lmm2_chunk_f32(s_, xcov_, y_, low, high, g_rot, nullml, max_iter, tol, threads,
                                              init_reml=init_log10_lbd, init_ml=init_log10_lbd_ml)
            elif model == "fvlmm":
                res, _ = lmm_assoc_chunk_f32(s_, xcov_, y_, log10_lbd, g_rot, threads, nullml)
            else:
                raise ValueError(model)
            for k, j in enumerate(idx):
                chrom, snp, pos, a0, a1 = sites[c0 + j]
                # lmm.rs:2667-2670: miss column = missing_count / n as f32
                miss_rate = np.float32(missing[j]) / np.float32(n)
                fh.write(format_row(chrom, pos, snp, a0, a1, float(af[j]), float(miss_rate), res[k]))
            rows_written += int(idx.size)
    return rows_written
