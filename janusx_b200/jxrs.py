"""`janusx.janusx`-compatible function surface for the exact-LMM path, backed by libjxb200.so.

Every function here keeps the name, argument order, defaults, return layout and error text of the
reference PyO3 function it replaces (registered at src/lib.rs:911-941 of JanusX); the body is a thin
ctypes call into the C ABI (include/jxb200.h).  `threads`, `rotate_block_rows` and `mmap_window_mb`
are accepted for signature compatibility; they steer CPU thread pools / host tiling in the reference
and have no meaning on the device path (rotate_block_rows is used as the device batch size hint by
the file-level scans).

The null model (S, Xcov, y_rot, U^T) is uploaded once per distinct set of arrays and kept resident in
HBM (`DeviceModel`); the array-argument functions look the handle up in a small cache so a chunk loop
that passes the same arrays every call -- exactly what python/janusx/pyBLUP/assoc.py does -- pays the
U^T upload once.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
from collections import OrderedDict
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from ._cabi import BedScanCfg, JxbError, QcCfg, SolveCfg, check, lib, ptr

MODEL_CODES = {"add": 0, "dom": 1, "rec": 2, "het": 3}
# SNP rows per device batch of the file-level scans.  75,776 = one full wave of the thread-per-SNP solve kernel on a
# 148-SM B200 (148 SMs x 16 warps x 32 SNPs); two waves per batch hide most of the wave's ragged tail (Brent paths of
# different length): +2.9 % at n = 20,000.  Per-batch workspace is ~300 bytes x n per row, so large n stays at one wave.
# The reference's rotate_block_rows (default 512) sizes CPU tiles and has no meaning here.
DEFAULT_DEVICE_BATCH = 75776


def default_device_batch(n: int) -> int:
    return 2 * DEFAULT_DEVICE_BATCH if n <= 24000 else DEFAULT_DEVICE_BATCH


__all__ = [
    "DeviceModel", "lmm_reml_chunk_f32", "lmm_reml_chunk_from_snp_f32", "lmm_reml_lmm2_chunk_from_snp_f32",
    "lmm_assoc_chunk_f32", "lmm_assoc_chunk_from_snp_f32", "lmm_reml_null_f32", "ml_loglike_null_f32",
    "lmm_rotate_x_y_with_ut_f64", "lmm_reml_assoc_bed_to_tsv_f32", "lmm_reml_lmm2_assoc_bed_to_tsv_f32",
    "fvlmm_assoc_bed_to_tsv_f32", "fvlmm_assoc_chunk_f32", "fvlmm_assoc_chunk_from_snp_f32",
    "gwas_lmm_lm_null_lrt_decision",
]


def _f64(a, name="array"):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _model_code(genetic_model: str) -> int:
    key = str(genetic_model).lower()
    if key not in MODEL_CODES:
        raise RuntimeError("model must be one of: add, dom, rec, het")
    return MODEL_CODES[key]


def _solve_cfg(low=-5.0, high=5.0, max_iter=30, tol=1e-2, init=None, nullml=None) -> SolveCfg:
    c = SolveCfg()
    c.low, c.high, c.tol, c.max_iter = float(low), float(high), float(tol), int(max_iter)
    c.has_init = 0 if init is None else 1
    c.init_log10_lbd = 0.0 if init is None else float(init)
    c.has_nullml = 0 if nullml is None else 1
    c.nullml = 0.0 if nullml is None else float(nullml)
    return c


class DeviceModel:
    """The null model resident in HBM on one device (opaque jxb_model handle)."""

    def __init__(self, s, xcov, y_rot, u_t=None, device: int = 0, u_t_on_device: bool = False):
        _cabi.require_gpu()
        self.s = _f64(s).reshape(-1)
        self.xcov = _f64(xcov)
        self.y = _f64(y_rot).reshape(-1)
        n = self.y.shape[0]
        if self.xcov.ndim != 2 or self.xcov.shape[0] != n:
            raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
        if self.s.shape[0] != n:
            raise RuntimeError("len(S) must equal len(y_rot)")
        self.n, self.p = n, int(self.xcov.shape[1])
        self.device = int(device)
        self.has_ut = u_t is not None
        h = C.c_void_p()
        if u_t is not None and u_t_on_device:
            # torch CUDA tensor (f32, contiguous, [n, n]) e.g. straight out of an NCCL broadcast
            if tuple(u_t.shape) != (n, n):
                raise RuntimeError("u_t must be (n, n) and row-major U^T")
            import torch  # plumbing only

            sd = torch.as_tensor(self.s, device=u_t.device)
            xd = torch.as_tensor(self.xcov, device=u_t.device)
            yd = torch.as_tensor(self.y, device=u_t.device)
            check(lib().jxb_model_create_dev(self.device, n, self.p, ptr(sd), ptr(xd), ptr(yd), ptr(u_t), C.byref(h)))
        else:
            ut = None
            if u_t is not None:
                ut = _f32(u_t)
                if ut.shape != (n, n):
                    raise RuntimeError("u_t must be (n, n) and row-major U^T")
            check(lib().jxb_model_create(self.device, n, self.p, ptr(self.s), ptr(self.xcov), ptr(self.y), ptr(ut),
                                         C.byref(h)))
        self._h = h

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib().jxb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise JxbError("model is closed")
        return self._h

    def set_xy(self, xcov, y_rot):
        xcov = _f64(xcov)
        y = _f64(y_rot).reshape(-1)
        if xcov.shape != (self.n, self.p) or y.shape[0] != self.n:
            raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
        check(lib().jxb_model_set_xy(self.handle, ptr(xcov), ptr(y)))
        self.xcov, self.y = xcov, y

    def sync(self):
        check(lib().jxb_model_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(lib().jxb_model_stream(self.handle) or 0)

    # -- null model -------------------------------------------------------------------------------
    def reml_null(self, low, high, max_iter=50, tol=1e-2) -> Tuple[float, float, float]:
        if low >= high:
            raise RuntimeError("low must be < high")
        out = (C.c_double * 3)()
        check(lib().jxb_reml_null(self.handle, low, high, int(max_iter), tol, out))
        return float(out[0]), float(out[1]), float(out[2])

    def ml_loglike_null(self, log10_lbd) -> float:
        out = C.c_double()
        check(lib().jxb_ml_loglike_null(self.handle, float(log10_lbd), C.byref(out)))
        return float(out.value)

    def ml_null(self, low, high, max_iter=30, tol=1e-2, init=None) -> Tuple[float, float]:
        out = (C.c_double * 2)()
        check(lib().jxb_ml_null(self.handle, low, high, int(max_iter), tol, 0 if init is None else 1,
                                0.0 if init is None else float(init), out))
        return float(out[0]), float(out[1])

    def rotate_xy(self, x, y):
        x = _f64(x)
        y = _f64(y).reshape(-1)
        xr = np.empty((self.n, x.shape[1]), dtype=np.float64)
        yr = np.empty((self.n, 1), dtype=np.float64)
        check(lib().jxb_rotate_xy(self.handle, ptr(x), x.shape[1], ptr(y), ptr(xr), ptr(yr)))
        return xr, yr

    # -- chunk scans ------------------------------------------------------------------------------
    def _chunk_in(self, a, what):
        a = _f32(a)
        if a.ndim != 2 or a.shape[1] != self.n:
            raise RuntimeError(f"{what} must be (m_chunk, n)")
        return a

    def lmm_reml_chunk(self, g, low, high, max_iter=50, tol=1e-2, nullml=None, rotated=True, init=None,
                       return_evals=False):
        g = self._chunk_in(g, "g_rot_chunk" if rotated else "snp_chunk")
        if low >= high:
            raise RuntimeError("low must be < high")
        m = g.shape[0]
        cfg = _solve_cfg(low, high, max_iter, tol, init, nullml)
        out = np.zeros((m, 4 if nullml is not None else 3), dtype=np.float64)
        ev = np.zeros(m, dtype=np.int32)
        fn = lib().jxb_lmm_reml_chunk_f32 if rotated else lib().jxb_lmm_reml_chunk_from_snp_f32
        check(fn(self.handle, ptr(g), m, C.byref(cfg), ptr(out), ptr(ev)))
        return (out, ev) if return_evals else out

    def lmm2_chunk(self, g, low, high, nullml, max_iter=50, tol=1e-2, rotated=False, init=None, return_evals=False):
        g = self._chunk_in(g, "g_rot_chunk" if rotated else "snp_chunk")
        if low >= high:
            raise RuntimeError("low must be < high")
        m = g.shape[0]
        cfg = _solve_cfg(low, high, max_iter, tol, init, nullml)
        out = np.zeros((m, 6), dtype=np.float64)
        ev = np.zeros(m, dtype=np.int32)
        check(lib().jxb_lmm2_chunk_f32(self.handle, ptr(g), m, 1 if rotated else 0, C.byref(cfg), ptr(out), ptr(ev)))
        return (out, ev) if return_evals else out

    def fixed_chunk(self, g, log10_lbd, nullml=None, rotated=False, return_meta=False):
        g = self._chunk_in(g, "g_rot_chunk" if rotated else "snp_chunk")
        if self.n <= self.p + 1:
            raise RuntimeError("n must be > p_cov+1")
        m = g.shape[0]
        out = np.zeros((m, 4 if nullml is not None else 3), dtype=np.float64)
        meta = (C.c_double * 3)()
        nm = None if nullml is None else C.c_double(float(nullml))
        check(lib().jxb_lmm_fixed_chunk_f32(self.handle, ptr(g), m, 1 if rotated else 0, float(log10_lbd),
                                            C.byref(nm) if nm is not None else None, ptr(out), meta))
        if return_meta:
            return out, {"ypy": float(meta[0]), "log_det_v": float(meta[1]), "df": int(meta[2])}
        return out

    def rotate_block(self, snp, variant: int = 0) -> np.ndarray:
        snp = self._chunk_in(snp, "snp_chunk")
        out = np.empty_like(snp)
        check(lib().jxb_rotate_block_f32(self.handle, ptr(snp), snp.shape[0], ptr(out), int(variant)))
        return out

    # -- packed scans -----------------------------------------------------------------------------
    def decode_packed(self, packed, n_full, sample_idx=None, maf_thr=0.02, miss_thr=0.05, het_thr=1.0,
                      genetic_model="add", want_g=True):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        rows, bps = packed.shape
        sidx = self._sample_idx(sample_idx)
        qc = QcCfg(maf_thr, miss_thr, het_thr, _model_code(genetic_model))
        counts = np.zeros((rows, 4), dtype=np.int32)
        af = np.zeros(rows, dtype=np.float32)
        mr = np.zeros(rows, dtype=np.float32)
        nk = C.c_size_t()
        g = np.zeros((rows, self.n), dtype=np.float32) if want_g else None
        check(lib().jxb_decode_packed(self.handle, ptr(packed), bps, rows, int(n_full), ptr(sidx), C.byref(qc),
                                      ptr(counts), ptr(af), ptr(mr), ptr(g), C.byref(nk)))
        return counts, af, mr, (g[: nk.value] if want_g else None)

    def _sample_idx(self, sample_idx):
        if sample_idx is None:
            return None
        sidx = np.ascontiguousarray(sample_idx, dtype=np.int64).reshape(-1)
        if sidx.shape[0] != self.n:      # the C ABI reads exactly n entries
            raise RuntimeError(f"sample_indices length mismatch: got {sidx.shape[0]}, expected {self.n}")
        return sidx

    def scan_packed(self, packed, n_full, sample_idx=None, pre_keep=None, maf_thr=0.02, miss_thr=0.05, het_thr=1.0,
                    genetic_model="add", mode="lmm", low=-5.0, high=5.0, max_iter=30, tol=1e-2, init=None,
                    nullml=None, log10_lbd=None, return_evals=False, row_af=None, row_flip=None):
        """One batch of packed SNP rows -> (keep, af, missing, out[n_kept, cols]).  row_af / row_flip: prepared row
        metadata (the caller's imputation frequency and LUT flip per row, src/decode/decode.rs:163-219)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8) if isinstance(packed, np.ndarray) else packed
        rows, bps = int(packed.shape[0]), int(packed.shape[1])
        mode_i = {"lmm": 0, "lmm2": 1, "fvlmm": 2}[mode]
        sidx = self._sample_idx(sample_idx)
        pk = None if pre_keep is None else np.ascontiguousarray(pre_keep, dtype=np.uint8)
        raf = None if row_af is None else np.ascontiguousarray(row_af, dtype=np.float32)
        rfl = None if row_flip is None else np.ascontiguousarray(np.asarray(row_flip, dtype=bool), dtype=np.uint8)
        if (raf is not None and raf.shape[0] != rows) or (rfl is not None and rfl.shape[0] != rows):
            raise RuntimeError("row_flip/row_maf length mismatch with packed rows")
        qc = QcCfg(maf_thr, miss_thr, het_thr, _model_code(genetic_model))
        if mode_i == 2:
            init = log10_lbd
        cfg = _solve_cfg(low, high, max_iter, tol, init, nullml)
        cols = 6 if mode_i == 1 else (4 if nullml is not None else 3)
        keep = np.zeros(rows, dtype=np.uint8)
        af = np.zeros(rows, dtype=np.float32)
        missing = np.zeros(rows, dtype=np.int32)
        out = np.zeros((rows, cols), dtype=np.float64)
        ev = np.zeros(rows, dtype=np.int32)
        nk = C.c_size_t()
        check(lib().jxb_scan_packed_prepared(self.handle, ptr(packed), bps, rows, int(n_full), ptr(sidx), ptr(pk),
                                             ptr(raf), ptr(rfl), C.byref(qc), C.byref(cfg), mode_i, ptr(keep), ptr(af),
                                             ptr(missing), ptr(out), ptr(ev), C.byref(nk)))
        res = (keep.astype(bool), af, missing, out[: nk.value])
        return res + (ev[: nk.value],) if return_evals else res

    def scan_packed_dev(self, packed_dev_ptr: int, rows: int, bps: int, n_full: int, sample_idx_dev_ptr=None,
                        maf_thr=0.02, miss_thr=0.05, het_thr=1.0, genetic_model="add", mode="lmm", low=-5.0,
                        high=5.0, max_iter=30, tol=1e-2, init=None, nullml=None, log10_lbd=None):
        """Asynchronous batch scan of packed rows already in HBM (results stay in the workspace)."""
        mode_i = {"lmm": 0, "lmm2": 1, "fvlmm": 2}[mode]
        qc = QcCfg(maf_thr, miss_thr, het_thr, _model_code(genetic_model))
        if mode_i == 2:
            init = log10_lbd
        cfg = _solve_cfg(low, high, max_iter, tol, init, nullml)
        check(lib().jxb_scan_packed_dev(self.handle, packed_dev_ptr, bps, rows, int(n_full), sample_idx_dev_ptr,
                                        C.byref(qc), C.byref(cfg), mode_i))
        return 6 if mode_i == 1 else (4 if nullml is not None else 3)

    def scan_fetch(self, rows: int, cols: int):
        keep = np.zeros(rows, dtype=np.uint8)
        af = np.zeros(rows, dtype=np.float32)
        missing = np.zeros(rows, dtype=np.int32)
        out = np.zeros((rows, cols), dtype=np.float64)
        ev = np.zeros(rows, dtype=np.int32)
        nk = C.c_size_t()
        check(lib().jxb_scan_fetch(self.handle, rows, cols, ptr(keep), ptr(af), ptr(missing), ptr(out), ptr(ev),
                                   C.byref(nk)))
        return keep.astype(bool), af, missing, out[: nk.value], ev[: nk.value]

    def scan_fetch_dev(self, rows: int, cols: int, out_dst_ptr=None, af_dst_ptr=None, counts_dst_ptr=None) -> int:
        """Copy the last scan's results device-to-device (out f64[n_kept, cols] compacted; af f32[rows]; counts
        i32[rows, 4] by source row) and return n_kept."""
        nk = C.c_size_t()
        check(lib().jxb_scan_fetch_dev(self.handle, rows, cols, out_dst_ptr, af_dst_ptr, counts_dst_ptr, C.byref(nk)))
        return int(nk.value)

    def stage_ms(self):
        """Device timers of the last scan (ms).  Streamed scans (`streamed` = 1): `rotate` covers all rotation slabs with
        the solve running underneath, `solve` is what was left of the solve afterwards, `solve_kernel` the solve
        kernel's own duration on its stream."""
        ms = (C.c_float * 8)()
        check(lib().jxb_last_stage_ms8(self.handle, ms))
        return dict(zip(("count_qc", "decode", "rotate", "solve", "h2d", "d2h", "solve_kernel", "streamed"),
                        [float(v) for v in ms]))

    # -- file level -------------------------------------------------------------------------------
    def scan_bed_to_tsv(self, bed_prefix, out_tsv, maf_thr, miss_thr, het_thr, genetic_model="add", snps_only=False,
                        sample_ids=None, mode="lmm", low=-5.0, high=5.0, max_iter=30, tol=1e-2, nullml=None,
                        init=None, log10_lbd=None, batch_rows=4096, progress_callback=None, progress_every=0,
                        snp_begin=0, snp_end=0, write_header=True, row_indices=None, row_maf=None, row_flip=None,
                        row_missing=None, mmap_window_mb=None) -> int:
        cfg = BedScanCfg()
        cfg.bed_prefix = str(bed_prefix).encode()
        cfg.out_tsv = str(out_tsv).encode()
        cfg.qc = QcCfg(maf_thr, miss_thr, het_thr, _model_code(genetic_model))
        mode_i = {"lmm": 0, "lmm2": 1, "fvlmm": 2}[mode]
        if mode_i == 2:
            init = log10_lbd
        cfg.solve = _solve_cfg(low, high, max_iter, tol, init, nullml)
        cfg.mode = mode_i
        cfg.snps_only = 1 if snps_only else 0
        keepalive = None
        if sample_ids is not None:
            ids = [str(x).encode() for x in sample_ids]
            keepalive = (C.c_char_p * len(ids))(*ids)
            cfg.sample_ids = keepalive
            cfg.n_sample_ids = len(ids)
        cfg.batch_rows = int(batch_rows)
        cfg.snp_begin, cfg.snp_end = int(snp_begin), int(snp_end)
        cfg.write_header = 1 if write_header else 0
        cfg.progress_every = int(progress_every)
        rows_keepalive = None
        if row_indices is not None:
            rows_keepalive = [np.ascontiguousarray(np.asarray(row_indices, dtype=np.int64))]
            cfg.row_indices = rows_keepalive[0].ctypes.data_as(C.POINTER(C.c_int64))
            cfg.n_row_indices = rows_keepalive[0].shape[0]
            for name, arr, ty, cty in (("row_maf", row_maf, np.float32, C.c_float), ("row_flip", row_flip, np.uint8, C.c_uint8),
                                       ("row_missing", row_missing, np.float32, C.c_float)):
                if arr is not None:
                    a = np.ascontiguousarray(np.asarray(arr).astype(ty, copy=False).reshape(-1))
                    if a.shape[0] != cfg.n_row_indices:
                        raise RuntimeError(f"prepared row metadata length mismatch: row_indices={cfg.n_row_indices}, {name}={a.shape[0]}")
                    rows_keepalive.append(a)
                    setattr(cfg, name, a.ctypes.data_as(C.POINTER(cty)))
        if mmap_window_mb:
            cfg.mmap_window_mb = max(1, int(mmap_window_mb))
        err = []

        def _cb(done, total, _user):
            if progress_callback is None:
                return 0
            try:
                progress_callback(int(done), int(total))
                return 0
            except BaseException as ex:  # KeyboardInterrupt included: abort the scan, re-raise after
                err.append(ex)
                return 1

        cb = _cabi.PROGRESS_CB(_cb)
        rows = C.c_size_t()
        rc = lib().jxb_scan_bed_to_tsv(self.handle, C.byref(cfg), C.byref(rows), cb, None)
        if err:
            raise err[0]
        check(rc)
        return int(rows.value)


# ------------------------------------------------------------------------------------------------------
# handle cache for the array-argument functions
# ------------------------------------------------------------------------------------------------------
_CACHE: "OrderedDict[tuple, DeviceModel]" = OrderedDict()
_CACHE_MAX = 2


def _digest(a: np.ndarray, sample: int = 1 << 16) -> bytes:
    flat = a.reshape(-1)
    if flat.shape[0] > sample:
        step = flat.shape[0] // sample
        flat = flat[::step]
    return hashlib.blake2b(np.ascontiguousarray(flat).tobytes(), digest_size=12).digest()


def _ut_key(ut: np.ndarray):
    """Cache key of a U^T argument: shape, dtype and a position-dependent checksum over EVERY entry (one threaded pass at
    memory bandwidth in libjxb200, ~30 ms for the 1.6 GB of n = 20,000), recomputed on every call -- so a matrix that
    differs anywhere from the resident one, including one modified in place, is never mistaken for it."""
    c = np.ascontiguousarray(ut)
    out = (C.c_uint64 * 2)()
    lib().jxb_host_checksum(c.ctypes.data, c.nbytes, out)
    return (ut.shape, str(ut.dtype), int(out[0]), int(out[1]))


def _evict_over_capacity() -> None:
    while len(_CACHE) > _CACHE_MAX:
        _, old = _CACHE.popitem(last=False)
        old.close()


def _get_model(s, xcov, y_rot, u_t=None) -> DeviceModel:
    s = _f64(s).reshape(-1)
    xcov = _f64(xcov)
    y = _f64(y_rot).reshape(-1)
    n = y.shape[0]
    if xcov.ndim != 2 or xcov.shape[0] != n:
        raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
    if s.shape[0] != n:
        raise RuntimeError("len(S) must equal len(y_rot)")
    ut = None
    if u_t is not None:
        ut = np.asarray(u_t)
        if ut.ndim != 2 or ut.shape != (n, n):
            raise RuntimeError("u_t must be (n, n) and row-major U^T")
    ut_key = None if ut is None else _ut_key(ut)
    key_small = (n, xcov.shape[1], _digest(s, 1 << 30), _digest(xcov, 1 << 30), _digest(y, 1 << 30))
    for key, mdl in list(_CACHE.items()):
        if key[0] == key_small and (ut_key is None or key[1] == ut_key):
            _CACHE.move_to_end(key)
            return mdl
    # same U^T, new trait/covariates: keep the resident U^T and swap the small arrays
    if ut_key is not None:
        for key, mdl in list(_CACHE.items()):
            if key[1] == ut_key and key[0][0] == n and key[0][1] == xcov.shape[1] and key[0][2] == key_small[2]:
                del _CACHE[key]
                mdl.set_xy(xcov, y)
                _CACHE[(key_small, ut_key)] = mdl
                return mdl
        # a rotate-only placeholder (lmm_rotate_x_y_with_ut_f64) holding the same U^T: release it first so the matrix
        # and its digit planes are never resident twice (~37 GB each at n = 50,000)
        for key, mdl in list(_CACHE.items()):
            if key[1] == ut_key and key[0][2:4] == (b"rot", b"rot"):
                del _CACHE[key]
                mdl.close()
    mdl = DeviceModel(s, xcov, y, ut)
    _CACHE[(key_small, ut_key)] = mdl
    _evict_over_capacity()
    return mdl


def set_rotate_variant(variant: int) -> None:
    """Rotation kernel used by the packed scans (default 3): 0 = FP64 DMMA/TMA GEMM, 1 = CUDA-core cross-check,
    2 = exact int8-sliced tensor-core rotation, slice GEMMs through cuBLASLt (additive coding; csrc/k2_int8.cu),
    3 = the same arithmetic on the hand-written tcgen05/TMEM kernel (csrc/k2_i8mma.cu)."""
    lib().jxb_set_rotate_variant(int(variant))


def set_thread_solve_min_rows(rows: int) -> None:
    """Batches with at least `rows` kept SNPs use the one-thread-per-SNP solve kernel (default 32768)."""
    lib().jxb_set_thread_solve_min_rows(int(rows))


def set_big_solve_kernel(variant: int) -> None:
    """Large-batch solve kernel: 0 = lane-per-SNP with refill (default), 1 = thread-per-SNP on an SNP-minor block."""
    lib().jxb_set_big_solve_kernel(int(variant))


def set_prefix_evals(mode: int) -> None:
    """Shared-abscissa prefix of the per-SNP REML searches in the lane-per-SNP solve: 1 = batches of >= 2048 kept SNPs
    (default), 2 = every batch, 0 = off.  Results and evaluation counts do not depend on it."""
    lib().jxb_set_prefix_evals(int(mode))


def set_stream_overlap(on: bool, slab_rows: int = 0) -> None:
    """Streamed scan (default on): rotate large batches in slabs while one persistent solve kernel consumes the rotated
    rows, so the tensor pipe and the FP64 pipe run concurrently.  slab_rows = 0 keeps the current slab size (8192)."""
    lib().jxb_set_stream_overlap(int(on) if on in (0, 1, 2) else (1 if on else 0), int(slab_rows))


def clear_model_cache() -> None:
    while _CACHE:
        _, m = _CACHE.popitem()
        m.close()


# ------------------------------------------------------------------------------------------------------
# PyO3-compatible functions
# ------------------------------------------------------------------------------------------------------
def lmm_reml_chunk_f32(s, xcov, y_rot, low, high, g_rot_chunk, max_iter=50, tol=1e-2, threads=0, nullml=None):
    """src/stats/lmm.rs:333-518 -> f64[m, 3|4] = beta, se, pwald[, plrt]."""
    return _get_model(s, xcov, y_rot).lmm_reml_chunk(g_rot_chunk, low, high, max_iter, tol, nullml, rotated=True)


def lmm_reml_chunk_from_snp_f32(s, xcov, y_rot, low, high, snp_chunk, u_t, max_iter=50, tol=1e-2, threads=0,
                                nullml=None, rotate_block_rows=256):
    """src/stats/lmm.rs:1479-1630."""
    return _get_model(s, xcov, y_rot, u_t).lmm_reml_chunk(snp_chunk, low, high, max_iter, tol, nullml, rotated=False)


def lmm_reml_lmm2_chunk_from_snp_f32(s, xcov, y_rot, low, high, snp_chunk, u_t, nullml, max_iter=50, tol=1e-2,
                                     threads=0, rotate_block_rows=256):
    """src/stats/lmm.rs:1632-1780 -> f64[m, 6] = beta, se, pwald, lambda_reml, ml_alt, plrt."""
    return _get_model(s, xcov, y_rot, u_t).lmm2_chunk(snp_chunk, low, high, nullml, max_iter, tol, rotated=False)


def lmm_assoc_chunk_f32(s, xcov, y_rot, log10_lbd, g_rot_chunk, threads=0, nullml=None):
    """Fixed-lambda scan of a rotated block: src/stats/lmm.rs:2010-2223 / src/stats/fvlmm.rs:1691-1805."""
    return _get_model(s, xcov, y_rot).fixed_chunk(g_rot_chunk, log10_lbd, nullml, rotated=True)


def lmm_assoc_chunk_from_snp_f32(s, xcov, y_rot, log10_lbd, snp_chunk, u_t, threads=0, nullml=None,
                                 rotate_block_rows=512):
    """src/stats/lmm.rs:2225-2238 (fvlmm_assoc_chunk_from_snp_f32, src/stats/fvlmm.rs:2114-2116, has the same signature;
    the reference's Python passes all nine arguments positionally, pyBLUP/assoc.py:1512-1522)."""
    return _get_model(s, xcov, y_rot, u_t).fixed_chunk(snp_chunk, log10_lbd, nullml, rotated=False)


fvlmm_assoc_chunk_f32 = lmm_assoc_chunk_f32
fvlmm_assoc_chunk_from_snp_f32 = lmm_assoc_chunk_from_snp_f32


def lmm_reml_null_f32(s, xcov, y_rot, low, high, max_iter=50, tol=1e-2):
    """src/stats/reml.rs:570-616 -> (lambda, ml, reml)."""
    if low >= high:
        raise RuntimeError("low must be < high")
    return _get_model(s, xcov, y_rot).reml_null(low, high, max_iter, tol)


def ml_loglike_null_f32(s, xcov, y_rot, log10_lbd):
    """src/stats/reml.rs:618-646."""
    return _get_model(s, xcov, y_rot).ml_loglike_null(log10_lbd)


def lmm_rotate_x_y_with_ut_f64(u_t, x, y, threads=0):
    """src/stats/reml.rs:107-198 -> (f64[n, q], f64[n, 1])."""
    y = _f64(y).reshape(-1)
    n = y.shape[0]
    if n == 0:
        raise RuntimeError("y must not be empty")
    x = _f64(x)
    if x.ndim != 2:
        raise RuntimeError("x must be 2D (n, q)")
    if x.shape[0] != n:
        raise RuntimeError(f"x rows must equal len(y): rows={x.shape[0]}, len(y)={n}")
    ut = np.asarray(u_t)
    if ut.ndim != 2:
        raise RuntimeError("u_t must be 2D (n, n)")
    if ut.shape != (n, n):
        raise RuntimeError("u_t must be shape (n, n) and row-major U^T")
    # the model only needs U^T here; S / Xcov / y are placeholders of the right shape
    ut_key = _ut_key(ut)
    for key, mdl in _CACHE.items():
        if key[1] == ut_key:
            return mdl.rotate_xy(x, y)
    mdl = DeviceModel(np.ones(n), np.ones((n, 1)), np.zeros(n), ut)
    _CACHE[((n, 1, b"rot", b"rot", _digest(y, 1 << 30)), ut_key)] = mdl
    while len(_CACHE) > _CACHE_MAX:
        _, old = _CACHE.popitem(last=False)
        old.close()
    return mdl.rotate_xy(x, y)


def _check_bed_args(s, xcov, y_rot, u_t, low, high, tol):
    if low >= high:
        raise RuntimeError("low must be < high")
    if not (math.isfinite(tol) and tol > 0.0):
        raise RuntimeError("tol must be positive and finite")
    y = np.asarray(y_rot).reshape(-1)
    n = y.shape[0]
    xc = np.asarray(xcov)
    if xc.shape[0] != n:
        raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
    if np.asarray(s).reshape(-1).shape[0] != n:
        raise RuntimeError("len(S) must equal len(y_rot)")
    if tuple(np.asarray(u_t).shape) != (n, n):
        raise RuntimeError("u_t must be (n, n) row-major U^T")
    if n <= xc.shape[1] + 1:
        raise RuntimeError("n must be > p+1")


def _prepared_rows(row_indices, row_flip, row_missing, row_maf):
    """Prepared row metadata (src/stats/lmm.rs:2576-2612): all four or none.  The listed rows are scanned without
    re-applying QC; `row_maf` is the imputation frequency (and the TSV af column), `row_flip` reverses the code LUT,
    `row_missing` (a rate) becomes the TSV miss column through round(rate * n) -- all as the reference uses them."""
    given = [v is not None for v in (row_indices, row_flip, row_missing, row_maf)]
    if any(given) and not all(given):
        raise RuntimeError(
            "prepared row metadata must provide all or none of: row_indices, row_flip, row_missing, row_maf")
    if not all(given):
        return None
    idx = np.ascontiguousarray(np.asarray(row_indices, dtype=np.int64).reshape(-1))
    m = idx.shape[0]
    if not (np.asarray(row_flip).shape[0] == m and np.asarray(row_missing).shape[0] == m
            and np.asarray(row_maf).shape[0] == m):
        raise RuntimeError(
            f"prepared row metadata length mismatch: row_indices={m}, row_flip={np.asarray(row_flip).shape[0]}, "
            f"row_maf={np.asarray(row_maf).shape[0]}, row_missing={np.asarray(row_missing).shape[0]}")
    if m > 1 and np.any(np.diff(idx) < 0):
        raise RuntimeError("prepared row_indices must be sorted in ascending BED order")
    return dict(row_indices=idx, row_maf=np.asarray(row_maf, dtype=np.float32).reshape(-1),
                row_flip=np.asarray(row_flip, dtype=bool).reshape(-1).astype(np.uint8),
                row_missing=np.asarray(row_missing, dtype=np.float32).reshape(-1))


def lmm_reml_assoc_bed_to_tsv_f32(bed_prefix, out_tsv, s, xcov, y_rot, u_t, maf_thr, miss_thr, het_thr,
                                  genetic_model="add", snps_only=False, sample_ids=None, row_indices=None,
                                  row_flip=None, row_missing=None, row_maf=None, low=-5.0, high=5.0, max_iter=30,
                                  tol=1e-2, threads=0, nullml=None, init_log10_lbd=None, rotate_block_rows=512,
                                  progress_callback=None, progress_every=0, mmap_window_mb=None) -> int:
    """src/stats/lmm.rs:2488-2750.  `init_log10_lbd` (clamped into [low, high], lmm.rs:2573-2575) seeds the Brent search
    of EVERY SNP.  The reference additionally carries each rayon worker's last optimum to its next SNP
    (use_warm_start, lmm.rs:134-161), which makes its rows depend on the thread schedule; that carry-over is never
    used here (JX_LMM_UNIFIED_NO_WARM_START=1 semantics), so rows are independent of batch size and GPU count.  With
    tol = 1e-2 a different Brent start moves lambda within the optimiser's tolerance: against a default-mode reference
    run expect |d log10 lambda| <~ 1e-2 and beta/se agreement to ~1e-4 relative, not the 1e-8 of the no-warm-start mode."""
    _check_bed_args(s, xcov, y_rot, u_t, low, high, tol)
    _model_code(genetic_model)
    rows_sel = _prepared_rows(row_indices, row_flip, row_missing, row_maf)
    mdl = _get_model(s, xcov, y_rot, u_t)
    init = None
    if init_log10_lbd is not None and math.isfinite(init_log10_lbd):
        init = min(max(float(init_log10_lbd), low), high)
    return mdl.scan_bed_to_tsv(bed_prefix, out_tsv, maf_thr, miss_thr, het_thr, genetic_model, snps_only, sample_ids,
                               "lmm", low, high, max_iter, tol, nullml, init, None, mmap_window_mb=mmap_window_mb, **(rows_sel or {}),
                               batch_rows=max(int(rotate_block_rows), default_device_batch(mdl.n)), progress_callback=progress_callback,
                               progress_every=progress_every)


def lmm_reml_lmm2_assoc_bed_to_tsv_f32(bed_prefix, out_tsv, s, xcov, y_rot, u_t, maf_thr, miss_thr, het_thr,
                                       genetic_model="add", snps_only=False, sample_ids=None, row_indices=None,
                                       row_flip=None, row_missing=None, row_maf=None, low=-5.0, high=5.0,
                                       max_iter=30, tol=1e-2, threads=0, nullml=None, init_log10_lbd_reml=None,
                                       init_log10_lbd_ml=None, rotate_block_rows=512, progress_callback=None,
                                       progress_every=0, mmap_window_mb=None) -> int:
    """src/stats/lmm.rs:2753-3038 (REML start = init_reml.or(init_ml); ML start = per-SNP REML optimum)."""
    _check_bed_args(s, xcov, y_rot, u_t, low, high, tol)
    _model_code(genetic_model)
    rows_sel = _prepared_rows(row_indices, row_flip, row_missing, row_maf)
    if nullml is not None and not math.isfinite(nullml):
        raise RuntimeError("nullml must be finite when provided")
    mdl = _get_model(s, xcov, y_rot, u_t)

    def _clamped(v):
        if v is None or not math.isfinite(v):
            return None
        return min(max(float(v), low), high)

    i_reml, i_ml = _clamped(init_log10_lbd_reml), _clamped(init_log10_lbd_ml)
    if nullml is None:
        _, nullml = mdl.ml_null(low, high, max_iter, tol, i_ml if i_ml is not None else i_reml)
        if not math.isfinite(nullml):
            raise RuntimeError("failed to optimize null ML for LMM2 unified scan")
    init = i_reml if i_reml is not None else i_ml
    return mdl.scan_bed_to_tsv(bed_prefix, out_tsv, maf_thr, miss_thr, het_thr, genetic_model, snps_only, sample_ids,
                               "lmm2", low, high, max_iter, tol, nullml, init, None, mmap_window_mb=mmap_window_mb, **(rows_sel or {}),
                               batch_rows=max(int(rotate_block_rows), default_device_batch(mdl.n)), progress_callback=progress_callback,
                               progress_every=progress_every)


def fvlmm_assoc_bed_to_tsv_f32(bed_prefix, out_tsv, s, xcov, y_rot, log10_lbd, u_t, maf_thr, miss_thr, het_thr,
                               genetic_model="add", snps_only=False, sample_ids=None, row_indices=None, row_flip=None,
                               row_missing=None, row_maf=None, threads=0, nullml=None, rotate_block_rows=512,
                               progress_callback=None, progress_every=0, mmap_window_mb=None):
    """src/stats/fvlmm.rs:2482-3202 -> (rows_written, pve, log_det_v)."""
    _check_bed_args(s, xcov, y_rot, u_t, -1.0, 1.0, 1e-2)
    lbd = 10.0 ** float(log10_lbd)
    if not (math.isfinite(lbd) and lbd > 0.0):
        raise RuntimeError("invalid log10_lbd")
    _model_code(genetic_model)
    rows_sel = _prepared_rows(row_indices, row_flip, row_missing, row_maf)
    mdl = _get_model(s, xcov, y_rot, u_t)
    rows = mdl.scan_bed_to_tsv(bed_prefix, out_tsv, maf_thr, miss_thr, het_thr, genetic_model, snps_only, sample_ids,
                               "fvlmm", nullml=nullml, log10_lbd=log10_lbd, mmap_window_mb=mmap_window_mb, **(rows_sel or {}),
                               batch_rows=max(int(rotate_block_rows), default_device_batch(mdl.n)), progress_callback=progress_callback,
                               progress_every=progress_every)
    # fvlmm.rs:2746-2753: pve = clamp(1 - ypy / sum y^2, 0, 1)
    _, meta = mdl.fixed_chunk(np.zeros((1, mdl.n), dtype=np.float32), log10_lbd, rotated=True, return_meta=True)
    y_sq = float(np.sum(mdl.y * mdl.y))
    pve = float(min(max(1.0 - meta["ypy"] / y_sq, 0.0), 1.0)) if y_sq > 0.0 else float("nan")
    return rows, pve, float(meta["log_det_v"])


# ------------------------------------------------------------------------------------------------------
# TSV writer + packed-array scans (route "B" callers that already hold packed rows / result blocks)
# ------------------------------------------------------------------------------------------------------
def _blob(strings) -> bytes:
    return b"\0".join(str(x).encode() for x in strings) + b"\0"


def _format_block(chrom, pos, snp, a0, a1, af, miss_rate, results, genetic_model="add") -> bytes:
    res = _f64(results)
    rows, cols = res.shape
    if rows == 0:
        return b""
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    af = np.ascontiguousarray(af, dtype=np.float32)
    mr = np.ascontiguousarray(miss_rate, dtype=np.float32)
    blobs = [_blob(v) for v in (chrom, snp, a0, a1)]
    cap = 160 * rows + sum(len(b) for b in blobs) * 3
    for _ in range(2):
        buf = C.create_string_buffer(cap)
        need = lib().jxb_format_block(buf, cap, rows, blobs[0], ptr(pos), blobs[1], blobs[2], blobs[3], ptr(af), ptr(mr),
                                      ptr(res), cols, _model_code(genetic_model))
        if need == 0:
            raise ValueError(f"results must have 3, 4 or 6 columns, got {cols}")
        if need <= cap:
            return buf.raw[:need]
        cap = need
    raise RuntimeError("TSV block formatting failed")


class SiteInfo:
    """chrom / pos / ref_allele / alt_allele record (PySiteInfo in the reference)."""
    __slots__ = ("chrom", "pos", "ref_allele", "alt_allele")

    def __init__(self, chrom, pos, ref_allele, alt_allele):
        self.chrom, self.pos, self.ref_allele, self.alt_allele = str(chrom), int(pos), str(ref_allele), str(alt_allele)


class GwasAssocTsvWriter:
    """src/io/assoc2tsv.rs:765-892: association TSV sink; the header is written when the first block fixes the
    schema (3 / 4 / 6 result columns); rows are formatted by the library (jxb_format_block)."""

    def __init__(self, path, genetic_model="add"):
        key = str(genetic_model).strip().lower()
        if key not in ("add", "dom", "rec", "het"):
            raise ValueError("genetic_model must be one of: add, dom, rec, het")
        self._path, self._model, self._cols, self._fh, self._rows, self._closed = str(path), key, None, None, 0, False

    def _ensure(self, cols):
        if cols not in (3, 4, 6):
            raise ValueError(f"unsupported results column count: {cols}")
        if self._closed:
            raise IOError("writer is closed")
        if self._cols is None:
            self._cols = cols
            self._fh = open(self._path, "wb", buffering=8 << 20)
            self._fh.write(lib().jxb_tsv_header(cols))
        elif self._cols != cols:
            raise ValueError(f"inconsistent results columns across chunks: expected {self._cols}, got {cols}")

    def write_chunk(self, sites, snp, maf, miss, results) -> int:
        sites = list(sites)
        if not sites:
            return 0
        snp = list(snp)
        if len(snp) != len(sites):
            raise ValueError(f"snp length mismatch: snp={len(snp)}, sites={len(sites)}")
        maf, miss = np.asarray(maf), np.asarray(miss)
        if maf.shape[0] != len(sites):
            raise ValueError(f"maf length mismatch: maf={maf.shape[0]}, sites={len(sites)}")
        if miss.shape[0] != len(sites):
            raise ValueError(f"miss length mismatch: miss={miss.shape[0]}, sites={len(sites)}")
        res = np.asarray(results)
        if res.ndim != 2:
            raise ValueError("results must be 2D")
        if res.shape[0] != len(sites):
            raise ValueError(f"results row mismatch: results={res.shape[0]}, sites={len(sites)}")
        if res.shape[1] < 3:
            raise ValueError(f"results must have at least 3 columns, got {res.shape[1]}")
        self._ensure(res.shape[1])
        get = (lambda s, k, i: getattr(s, k)) if hasattr(sites[0], "chrom") else (lambda s, k, i: s[i])
        text = _format_block([get(s, "chrom", 0) for s in sites], [int(get(s, "pos", 1)) for s in sites], snp,
                             [get(s, "ref_allele", 2) for s in sites], [get(s, "alt_allele", 3) for s in sites],
                             maf, miss, res, self._model)
        self._fh.write(text)
        self._rows += len(sites)
        return len(sites)

    def append_text(self, text, has_plrt, rows):
        if rows == 0 or not text:
            return
        self._ensure(4 if has_plrt else 3)
        self._fh.write(text.encode() if isinstance(text, str) else bytes(text))
        self._rows += int(rows)

    def send_block(self, data):
        if not data:
            return
        if self._fh is None or self._closed:
            raise IOError("writer is closed")
        self._fh.write(bytes(data))

    def flush(self):
        if self._fh is not None and not self._closed:
            self._fh.flush()

    def close(self):
        if self._fh is not None and not self._closed:
            self._fh.close()
        self._closed = True

    @property
    def rows_written(self) -> int:
        return self._rows

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _packed_args(packed, n_samples, row_flip, row_maf, s, xcov, y_rot, u_t, sample_indices, row_indices, low, high, tol):
    """Validation of src/stats/lmm.rs:3089-3187 (same messages); returns (packed rows to scan, sample index)."""
    if n_samples == 0:
        raise RuntimeError("n_samples must be > 0")
    if low >= high:
        raise RuntimeError("low must be < high")
    if not (math.isfinite(tol) and tol > 0.0):
        raise RuntimeError("tol must be positive and finite")
    packed = np.asarray(packed)
    if packed.ndim != 2:
        raise RuntimeError("packed must be 2D (m, bytes_per_snp)")
    m_packed, bps = packed.shape
    if bps != (n_samples + 3) // 4:
        raise RuntimeError(f"packed second dimension mismatch: got {bps}, expected {(n_samples + 3) // 4}")
    if row_indices is not None:
        ridx = np.asarray(row_indices, dtype=np.int64).reshape(-1)
        if ridx.size and (ridx.min() < 0 or ridx.max() >= m_packed):
            raise RuntimeError("row_indices out of range")
        packed = packed[ridx]
    m = packed.shape[0]
    if np.asarray(row_flip).shape[0] != m or np.asarray(row_maf).shape[0] != m:
        raise RuntimeError("row_flip/row_maf length mismatch with packed rows")
    n = np.asarray(y_rot).reshape(-1).shape[0]
    if n == 0:
        raise RuntimeError("y_rot must not be empty")
    sidx = None
    if sample_indices is not None:
        sidx = np.asarray(sample_indices, dtype=np.int64).reshape(-1)
        if sidx.shape[0] != n:
            raise RuntimeError(f"sample_indices length mismatch: got {sidx.shape[0]}, expected {n}")
        if sidx.size and (sidx.min() < 0 or sidx.max() >= n_samples):
            raise RuntimeError("sample_indices out of range")
    elif n != n_samples:
        raise RuntimeError(f"len(y_rot)={n} must equal n_samples={n_samples} when sample_indices is not provided")
    _check_bed_args(s, xcov, y_rot, u_t, low, high, tol)
    return np.ascontiguousarray(packed, dtype=np.uint8), sidx


def _scan_packed_all_rows(mdl, packed, n_samples, sidx, row_maf, row_flip, model, low, high, max_iter, tol, init, nullml,
                          progress_callback, progress_every):
    """Every supplied row is scanned (thresholds off: the caller already filtered).  `row_maf` is the imputation
    frequency exactly as the reference uses it (mean_g = 2 * row_maf, src/decode/decode.rs:213-219) -- it need not be
    the frequency over the selected samples (workflow_model_packed.py:1296 passes full-sample values) -- and `row_flip`
    reverses the code LUT."""
    m = packed.shape[0]
    cols = 4 if nullml is not None else 3
    out = np.zeros((m, cols), dtype=np.float64)
    miss_all = np.zeros(m, dtype=np.int32)
    raf = np.ascontiguousarray(np.asarray(row_maf, dtype=np.float32).reshape(-1))
    rfl = np.ascontiguousarray(np.asarray(row_flip, dtype=bool).reshape(-1))
    step = default_device_batch(mdl.n) if not progress_every else max(1, min(int(progress_every), default_device_batch(mdl.n)))
    for r0 in range(0, m, step):
        r1 = min(m, r0 + step)
        keep, af, missing, res = mdl.scan_packed(packed[r0:r1], n_samples, sidx, None, 0.0, 1.0, 0.0, model, "lmm", low, high,
                                                 max_iter, tol, init, nullml, row_af=raf[r0:r1],
                                                 row_flip=(rfl[r0:r1] if rfl.any() else None))
        if not keep.all():
            raise RuntimeError("internal error: packed scan dropped rows with QC disabled")
        out[r0:r1], miss_all[r0:r1] = res, missing
        if progress_callback is not None:
            progress_callback(r1, m)
    return out, miss_all


def lmm_reml_assoc_packed_f32(packed, n_samples, row_flip, row_maf, s, xcov, y_rot, u_t, sample_indices=None,
                              row_indices=None, low=-5.0, high=5.0, max_iter=50, tol=1e-2, threads=0, model="add",
                              progress_callback=None, progress_every=0, nullml=None, init_log10_lbd=None,
                              rotate_block_rows=256):
    """src/stats/lmm.rs:3040-3362 -> f64[m, 3|4].  `init_log10_lbd` seeds every SNP (no carried warm start)."""
    _model_code(model)
    packed, sidx = _packed_args(packed, n_samples, row_flip, row_maf, s, xcov, y_rot, u_t, sample_indices, row_indices,
                                low, high, tol)
    init = None
    if init_log10_lbd is not None and math.isfinite(init_log10_lbd):
        init = min(max(float(init_log10_lbd), low), high)
    mdl = _get_model(s, xcov, y_rot, u_t)
    out, _ = _scan_packed_all_rows(mdl, packed, n_samples, sidx, row_maf, row_flip, model, low, high, max_iter, tol, init,
                                   nullml, progress_callback, progress_every)
    return out


def _read_bim_columns(prefix, row_indices):
    chrom, pos, snp, a0, a1 = [], [], [], [], []
    with open(str(prefix) + ".bim") as fh:
        for ln, line in enumerate(fh, 1):
            tok = line.split()
            if len(tok) < 6:
                raise RuntimeError(f"Malformed BIM line at {prefix}.bim:{ln}: {line.rstrip()}")
            chrom.append(tok[0]); snp.append(tok[1]); a0.append(tok[4]); a1.append(tok[5])
            try:
                v = int(tok[3])
                pos.append(v if -(1 << 31) <= v < (1 << 31) else 0)
            except ValueError:
                pos.append(0)
    cols = (chrom, pos, snp, a0, a1)
    if row_indices is not None:
        cols = tuple([c[int(i)] for i in row_indices] for c in cols)
    return cols


def lmm_reml_assoc_packed_f32_to_tsv(packed, n_samples, row_flip, row_maf, row_missing, s, xcov, y_rot, u_t, chrom, pos,
                                     snp, allele0, allele1, out_tsv, sample_indices=None, row_indices=None, low=-5.0,
                                     high=5.0, max_iter=50, tol=1e-2, threads=0, model="add", progress_callback=None,
                                     progress_every=0, nullml=None, init_log10_lbd=None, rotate_block_rows=256,
                                     bed_prefix=None) -> int:
    """src/stats/lmm.rs:3364-3800 -> rows written.  af column = `row_maf`; miss column = the count recovered from
    `row_missing` (round(rate * n)) divided by n again, in f32 (lmm.rs:1934-1950, 3704-3707)."""
    _model_code(model)
    packed, sidx = _packed_args(packed, n_samples, row_flip, row_maf, s, xcov, y_rot, u_t, sample_indices, row_indices,
                                low, high, tol)
    m = packed.shape[0]
    if np.asarray(row_missing).shape[0] != m:
        raise RuntimeError("row_flip/row_maf/row_missing length mismatch with packed rows")
    cols_meta = [list(chrom), list(pos), list(snp), list(allele0), list(allele1)]
    if all(len(c) == 0 for c in cols_meta):
        if bed_prefix is None or not str(bed_prefix).strip():
            raise RuntimeError("empty TSV metadata requires non-empty bed_prefix")
        cols_meta = list(_read_bim_columns(str(bed_prefix).strip(), row_indices))
        if any(len(c) != m for c in cols_meta):
            raise RuntimeError(f"BIM metadata length mismatch: expected={m}")
    elif any(len(c) != m for c in cols_meta):
        raise RuntimeError(f"TSV metadata length mismatch: rows={m}, chrom={len(cols_meta[0])}, pos={len(cols_meta[1])}, "
                           f"snp={len(cols_meta[2])}, allele0={len(cols_meta[3])}, allele1={len(cols_meta[4])}")
    init = None
    if init_log10_lbd is not None and math.isfinite(init_log10_lbd):
        init = min(max(float(init_log10_lbd), low), high)
    mdl = _get_model(s, xcov, y_rot, u_t)
    out, _ = _scan_packed_all_rows(mdl, packed, n_samples, sidx, row_maf, row_flip, model, low, high, max_iter, tol, init,
                                   nullml, progress_callback, progress_every)
    n = mdl.n
    rate = np.asarray(row_missing, dtype=np.float32)
    cnt = np.where(np.isfinite(rate) & (rate > 0), np.round(rate.astype(np.float64) * n), 0.0)
    miss_rate = cnt.astype(np.float32) / np.float32(n)
    text = _format_block(*cols_meta, np.asarray(row_maf, dtype=np.float32), miss_rate, out, "add")
    with open(out_tsv, "wb") as fh:
        fh.write(lib().jxb_tsv_header(out.shape[1]))
        fh.write(text)
    return m


class FvLmmAssocCache:
    """src/stats/fvlmm.rs:1412-1436: handle of the fixed-lambda precomputation (weights, Cholesky of X'WX, ypy).
    Here the cached state lives with the device model; the object pins (model, log10 lambda)."""

    def __init__(self, mdl: DeviceModel, log10_lbd: float, arrays=None):
        self._mdl, self._l10 = mdl, float(log10_lbd)
        self._arrays = arrays                     # (s, xcov, y_rot) for the from-SNP variant, which also needs U^T
        self.n, self.p, self.lbd = mdl.n, mdl.p, 10.0 ** float(log10_lbd)


def fvlmm_assoc_prepare_cache_f32(s, xcov, y_rot, log10_lbd) -> FvLmmAssocCache:
    """src/stats/fvlmm.rs:1807-1846."""
    y = _f64(y_rot).reshape(-1)
    xc = _f64(xcov)
    if xc.ndim != 2 or xc.shape[0] != y.shape[0]:
        raise RuntimeError("Xcov.n_rows must equal len(y_rot)")
    if _f64(s).reshape(-1).shape[0] != y.shape[0]:
        raise RuntimeError("len(S) must equal len(y_rot)")
    if y.shape[0] <= xc.shape[1] + 1:
        raise RuntimeError("n must be > p_cov+1")
    lbd = 10.0 ** float(log10_lbd)
    if not (math.isfinite(lbd) and lbd > 0.0):
        raise RuntimeError("invalid log10_lbd")
    # a private device model: the handle must outlive evictions from the array-argument cache
    return FvLmmAssocCache(DeviceModel(s, xc, y), log10_lbd, (_f64(s).reshape(-1), xc, y))


def fvlmm_assoc_chunk_with_cache_f32(cache: FvLmmAssocCache, g_rot_chunk, threads=0, nullml=None):
    """src/stats/fvlmm.rs:1917-1937."""
    g = np.asarray(g_rot_chunk)
    if g.ndim != 2 or g.shape[1] != cache.n:
        raise RuntimeError("g_rot_chunk must be (m, n)")
    return cache._mdl.fixed_chunk(g, cache._l10, nullml, rotated=True)


def fvlmm_assoc_chunk_from_snp_with_cache_f32(cache: FvLmmAssocCache, snp_chunk, u_t, threads=0, nullml=None,
                                              rotate_block_rows=512):
    """src/stats/fvlmm.rs:1997-2060: the cached fixed-lambda scan of an unrotated chunk."""
    g = np.asarray(snp_chunk)
    if g.ndim != 2 or g.shape[1] != cache.n:
        raise RuntimeError("snp_chunk must be (m, n)")
    s_, xc_, y_ = cache._arrays
    return _get_model(s_, xc_, y_, u_t).fixed_chunk(g, cache._l10, nullml, rotated=False)


def fvlmm_assoc_chunk_from_snp_to_tsv_f32(s, xcov, y_rot, log10_lbd, snp_chunk, u_t, chrom, pos, snp, allele0, allele1, maf,
                                          miss, threads=0, nullml=None, rotate_block_rows=512, progress_callback=None,
                                          progress_every=0):
    """src/stats/fvlmm.rs:2257-2480 -> ([tsv text blocks], rows): fixed-lambda scan of an unrotated chunk returned as
    ready-made TSV rows (miss is a rate)."""
    g = np.asarray(snp_chunk)
    m = g.shape[0]
    if any(len(c) != m for c in (chrom, pos, snp, allele0, allele1, maf, miss)):
        raise RuntimeError(f"TSV metadata length mismatch: rows={m}")
    res = lmm_assoc_chunk_from_snp_f32(s, xcov, y_rot, log10_lbd, g, u_t, threads, nullml)
    if progress_callback is not None:
        progress_callback(m, m)
    if m == 0:
        return [], 0
    return [_format_block(chrom, pos, snp, allele0, allele1, maf, miss, res, "add")], m


def _outside_scope(name):
    def _f(*_a, **_k):
        raise NotImplementedError(f"{name} belongs to a different algorithm (FaST-LMM low-rank / plain LM) and is outside "
                                  "the exact-LMM path this library implements")
    _f.__name__ = name
    _f.__doc__ = "Present so that `from janusx.janusx import ...` in python/janusx/pyBLUP/assoc.py:207-218 resolves."
    return _f


# hard imports of the reference's Python layer that are not part of this path (pyBLUP/assoc.py:207-218)
fastlmm_prepare_lowrank_f64 = _outside_scope("fastlmm_prepare_lowrank_f64")
fastlmm_assoc_from_snp_f32 = _outside_scope("fastlmm_assoc_from_snp_f32")
fastlmm_reml_chunk_f32 = _outside_scope("fastlmm_reml_chunk_f32")
fastlmm_reml_null_f32 = _outside_scope("fastlmm_reml_null_f32")
fastlmm_assoc_chunk_f32 = _outside_scope("fastlmm_assoc_chunk_f32")
lm_block_assoc_f32 = _outside_scope("lm_block_assoc_f32")


# ------------------------------------------------------------------------------------------------------
# GRM and its eigendecomposition on the device
# ------------------------------------------------------------------------------------------------------
class DeviceGrm:
    """Streaming centred-additive GRM accumulator in HBM (csrc/grm.cu; src/stats/grm.rs:204-608)."""

    def __init__(self, n_samples: int, sample_indices=None, method: int = 1, device: int = 0):
        _cabi.require_gpu()
        if n_samples == 0:
            raise RuntimeError("n_samples must be > 0")
        self._sidx = None if sample_indices is None else np.ascontiguousarray(sample_indices, dtype=np.int64).reshape(-1)
        self.n_full = int(n_samples)
        self.n = self.n_full if self._sidx is None else int(self._sidx.shape[0])
        self.device = int(device)
        h = C.c_void_p()
        check(lib().jxb_grm_create(self.device, self.n_full, ptr(self._sidx), 0 if self._sidx is None else self.n,
                                   int(method), C.byref(h)))
        self._h = h

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("GRM handle is closed")
        return self._h

    def update(self, packed, row_maf=None, qc=None) -> None:
        """Add packed SNP rows.  row_maf: the prepared allele frequencies (reference semantics); None => computed on
        the device, optionally with `qc=(maf_thr, miss_thr, het_thr)` leaving failing rows out."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if packed.ndim != 2:
            raise RuntimeError("packed must be 2D (m, bytes_per_snp)")
        maf = None if row_maf is None else np.ascontiguousarray(row_maf, dtype=np.float32).reshape(-1)
        if maf is not None and maf.shape[0] != packed.shape[0]:
            raise RuntimeError(f"row_maf length mismatch: got {maf.shape[0]}, expected {packed.shape[0]}")
        cfg = None if qc is None else QcCfg(float(qc[0]), float(qc[1]), float(qc[2]), 0)
        check(lib().jxb_grm_update(self.handle, ptr(packed), packed.shape[1], packed.shape[0], ptr(maf),
                                   C.byref(cfg) if cfg is not None else None))

    def update_dev(self, packed_dev_ptr: int, rows: int, bps: int, row_maf_dev_ptr=None, qc=None) -> None:
        """Same as update() for packed rows already in HBM on this handle's device (raw device pointers)."""
        cfg = None if qc is None else QcCfg(float(qc[0]), float(qc[1]), float(qc[2]), 0)
        check(lib().jxb_grm_update_dev(self.handle, packed_dev_ptr, int(bps), int(rows), row_maf_dev_ptr,
                                       C.byref(cfg) if cfg is not None else None))

    def eigh_dev(self, evals_dev_ptr: int, ut_f32_dev_ptr: int, diag_shift: float = 1e-6) -> None:
        """Finish, then decompose in place with outputs left on the device (f64[n], f32[n, n] caller buffers)."""
        self.finish(to_host=False)
        check(lib().jxb_eigh_dev(self.device, self.n, lib().jxb_grm_device_matrix(self.handle), float(diag_shift),
                                 evals_dev_ptr, ut_f32_dev_ptr, lib().jxb_grm_stream(self.handle)))
        check(lib().jxb_grm_finish(self.handle, None, None))   # synchronises the handle's stream

    @property
    def rows_used(self) -> int:
        return int(lib().jxb_grm_rows_used(self.handle))

    def finish(self, to_host: bool = True):
        """-> (K f64[n, n] or None, varsum)."""
        k = np.empty((self.n, self.n), dtype=np.float64) if to_host else None
        vs = C.c_double()
        check(lib().jxb_grm_finish(self.handle, ptr(k), C.byref(vs)))
        return k, float(vs.value)

    def eigh(self, diag_shift: float = 1e-6):
        """Finish, then decompose K + diag_shift*I in place on the device -> (S f64[n], U^T f32[n, n]) on the host.
        The n x n matrix never visits the host in f64."""
        w = np.empty(self.n, dtype=np.float64)
        ut32 = np.empty((self.n, self.n), dtype=np.float32)
        check(lib().jxb_grm_eigh(self.handle, float(diag_shift), ptr(w), ptr(ut32)))
        return w, ut32

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().jxb_grm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _grm_packed(packed, n_samples, row_flip, row_maf, sample_indices, method, progress_callback, progress_every):
    if n_samples == 0:
        raise RuntimeError("n_samples must be > 0")
    if method not in (1, 2, 3):
        raise RuntimeError(f"unsupported method={method}; expected 1 (centered additive), 2 (standardized additive), "
                           "or 3 (centered dominance)")
    if method != 1:
        raise NotImplementedError("only method=1 (centred additive, the `-k 1` default) runs on the device in this build")
    packed = np.asarray(packed)
    if packed.ndim != 2:
        raise RuntimeError("packed must be 2D (m, bytes_per_snp)")
    m, bps = packed.shape
    if m == 0:
        raise RuntimeError("packed must contain at least one SNP row")
    if bps != (n_samples + 3) // 4:
        raise RuntimeError(f"packed second dimension mismatch: got {bps}, expected {(n_samples + 3) // 4}")
    if np.asarray(row_flip).shape[0] != m:
        raise RuntimeError(f"row_flip length mismatch: got {np.asarray(row_flip).shape[0]}, expected {m}")
    if np.asarray(row_maf).shape[0] != m:
        raise RuntimeError(f"row_maf length mismatch: got {np.asarray(row_maf).shape[0]}, expected {m}")
    if np.any(np.asarray(row_flip, dtype=bool)):
        raise NotImplementedError("row_flip=True is not supported")
    if sample_indices is not None and np.asarray(sample_indices).size == 0:
        raise RuntimeError("sample_indices must not be empty")
    g = DeviceGrm(n_samples, sample_indices, method)
    try:
        maf = np.asarray(row_maf, dtype=np.float32)
        step = 65536 if not progress_every else max(1, int(progress_every))
        for r0 in range(0, m, step):
            g.update(packed[r0:r0 + step], maf[r0:r0 + step])
            if progress_callback is not None:
                progress_callback(min(m, r0 + step), m)
        k, _ = g.finish()
    finally:
        g.close()
    return k


def grm_packed_f64(packed, n_samples, row_flip, row_maf, sample_indices=None, method=1, block_cols=65536, threads=0,
                   progress_callback=None, progress_every=0):
    """src/stats/grm.rs:3583-3623 -> f64[n, n]."""
    return _grm_packed(packed, n_samples, row_flip, row_maf, sample_indices, method, progress_callback, progress_every)


def grm_packed_f32(packed, n_samples, row_flip, row_maf, sample_indices=None, method=1, block_cols=65536, threads=0,
                   progress_callback=None, progress_every=0):
    """src/stats/grm.rs:3053-3581 -> f32[n, n] (the f64 accumulator cast once, grm.rs:484)."""
    return _grm_packed(packed, n_samples, row_flip, row_maf, sample_indices, method, progress_callback,
                       progress_every).astype(np.float32)


def rust_eigh_from_array_f64(a, threads=0, driver=None, jobz="V", require_lapack=False):
    """src/math/eigh.rs:1621-1705 -> (evals, evecs | None, backend, evd_backend, n, threads_before, threads_in_stage,
    threads_after, lapack_used, elapsed_s).  Eigenvalues ascending, eigenvectors in columns (numpy convention)."""
    import time
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[0] == 0 or a.shape[0] != a.shape[1]:
        shape = tuple(a.shape) + (0, 0)
        raise RuntimeError(f"rust_eigh_from_array_f64 expects a non-empty square matrix; got shape=({shape[0]}, {shape[1]})")
    if require_lapack:
        raise RuntimeError("rust_eigh_from_array_f64 expected LAPACK backend, got cusolver_xsyevd")
    _cabi.require_gpu()
    a = _f64(a)
    n = a.shape[0]
    want_v = str(jobz).strip().upper() != "N"
    w = np.empty(n, dtype=np.float64)
    ut = np.empty((n, n), dtype=np.float64) if want_v else None
    t0 = time.perf_counter()
    check(lib().jxb_eigh(0, n, ptr(a), 0.0, ptr(w), ptr(ut), None))
    dt = time.perf_counter() - t0
    return w, (ut.T if want_v else None), "cuda", "cusolver_xsyevd", n, 0, 0, 0, False, dt


rust_eigh_from_array_f64_inplace = rust_eigh_from_array_f64


def vcf_to_plink(vcf_path, out_prefix, snps_only=False):
    """VCF(.gz) -> out_prefix.bed/.bim/.fam with the reference's GT rules (src/io/gfcore.rs:2875-2980); host only.
    -> (n_samples, n_sites)."""
    ns, nv = C.c_size_t(), C.c_size_t()
    check(lib().jxb_vcf_to_plink(str(vcf_path).encode(), str(out_prefix).encode(), 1 if snps_only else 0,
                                 C.byref(ns), C.byref(nv)))
    return int(ns.value), int(nv.value)


def gwas_lmm_lm_null_lrt_decision(y, x_cov, lmm_ml0, alpha=0.05, boundary_mixture=True):
    """src/stats/gwas_unified.rs:119-175 -> (switch_to_lm, lrt_stat, pval, lm_ml0).

    Host-side bookkeeping, once per trait (an n x (p+1) least-squares fit): H0 "Va = 0" is tested with the LRT of
    the LMM null ML against the plain linear-model ML; a non-significant test means the caller may fall back to LM.
    """
    if not math.isfinite(lmm_ml0):
        raise RuntimeError("lmm_ml0 must be finite")
    if not (math.isfinite(alpha) and 0.0 < alpha < 1.0):
        raise RuntimeError("alpha must be in (0,1)")
    y = _f64(y).reshape(-1)
    x = _f64(x_cov)
    if x.ndim != 2:
        raise RuntimeError("x_cov must be 2D")
    n, p_cov = y.shape[0], x.shape[1]
    if x.shape[0] != n:
        raise RuntimeError("x_cov rows must equal len(y)")
    if n <= p_cov + 1:
        raise RuntimeError("insufficient samples: require n > p_cov + 1")
    design = np.concatenate([np.ones((n, 1)), x], axis=1)          # gwas_unified.rs:63-70: intercept + covariates
    xtx = design.T @ design
    xty = design.T @ y
    try:
        chol = np.linalg.cholesky(xtx)
    except np.linalg.LinAlgError:
        try:
            chol = np.linalg.cholesky(xtx + 1e-8 * np.eye(p_cov + 1))
        except np.linalg.LinAlgError:
            raise RuntimeError("failed to compute LM null log-likelihood")
    beta = np.linalg.solve(chol.T, np.linalg.solve(chol, xty))
    resid = y - design @ beta
    rss = float(resid @ resid)
    if not (math.isfinite(rss) and rss > 0.0):
        raise RuntimeError("failed to compute LM null log-likelihood")
    n_f = float(n)
    lm_ml0 = n_f * (math.log(n_f) - 1.0 - math.log(2.0 * math.pi)) / 2.0 - 0.5 * n_f * math.log(rss)
    stat = 2.0 * (lmm_ml0 - lm_ml0)
    if not math.isfinite(stat) or stat < 0.0:
        stat = 0.0
    pval = math.erfc(math.sqrt(0.5 * stat)) if stat > 0.0 else 1.0   # chi2_sf_df1, src/math/linalg.rs:7-17
    if boundary_mixture:
        pval *= 0.5
    if not math.isfinite(pval):
        pval = 1.0
    pval = min(max(pval, 2.2250738585072014e-308), 1.0)
    return bool(pval >= alpha), float(stat), float(pval), float(lm_ml0)


def __getattr__(name):
    # `janusx.janusx.BedChunkReader` lives in gfreader.py (which imports this module): resolve it lazily
    if name in ("BedChunkReader", "BedChunkReaderFromMeta", "prepare_bed_logic_meta_selected", "prepare_bed_logic_keep_mask"):
        from . import gfreader
        return getattr(gfreader, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
