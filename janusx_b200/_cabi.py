"""ctypes binding of libjxb200.so (include/jxb200.h).

This is the only way the package reaches compute: there is no Python/NumPy/torch implementation of the
scan behind it.  Importing works without a GPU (so CPU-only hosts can build and inspect symbols), but
every compute call fails loudly when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["JXB_LIB_PATH"]) if os.environ.get("JXB_LIB_PATH") else _PKG / "libjxb200.so"   # override: kernel A/B builds
_lib = None


class JxbError(RuntimeError):
    """Raised for every non-zero return of the C ABI (mirrors PyRuntimeError in the reference)."""


class SolveCfg(C.Structure):
    _fields_ = [("low", C.c_double), ("high", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int32),
                ("has_init", C.c_int32), ("init_log10_lbd", C.c_double), ("has_nullml", C.c_int32),
                ("nullml", C.c_double)]


class QcCfg(C.Structure):
    _fields_ = [("maf_thr", C.c_float), ("miss_thr", C.c_float), ("het_thr", C.c_float),
                ("genetic_model", C.c_int32)]


PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_size_t, C.c_void_p)


class BedScanCfg(C.Structure):
    _fields_ = [("bed_prefix", C.c_char_p), ("out_tsv", C.c_char_p), ("qc", QcCfg), ("solve", SolveCfg),
                ("mode", C.c_int32), ("snps_only", C.c_int32), ("sample_ids", C.POINTER(C.c_char_p)),
                ("n_sample_ids", C.c_size_t), ("batch_rows", C.c_size_t), ("snp_begin", C.c_size_t),
                ("snp_end", C.c_size_t), ("write_header", C.c_int32), ("progress_every", C.c_size_t),
                ("row_indices", C.POINTER(C.c_int64)), ("n_row_indices", C.c_size_t),
                ("row_maf", C.POINTER(C.c_float)), ("row_flip", C.POINTER(C.c_uint8)),
                ("row_missing", C.POINTER(C.c_float)), ("mmap_window_mb", C.c_size_t)]


# every symbol include/jxb200.h declares: (name, restype, argtypes or None)
_vp = C.c_void_p
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pu8 = C.POINTER(C.c_uint8)
_psz = C.POINTER(C.c_size_t)
SYMBOLS = {
    "jxb_last_error": (C.c_char_p, []),
    "jxb_device_count": (C.c_int, []),
    "jxb_build_info": (C.c_char_p, []),
    "jxb_launch_count": (C.c_uint64, []),
    "jxb_model_create": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "jxb_model_create_dev": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "jxb_model_destroy": (None, [_vp]),
    "jxb_model_set_xy": (C.c_int, [_vp, _vp, _vp]),
    "jxb_model_sync": (C.c_int, [_vp]),
    "jxb_rotate_xy": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "jxb_reml_null": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_double, _pd]),
    "jxb_ml_loglike_null": (C.c_int, [_vp, C.c_double, _pd]),
    "jxb_ml_null": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, _pd]),
    "jxb_lmm_reml_chunk_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(SolveCfg), _vp, _vp]),
    "jxb_lmm_reml_chunk_from_snp_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(SolveCfg), _vp, _vp]),
    "jxb_lmm2_chunk_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.POINTER(SolveCfg), _vp, _vp]),
    "jxb_lmm_fixed_chunk_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_double, _pd, _vp, _pd]),
    "jxb_rotate_block_f32": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_int]),
    "jxb_scan_packed": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, C.POINTER(QcCfg),
                                  C.POINTER(SolveCfg), C.c_int, _vp, _vp, _vp, _vp, _vp, _psz]),
    "jxb_scan_packed_prepared": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp,
                                           C.POINTER(QcCfg), C.POINTER(SolveCfg), C.c_int, _vp, _vp, _vp, _vp, _vp, _psz]),
    "jxb_stage_packed": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t]),
    "jxb_scan_staged_begin": (C.c_int, [_vp]),
    "jxb_scan_staged": (C.c_int, [_vp, C.c_size_t, _vp, _vp, _vp, _vp, C.POINTER(QcCfg), C.POINTER(SolveCfg), C.c_int,
                                  _vp, _vp, _vp, _vp, _vp, _psz]),
    "jxb_stage_cancel": (None, [_vp]),
    "jxb_scan_packed_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, C.POINTER(QcCfg),
                                      C.POINTER(SolveCfg), C.c_int]),
    "jxb_scan_fetch": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp, _vp, _vp, _vp, _vp, _psz]),
    "jxb_scan_fetch_dev": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp, _vp, _vp, _psz]),
    "jxb_decode_packed": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, C.POINTER(QcCfg), _vp, _vp,
                                    _vp, _vp, _psz]),
    "jxb_decode_packed_prepared": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp, C.c_int, _vp,
                                             _psz]),
    "jxb_decode_packed_lut": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp]),
    "jxb_debug_fetch_rot": (C.c_int, [_vp, C.c_size_t, C.c_size_t, _vp]),
    "jxb_set_timing": (None, [C.c_int]),
    "jxb_last_stage_ms": (C.c_int, [_vp, _pf]),
    "jxb_model_stream": (_vp, [_vp]),
    "jxb_fp64_probe": (C.c_int, [C.c_int, _pd]),
    "jxb_set_rotate_variant": (None, [C.c_int]),
    "jxb_set_thread_solve_min_rows": (None, [C.c_size_t]),
    "jxb_set_big_solve_kernel": (None, [C.c_int]),
    "jxb_set_stream_overlap": (None, [C.c_int, C.c_size_t]),
    "jxb_set_generic_divide": (None, [C.c_int]),
    "jxb_set_prefix_evals": (None, [C.c_int]),
    "jxb_set_fixed_lane_min_rows": (None, [C.c_size_t]),
    "jxb_selftest_rcp": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "jxb_last_stage_ms8": (C.c_int, [_vp, _pf]),
    "jxb_scan_bed_to_tsv": (C.c_int, [_vp, C.POINTER(BedScanCfg), _psz, PROGRESS_CB, _vp]),
    "jxb_format_row": (C.c_size_t, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64, C.c_char_p, C.c_char_p,
                                    C.c_char_p, C.c_float, C.c_float, _pd, C.c_int]),
    "jxb_format_block": (C.c_size_t, [_vp, C.c_size_t, C.c_size_t, C.c_char_p, _vp, C.c_char_p, C.c_char_p, C.c_char_p,
                                      _vp, _vp, _vp, C.c_int, C.c_int]),
    "jxb_tsv_header": (C.c_char_p, [C.c_int]),
    "jxb_selftest_format": (C.c_size_t, [C.c_size_t, C.c_uint64, C.c_int, C.c_char_p, C.c_size_t]),
    "jxb_host_checksum": (None, [_vp, C.c_size_t, C.POINTER(C.c_uint64)]),
    "jxb_grm_create": (C.c_int, [C.c_int, C.c_size_t, _vp, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "jxb_grm_update": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.POINTER(QcCfg)]),
    "jxb_grm_update_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.POINTER(QcCfg)]),
    "jxb_grm_rows_used": (C.c_size_t, [_vp]),
    "jxb_grm_finish": (C.c_int, [_vp, _vp, _pd]),
    "jxb_grm_device_matrix": (_vp, [_vp]),
    "jxb_grm_stream": (_vp, [_vp]),
    "jxb_grm_destroy": (None, [_vp]),
    "jxb_vcf_to_plink": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _psz, _psz]),
    "jxb_grm_eigh": (C.c_int, [_vp, C.c_double, _vp, _vp]),
    "jxb_set_eigh_devices": (None, [C.c_int]),
    "jxb_eigh": (C.c_int, [C.c_int, C.c_size_t, _vp, C.c_double, _vp, _vp, _vp]),
    "jxb_eigh_dev": (C.c_int, [C.c_int, C.c_size_t, _vp, C.c_double, _vp, _vp, _vp]),
}


def _preload_cuda_math_libs() -> None:
    """libjxb200.so opens cuSOLVER / cuBLASLt by soname at first use.  A process that imported torch already holds the
    copies bundled with the `nvidia-*` wheels, one that did not would pick the toolkit's (another version): the same GRM
    then decomposes to different last bits depending on whether torch happens to be loaded (seen: `-gpus 1` vs `-gpus 2`).
    Load the bundled copies first, when they exist, so every process resolves the same libraries."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia")
    except (ImportError, ValueError):
        spec = None
    roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    for rel in ("cuda_runtime/lib/libcudart.so.12", "nvjitlink/lib/libnvJitLink.so.12", "cublas/lib/libcublasLt.so.12",
                "cublas/lib/libcublas.so.12", "cusparse/lib/libcusparse.so.12", "cusolver/lib/libcusolver.so.11"):
        for root in roots:
            path = Path(root) / rel
            if path.exists():
                try:
                    C.CDLL(str(path), mode=C.RTLD_GLOBAL)
                except OSError:
                    pass
                break


def lib() -> C.CDLL:
    """Load libjxb200.so (built in-tree by janusx_b200/build.py).  No fallback of any kind."""
    global _lib
    if _lib is None:
        _preload_cuda_math_libs()
        if not LIB_PATH.exists():
            raise JxbError(
                f"{LIB_PATH} is missing: build it with `python -m janusx_b200.build` "
                "(nvcc, sm_100a). janusx_b200 has no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().jxb_last_error()
        raise JxbError(msg.decode() if msg else f"libjxb200 error {rc}")


def require_gpu() -> int:
    n = int(lib().jxb_device_count())
    if n <= 0:
        raise JxbError("no CUDA device is visible: janusx_b200 has no CPU fallback")
    return n


def ptr(a):
    """Raw address of a C-contiguous numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(type(a))


def launch_count() -> int:
    return int(lib().jxb_launch_count())
