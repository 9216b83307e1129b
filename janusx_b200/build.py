"""Build libjxb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = Path(os.environ["JXB_BUILD_OUT"]) if os.environ.get("JXB_BUILD_OUT") else PKG / "libjxb200.so"   # override: kernel A/B builds
# (source, object stem, extra defines).  k3_inst.cu is compiled once per static covariate count so the
# heavy fully-unrolled K3 kernels build in parallel.
UNITS = [("cabi.cu", "cabi", []), ("k1_decode.cu", "k1_decode", []), ("k2_rotate.cu", "k2_rotate", []), ("k2_int8.cu", "k2_int8", []), ("k2_i8mma.cu", "k2_i8mma", []),
         ("k3_solve.cu", "k3_solve", ["-fmad=false"]), ("bed_scan.cpp", "bed_scan", []), ("grm.cu", "grm", []), ("eigh.cu", "eigh", []),
         ("vcf_cache.cpp", "vcf_cache", []), ("probe.cu", "probe", [])]
# -fmad=false: K3 reproduces the reference's separate multiply/add rounding (Rust never fuses)
_K3X = os.environ.get("JXB_K3_FLAGS", "").split()   # e.g. "-DJXB_K3_BUFS=1 -DJXB_K3_MINB=4" (tuning experiments)
UNITS += [("k3_inst.cu", f"k3_inst_p{p}", [f"-DJXB_P={p}", "-fmad=false", *_K3X]) for p in range(1, 9)]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = [CSRC / "jxb_common.cuh", CSRC / "k3_solve.cuh", CSRC / "tc05.cuh", PKG.parent / "include" / "jxb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fno-fast-math",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = Path(os.environ["JXB_BUILD_OBJDIR"]) if os.environ.get("JXB_BUILD_OBJDIR") else PKG / "build"
    objdir.mkdir(parents=True, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc wrapper without OpenMP specs; nvcc wants the system g++
    ccbin = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    objs = []
    procs = []
    for src, stem, defs in UNITS:
        obj = objdir / (stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-ccbin", ccbin, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if src.endswith(".cpp"):
            cmd[1:1] = ["-x", "cu"]
        procs.append((stem, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose and out:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", ccbin, "-o", str(OUT), *objs, "-lcudart", "-lpthread", "-ldl"]
    subprocess.run(link, check=True, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
