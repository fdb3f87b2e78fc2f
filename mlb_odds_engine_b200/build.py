"""Builds libmlbsim_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m mlb_odds_engine_b200.build [--force]

The replay kernel is compiled with -fmad=false (bit-exact float64 compares, see
csrc/replay_flat_kernel.cu); everything carries -lineinfo so ncu's source page maps to the .cu files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB_PATH = os.path.join(PKG, "libmlbsim_b200.so")
OBJ_DIR = os.path.join(PKG, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE] + ARCH
COMMON += os.environ.get("MLBSIM_NVCC_DEFS", "").split()   # e.g. "-DOPT_ANYSYNC=1" (tuning sweeps)
UNITS = {
    "native_kernel.cu": [],
    "replay_flat_kernel.cu": ["-fmad=false"],
    "summary_kernel.cu": [],
    "alias_kernel.cu": ["-fmad=false"],
    "table_kernel.cu": ["-fmad=false"],
    "mlbsim_api.cu": [],
}
DEPS = ["native_kernel.h", "philox.cuh", "game_rules.cuh", "alias_row.cuh", os.path.join(INCLUDE, "mlbsim.h")]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built on this machine")
    return exe


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    srcs = [os.path.join(CSRC, u) for u in UNITS] + [d if os.path.isabs(d) else os.path.join(CSRC, d) for d in DEPS]
    return max(_mtime(s) for s in srcs) > _mtime(LIB_PATH)


def build_library(force: bool = False, verbose: bool = False, defs: list[str] | None = None, out: str | None = None) -> str:
    """Default: the product library.  `defs` / `out`: a tuning variant (extra -D switches) written to another path with
    its own object directory (scripts/build_variants.py; loaded through the MLBSIM_LIB environment variable)."""
    lib_path = out or LIB_PATH
    obj_dir = OBJ_DIR if out is None else os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(out))[0])
    if out is None and not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    for unit, extra in UNITS.items():
        obj = os.path.join(obj_dir, unit.replace(".cu", ".o"))
        cmd = [nvcc, *COMMON, *(defs or []), *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {unit}")
        objs.append(obj)
    tmp = lib_path + ".tmp"
    r = subprocess.run([nvcc, "-shared", *ARCH, "-o", tmp, *objs, "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
