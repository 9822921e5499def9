"""Builds the CUDA library in-tree: game_engine_b200/libgame_engine_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libgame_engine_b200.so")
SOURCES = ["ge_capi.cu"]
HEADERS = ["ge_common.cuh", "ge_step_tps.cuh", "ge_step_coop.cuh", "ge_spec_gen.cuh", "ge_glue.cuh", os.path.join("..", "..", "include", "game_engine_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", *(os.environ.get("GE_EXTRA_NVCC", "").split()),
    "--shared", "-Xcompiler", "-fPIC", "-cudart", "static",
]


# bench-only helper (head-start spin kernel): kept OUT of the product library
AUX_SRC = os.path.join(os.path.dirname(_HERE), "tools", "benchaux", "ge_benchaux.cu")
AUX_LIB = os.path.join(os.path.dirname(_HERE), "tools", "benchaux", "libge_benchaux.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_aux() -> str:
    if os.path.exists(AUX_SRC) and (not os.path.exists(AUX_LIB) or os.path.getmtime(AUX_LIB) < os.path.getmtime(AUX_SRC)):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", AUX_LIB, AUX_SRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("nvcc failed on the bench helper (exit %d)" % res.returncode)
    return AUX_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    # constexpr views of the shipped games' tables (regenerated only when their content changes)
    sys.path.insert(0, os.path.dirname(_HERE))
    from game_engine_b200 import specgen
    specgen.write()
    build_aux()
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d)" % res.returncode)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
