"""Builds the CUDA library in-tree: game_engine_b200/libgame_engine_b200.so (sm_100a only).

The step kernels are instantiated in separate translation units (csrc/ge_kernels.h): ge_k_generic.cu once per
(family, record bucket), ge_k_spec.cu once per shipped table.  They are compiled in parallel, objects are cached
under csrc/_obj/ (rebuilt only when a source they include is newer), and linked with the host runtime ge_capi.cu.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
# GE_LIB_OUT: build a variant library (A/B of compile-time knobs with GE_EXTRA_NVCC) next to the product one; GE_LIB makes
# capi.py load it
LIB = os.environ.get("GE_LIB_OUT") or os.path.join(_HERE, "libgame_engine_b200.so")
PUBLIC_H = os.path.join("..", "..", "include", "game_engine_b200.h")
KERNEL_HEADERS = ["ge_common.cuh", "ge_step_tps.cuh", "ge_step_coop.cuh", "ge_spec_gen.cuh", "ge_kernels.h", PUBLIC_H]
KERNEL_SETS = [(1, 8), (1, 16), (1, 24), (1, 32), (2, 4), (2, 8), (2, 16), (2, 32)]      # ge_kernels.h GE_KERNEL_SETS

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]
EXTRA = os.environ.get("GE_EXTRA_NVCC", "").split()
NVCC_FLAGS = [*ARCH_FLAGS, *EXTRA, "--shared", "-Xcompiler", "-fPIC", "-cudart", "static"]       # single-file builds (bench helper)


# bench-only helper (head-start spin kernel): kept OUT of the product library
AUX_SRC = os.path.join(os.path.dirname(_HERE), "tools", "benchaux", "ge_benchaux.cu")
AUX_LIB = os.path.join(os.path.dirname(_HERE), "tools", "benchaux", "libge_benchaux.so")


def _nvcc() -> str:
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def units():
    """(object name, source, extra -D flags, headers it depends on) of every translation unit."""
    from game_engine_b200 import specgen
    out = [("ge_capi", "ge_capi.cu", [], ["ge_common.cuh", "ge_step_tps.cuh", "ge_spec_gen.cuh", "ge_kernels.h", "ge_glue.cuh", PUBLIC_H])]
    for fam, bucket in KERNEL_SETS:
        out.append(("ge_k_%d_%d" % (fam, bucket), "ge_k_generic.cu", ["-DGE_TU_FAM=%d" % fam, "-DGE_TU_BUCKET=%d" % bucket], KERNEL_HEADERS))
    for i in range(len(specgen.SPECS)):
        out.append(("ge_k_spec_%d" % i, "ge_k_spec.cu", ["-DGE_SPEC_INDEX=%d" % i], KERNEL_HEADERS))
    return out


def _flags_tag() -> str:
    return hashlib.sha1(" ".join(ARCH_FLAGS + EXTRA).encode()).hexdigest()[:8]


def _obj_path(name: str) -> str:
    return os.path.join(OBJ, "%s.%s.o" % (name, _flags_tag()))


def _unit_stale(name, src, headers) -> bool:
    o = _obj_path(name)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in headers] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(unit, verbose: bool):
    name, src, defs, _ = unit
    cmd = [_nvcc(), *ARCH_FLAGS, *EXTRA, *defs, "-Xcompiler", "-fPIC", "-c", "-o", _obj_path(name), os.path.join(CSRC, src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return name, res.returncode, res.stdout


def build_aux() -> str:
    if os.path.exists(AUX_SRC) and (not os.path.exists(AUX_LIB) or os.path.getmtime(AUX_LIB) < os.path.getmtime(AUX_SRC)):
        res = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", AUX_LIB, AUX_SRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
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
    us = units()
    # up to date when the library is newer than every source (the object cache does not travel to the GPU box)
    srcs = {os.path.join(CSRC, f) for u in us for f in [u[1]] + u[3]} | {__file__}
    if not force and not EXTRA and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(f) for f in srcs):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    todo = [u for u in us if force or _unit_stale(u[0], u[1], u[3])]
    jobs = max(1, min(len(todo), int(os.environ.get("GE_BUILD_JOBS", os.cpu_count() or 1))))
    failed = []
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as pool:
            for name, rc, out in pool.map(lambda u: _compile(u, verbose), todo):
                if verbose or rc != 0:
                    sys.stderr.write("---- %s\n%s" % (name, out))
                if rc != 0:
                    failed.append(name)
    if failed:
        raise RuntimeError("nvcc failed on: %s" % ", ".join(failed))
    cmd = [_nvcc(), *ARCH_FLAGS, "--shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", LIB] + [_obj_path(u[0]) for u in us]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc link failed (exit %d)" % res.returncode)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
