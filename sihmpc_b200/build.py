# -*- coding: utf-8 -*-
"""In-tree build of libsihmpc_b200.so for sm_100a (nvcc cross-compiles without a GPU).
The translation units are compiled in parallel and linked into one shared library."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libsihmpc_b200.so")
# (source, extra defines, object, dependencies): the tile kernels are instantiated one tile-row count per unit
_TILE_DEPS = ["tile_kernel.cu", "tile_active_set.inc", "tile_api.h", "devplan.h", "../../include/sihmpc_b200.h"]
_MMA_DEPS = ["mma_units.cu", "mma_kernel.cuh", "mma_api.h", "dist_kernel.cuh", "step_kernel.cuh", "devplan.h",
             "../../include/sihmpc_b200.h"]
MMA_SHAPES = [(n, x) for n in (6, 8, 10, 12) for x in (16, 24, 32)]       # (horizon, state capacity) of the tensor fast path
UNITS = [("capi.cu", [], "capi.o", None), ("tile_kernel.cu", [], "tile_host.o", _TILE_DEPS)] + [
    ("tile_kernel.cu", ["-DSIH_TL_ONLY_NT=%d" % k], "tile_nt%d.o" % k, _TILE_DEPS) for k in range(1, 9)] + [
    ("mma_units.cu", [], "mma_host.o", _MMA_DEPS)] + [
    ("mma_units.cu", ["-DSIH_MMA_N=%d" % n, "-DSIH_MMA_NXP=%d" % x], "mma_%d_%d.o" % (n, x), _MMA_DEPS) for n, x in MMA_SHAPES]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
    out.append(os.path.join(HERE, "..", "include", "sihmpc_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _stale(obj, deps) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    srcs = _sources() if deps is None else [os.path.join(CSRC, d) for d in deps]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs = [], []
    for src, defs, oname, deps in UNITS:
        obj = os.path.join(OBJ, oname)
        objs.append(obj)
        if not force and not extra and not _stale(obj, deps):
            continue
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        jobs.append((oname, cmd))
    width = max(1, (os.cpu_count() or 4))
    running = []

    def reap(pr_entry):
        u, pr = pr_entry
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (u, out, err))
        if verbose:
            sys.stderr.write(err)

    for oname, cmd in jobs:
        if len(running) >= width:
            reap(running.pop(0))
        running.append((oname, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for entry in running:
        reap(entry)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                       cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
