"""Builds libvenusaur_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m venusaur_b200.build [--force]

nvcc cross-compiles without a GPU.  path_kernels.cu and wavefront.cu are compiled twice:
  exact : -DVN_EXACT=1 -fmad=false           (IEEE, bit-identical to the oracle)
  fast  : -DVN_EXACT=0 -fmad=true -ftz=true  (the benchmarked build)
lbvh.cu is compiled with -fmad=false so Morton quantisation is reproducible on the host.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libvenusaur_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden"]

UNITS = [
    # (source, object, extra flags)
    ("vn_api.cu", "vn_api.o", []),
    ("lbvh.cu", "lbvh.o", ["-fmad=false"]),
    ("grid.cu", "grid.o", ["-fmad=false"]),
    ("path_kernels.cu", "path_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("path_kernels.cu", "path_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
    ("wavefront.cu", "wavefront_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("wavefront.cu", "wavefront_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
    ("pool_kernels.cu", "pool_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("pool_kernels.cu", "pool_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
    ("slot_kernels.cu", "slot_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("slot_kernels.cu", "slot_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libvenusaur_b200.so")
    return exe


def _sources() -> list[str]:
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    inc = os.path.join(HERE, "..", "include")
    for root, _, files in os.walk(inc):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.abspath(__file__))
    return out


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    jobs = []
    for src, obj, extra in UNITS:
        cmd = [cc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, obj) for _, obj, _ in UNITS]
    _run([cc, *ARCH, "-shared", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
