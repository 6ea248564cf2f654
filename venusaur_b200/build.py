"""Builds libvenusaur_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m venusaur_b200.build [--force]

nvcc cross-compiles without a GPU.  path_kernels.cu and wavefront.cu are compiled twice:
  exact : -DVN_EXACT=1 -fmad=false           (IEEE, bit-identical to the oracle: the default and the benchmarked build)
  fast  : -DVN_EXACT=0 -fmad=true -ftz=true  (opt-in VN_FAST: relaxed numerics, not within the image tolerance)
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
    ("vn_multi.cu", "vn_multi.o", []),
    ("lbvh.cu", "lbvh.o", ["-fmad=false"]),
    ("grid.cu", "grid.o", ["-fmad=false"]),
    ("path_kernels.cu", "path_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("path_kernels.cu", "path_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
    ("wavefront.cu", "wavefront_exact.o", ["-DVN_EXACT=1", "-fmad=false"]),
    ("wavefront.cu", "wavefront_fast.o", ["-DVN_EXACT=0", "-fmad=true", "-ftz=true"]),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libvenusaur_b200.so")
    return exe


def _sources() -> list[str]:
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    inc = os.path.join(HERE, "..", "include")
    for root, _, files in sorted(os.walk(inc)):
        out += [os.path.join(root, f) for f in sorted(files)]
    out.append(os.path.abspath(__file__))
    return out


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """Content hash of everything the library is built from.  File times do not survive a copy of the tree (a snapshot sent to
    another machine arrives with fresh mtimes in arbitrary order), contents do."""
    import hashlib
    h = hashlib.sha256()
    for path in _sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def up_to_date() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return False
    try:
        with open(STAMP) as f:
            return f.read().strip() == source_hash()
    except OSError:
        return False


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds the library unless it is up to date.  Safe against concurrent callers (one process per GPU all importing the package
    at once): an exclusive file lock serialises them, the objects and the library are written under process-private names and
    renamed into place, and whoever gets the lock second finds the work done."""
    if not force and up_to_date():
        return LIB
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():
                return LIB
            cc = nvcc()
            tag = ".%d.tmp" % os.getpid()
            jobs = []
            for src, obj, extra in UNITS:
                cmd = [cc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj + tag + ".o")]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                jobs.append(cmd)
            with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
                list(ex.map(_run, jobs))
            objs = []
            for _, obj, _ in UNITS:
                os.replace(os.path.join(OBJ, obj + tag + ".o"), os.path.join(OBJ, obj))
                objs.append(os.path.join(OBJ, obj))
            _run([cc, *ARCH, "-shared", "-o", LIB + tag, *objs])
            os.replace(LIB + tag, LIB)
            with open(STAMP + tag, "w") as f:
                f.write(source_hash())
            os.replace(STAMP + tag, STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_experiment(n: int) -> str:
    """A second library next to the shipped one, compiled with -DVN_EXP=n (code under `#if VN_EXP == n` in csrc/): A/B measurements of
    compile-time variants within one visit to the GPU (VN_EXPERIMENT=n selects it in _lib.load()).  Test infrastructure only."""
    cc = nvcc()
    out = os.path.join(HERE, "libvenusaur_b200_exp%d.so" % n)
    odir = os.path.join(OBJ, "exp%d" % n)
    os.makedirs(odir, exist_ok=True)
    jobs, objs = [], []
    for src, obj, extra in UNITS:
        o = os.path.join(odir, obj)
        jobs.append([cc, *ARCH, *COMMON, *extra, "-DVN_EXP=%d" % n, "-c", os.path.join(CSRC, src), "-o", o])
        objs.append(o)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        list(ex.map(_run, jobs))
    _run([cc, *ARCH, "-shared", "-o", out, *objs])
    return out


if __name__ == "__main__":
    if "--exp" in sys.argv:
        print(build_experiment(int(sys.argv[sys.argv.index("--exp") + 1])))
    else:
        path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(path)
