"""Build the sm_100a shared library in-tree (drvae_b200/lib/libdrvae_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the gpurun
snapshot.  `python -m drvae_b200.build` or `__graft_entry__.build()`.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdrvae_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in src_files)


def build(force=False, verbose=False, extra_flags=(), variant=""):
    """variant: suffix of an instrumented build (own object directory and library name), e.g. "waits" with
    extra_flags=["-DGEMM_PROFILE_WAITS"] -> lib/libdrvae_b200_waits.so."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + ("_" + variant if variant else ""))
    lib = LIB if not variant else os.path.join(LIBDIR, "libdrvae_b200_%s.so" % variant)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "drvae_b200.h"))
    jobs = []
    objs = []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = ["nvcc"] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if jobs or not os.path.exists(lib):
        run(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


if __name__ == "__main__":
    if "--waits" in sys.argv:  # instrumented variant for tools/wait_profile.py
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=["-DGEMM_PROFILE_WAITS"], variant="waits"))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
