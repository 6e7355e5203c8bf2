"""
In-tree build of libpyxu_b200.so with nvcc for sm_100a.

    python -m pyxu_b200._build            # build the CUDA library

(The checker's helpers -- the oracle's C port, the staged reference, the host emulation of the kernel bodies -- are built by
oracle/build.py and tests/emu/build.py: the package names nothing under oracle/ or tests/.)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pyxu_b200", "csrc")
LIBDIR = os.path.join(ROOT, "pyxu_b200", "lib")
LIB = os.path.join(LIBDIR, "libpyxu_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    *os.environ.get("PYXU_B200_NVCC_EXTRA", "").split(),  # e.g. -DPXB_EXPERIMENT for the A/B switches of tools/bench_criterion.py
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libpyxu_b200.so (sm_100a)")
    return nvcc


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "pyxu_b200.h")]
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    if force or _newer(LIB, deps):
        from concurrent.futures import ThreadPoolExecutor

        nvcc = _nvcc()

        def compile_one(src):  # one translation unit -> one object, all units in parallel
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, src]
            r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
            return obj

        with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            objs = list(ex.map(compile_one, srcs))
        subprocess.run([nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs], check=True)
    return LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
