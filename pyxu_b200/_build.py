"""
In-tree build of libpyxu_b200.so with nvcc for sm_100a (and of the test-only helpers).

    python -m pyxu_b200._build            # build the CUDA library
    python -m pyxu_b200._build --all      # + oracle C helper + host emulation used by CPU tests
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


def build_emu(force=False):
    """Host build of the per-voxel kernel bodies (tests/emu): CPU test infrastructure only."""
    src = os.path.join(ROOT, "tests", "emu", "pxb_emu.cpp")
    out = os.path.join(ROOT, "tests", "emu", "libpxb_emu.so")
    deps = [src, os.path.join(ROOT, "include", "pyxu_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or _newer(out, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", out, src], check=True)
    return out


def build_oracle(force=False):
    """C restatement used as the multi-threaded CPU baseline (oracle/): test/bench infrastructure only."""
    src = os.path.join(ROOT, "oracle", "tv_oracle.c")
    out = os.path.join(ROOT, "oracle", "libtv_oracle.so")
    if not os.path.exists(src):
        return None
    if force or _newer(out, [src]):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"], check=True)
    return out


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--all" in sys.argv:
        print(build_emu(), build_oracle())
