"""
Builds the checker's helpers (TEST / BENCH INFRASTRUCTURE -- never imported by the pyxu_b200 package):

* `build_port()`  oracle/tv_oracle.c -> oracle/libtv_oracle.so   (C/OpenMP restatement: second checker, CPU baseline "port")
* `build_ref()`   stages the reference's own pure-Python package (read-only at /root/reference in the build container) into the
                  git-ignored oracle/_ref/ so that it travels to the GPU box with the snapshot: `bench.py --impl reference` and
                  `cpu_baseline` then time the REAL pyxu NumPy/Numba path (kind "reference").  No file of it enters the history.
* `load_ref()`    imports the staged reference (two of its hard dependencies, dask and sparse, are absent from the image and
                  never touched by the NumPy path: inert stubs for this process only, see tests/golden/_ref_import.py).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/pyxu"
REF_DST = os.path.join(HERE, "_ref")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_port(force=False):
    src, out = os.path.join(HERE, "tv_oracle.c"), os.path.join(HERE, "libtv_oracle.so")
    if force or _newer(out, [src]):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"], check=True)
    return out


def build_ref(force=False):
    """oracle/_ref/pyxu = the reference package as it lies under /root/reference (only where that exists: the build container)."""
    dst = os.path.join(REF_DST, "pyxu")
    if not os.path.isdir(REF_SRC):
        return dst if os.path.isdir(dst) else None
    if force or not os.path.isdir(dst):
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(REF_DST, exist_ok=True)
        shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return dst


def load_ref():
    """Namespace (abc, operator, solver, stop) of the staged reference; raises ImportError with the reason when it cannot run here."""
    if not os.path.isdir(os.path.join(REF_DST, "pyxu")):
        raise ImportError("oracle/_ref/pyxu is not staged (run __graft_entry__.build() where /root/reference exists)")
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests", "golden"))
    import _ref_import

    return _ref_import.load(REF_DST)


if __name__ == "__main__":
    print(build_port(force="--force" in sys.argv), build_ref(force="--force" in sys.argv))
