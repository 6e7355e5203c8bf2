"""Host build of the kernel bodies (tests/emu/pxb_emu.cpp -> libpxb_emu.so): CPU test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pyxu_b200", "csrc")


def build_emu(force=False):
    src, out = os.path.join(HERE, "pxb_emu.cpp"), os.path.join(HERE, "libpxb_emu.so")
    deps = [src, os.path.join(ROOT, "include", "pyxu_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", out, src], check=True)
    return out
