#!/usr/bin/env python
"""Times Stencil.apply (TMA-tiled vs generic kernels) on the shapes of BASELINE.json configs[1] / [2].
Usage: python tools/bench_stencil.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return (k / k.sum()).astype(np.float32)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, shape, batch, kern, cen in [
    ("8192^2 separable 9x9", (8192, 8192), 1, [gauss(9, 1.7), gauss(9, 1.7)], (4, 4)),
    ("8192^2 dense 9x9", (8192, 8192), 1, np.outer(gauss(9, 1.7), gauss(9, 1.7)), (4, 4)),
    ("64 x 1024^2 dense 5x5", (1024, 1024), 64, np.outer(gauss(5, 1.0), gauss(5, 1.0)), (2, 2)),
    ("64 x 1024^2 separable 5x5", (1024, 1024), 64, [gauss(5, 1.0), gauss(5, 1.0)], (2, 2)),
]:
    x = torch.randn(batch, int(np.prod(shape)), device="cuda", dtype=torch.float32)
    nbytes = 8 * x.numel()
    for label, force in (("tiled", None), ("generic", False)):
        op = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        op._tiled_ok = force
        ms = timeit(lambda: op.apply(x))
        print(f"{name:28s} {label:8s} {ms:8.3f} ms   {nbytes / ms / 1e6:7.0f} GB/s (8 B/voxel)", flush=True)

# Folding boundary modes: Pad -> tiled stencil / tiled stencil -> Pad^T (Stencil._run_padded, PYXU_B200_STENCIL_PADDED=1)
# against the gather kernels.
from pyxu_b200.operator.linop import stencil as _st

if _st.PADDED_TILED:
    for name, shape, kern, cen in [
        ("8192^2 separable 9x9", (8192, 8192), [gauss(9, 1.7), gauss(9, 1.7)], (4, 4)),
        ("8192^2 dense 5x5", (8192, 8192), np.outer(gauss(5, 1.0), gauss(5, 1.0)) + np.float32(0.01) * np.arange(25, dtype=np.float32).reshape(5, 5), (2, 2)),
    ]:
        x = torch.randn(1, int(np.prod(shape)), device="cuda", dtype=torch.float32)
        nbytes = 8 * x.numel()
        for mode in ("reflect", "wrap"):
            for label, force in (("padded", None), ("gather", False)):
                op = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode=mode)
                op._padded_ok = force
                for adj in (False, True):
                    ms = timeit(lambda: op.adjoint(x) if adj else op.apply(x))
                    print(f"{name:22s} {mode:8s} {'adjoint' if adj else 'apply':8s} {label:8s} {ms:8.3f} ms   {nbytes / ms / 1e6:7.0f} GB/s (8 B/voxel)", flush=True)
    # 3-D separable 7x7x7 with a folding mode on every axis: streaming axis-0 pass with the boundary map + the two padded passes
    shape = (256, 1024, 1024)
    x = torch.randn(1, int(np.prod(shape)), device="cuda", dtype=torch.float32)
    for label, force in (("padded", None), ("gather", False)):
        op = pxo.Stencil(arg_shape=shape, kernel=[gauss(7, 1.2)] * 3, center=(3, 3, 3), mode="reflect")
        op._padded_ok = force
        for adj in (False, True):
            ms = timeit(lambda: op.adjoint(x) if adj else op.apply(x), reps=3)
            print(f"256x1024^2 sep 7x7x7 reflect {'adjoint' if adj else 'apply':8s} {label:8s} {ms:8.3f} ms   {8 * x.numel() / ms / 1e6:7.0f} GB/s (8 B/voxel)", flush=True)
