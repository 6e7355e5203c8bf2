#!/usr/bin/env python
"""Operator-level timings (LinOp.apply / adjoint, ProxFunc.prox, stopping-criterion norms) on the shapes of BASELINE.json,
with the algorithmic HBM bytes of each call -> achieved GB/s.   python tools/bench_ops.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _kernels as kr


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


def report(name, ms, nbytes):
    print(f"{name:58s} {ms:8.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


for shape in ((1024, 1024, 1024), (8192, 8192)):
    N, D = int(np.prod(shape)), len(shape)
    G = pxo.Gradient(arg_shape=shape, dtype=np.float32)
    x = torch.rand(N, device="cuda")
    z = G.apply(x)
    report(f"Gradient.apply   {shape} fp32 (4 + {4 * D} B/voxel)", timeit(lambda: G.apply(x)), (4 + 4 * D) * N)
    report(f"Gradient.adjoint {shape} fp32 (4 + {4 * D} B/voxel)", timeit(lambda: G.adjoint(z)), (4 + 4 * D) * N)
    del z
    y = torch.rand(N, device="cuda")
    report(f"PositiveOrthant.prox(x - tau*y) {shape} (12 B/voxel)", timeit(lambda: kr.prox_lincomb((1, 0.0, 0.0), 0.5, 1.0, x, -0.5, y)), 12 * N)
    report(f"RelError norms (pxb_sqnorms) {shape} (8 B/voxel)", timeit(lambda: kr.sqnorms(x, y, rows=1)), 8 * N)
    L = pxo.L21Norm(arg_shape=(D, *shape), l2_axis=(0,))
    zz = torch.rand(D * N, device="cuda")
    report(f"L21Norm.prox {shape} ({8 * D} B/voxel)", timeit(lambda: L.prox(zz, 0.3)), 8 * D * N)
    del x, y, zz
shape = (512, 1024, 1024)
N = int(np.prod(shape))
x = torch.rand(N, device="cuda")
g7 = gauss(7, 1.2)
S = pxo.Stencil(arg_shape=shape, kernel=[g7, g7, g7], center=(3, 3, 3))
report(f"Stencil.apply   separable 7x7x7 {shape} fp32 (8 B/voxel)", timeit(lambda: S.apply(x)), 8 * N)
report(f"Stencil.adjoint separable 7x7x7 {shape} fp32 (8 B/voxel)", timeit(lambda: S.adjoint(x)), 8 * N)
S4 = pxo.Stencil(arg_shape=shape, kernel=[gauss(4, 1.0), g7, g7], center=(1, 3, 3))
report(f"Stencil.apply   separable 4x7x7 (axis-0 pass + tiled pass, 16 B/voxel)", timeit(lambda: S4.apply(x)), 16 * N)
