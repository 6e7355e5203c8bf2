#!/usr/bin/env python
"""Stencil.apply / .adjoint of a separable K x K x K PSF on a 3-D volume ('constant' boundaries): the fully unrolled instances
(k_stencil3d_fast) against the general marching kernel (k_stencil3d), with and without the epilogue operand (out = a*S(x) + b*y:
the residual A x - y of a deblurring data term).   python tools/bench_stencil3d.py [--n0 256 --n1 1024 --n2 1024] [--only fast]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K

ap = argparse.ArgumentParser()
ap.add_argument("--n0", type=int, default=256)
ap.add_argument("--n1", type=int, default=1024)
ap.add_argument("--n2", type=int, default=1024)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default=None)
ap.add_argument("--taps", default="3,5,7,9")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--dense-only", action="store_true", help="only the dense (full-rank) kernels")
ap.add_argument("--no-tiled", action="store_true", help="dense kernels: the marching kernel only")
args = ap.parse_args()
lib = K.lib()
shape = (args.n0, args.n1, args.n2)
N = int(np.prod(shape))
tdt, ndt, isz = (torch.float32, np.float32, 4) if args.dtype == "f32" else (torch.float64, np.float64, 8)


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return (k / k.sum()).astype(ndt)


def timeit(fn, reps):
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


x = torch.randn(1, N, device="cuda", dtype=tdt)
y = torch.randn(N, device="cuda", dtype=tdt)
rows = []
for k in [] if args.dense_only else [int(v) for v in args.taps.split(",")]:
    op = pxo.Stencil(arg_shape=shape, kernel=[gauss(k, 0.2 * k)] * 3, center=(k // 2,) * 3, mode="constant")
    for label, path in (("fast", 0), ("general", 1)):
        if args.only and label != args.only:
            continue
        K.check(lib.pxb_set_stencil3d_path(path), "path")
        for what, fn, bpv in (("apply", lambda: op.apply(x), 2 * isz), ("adjoint", lambda: op.adjoint(x), 2 * isz),
                              ("apply - y", lambda: op._run_tiled(x, False, alpha=1.0, beta=-1.0, add=y), 3 * isz)):
            ms = timeit(fn, args.reps)
            rows.append({"taps": k, "kernel": label, "what": what, "ms": ms, "bytes_per_voxel": bpv, "GBps": bpv * N / ms / 1e6})
            print(f"{k}x{k}x{k} {label:8s} {what:10s} {ms:8.3f} ms  {bpv * N / ms / 1e6:7.0f} GB/s ({bpv} B/voxel)", flush=True)
    lib.pxb_set_stencil3d_path(0)
# a dense PSF of full rank (not an outer product): the marching kernel (K^3 FMAs per sample, one pass), one tiled dense 2-D pass per kernel
# plane (Stencil._run_dense3d), the gather kernel
for k in [int(v) for v in args.taps.split(",") if int(v) <= 7]:
    kern = np.random.default_rng(k).random((k, k, k)).astype(ndt)
    for label, force in (("march", None), ("tiled", None), ("gather", False)):
        if (args.only and label not in ("tiled", "march")) or (args.no_tiled and label == "tiled"):
            continue
        from pyxu_b200.operator.linop import stencil as st_mod

        st_mod.DENSE3D_MARCH = label == "march"  # one marching pass (pxb_stencil3d_dense_apply) / one tiled 2-D pass per kernel plane
        op = pxo.Stencil(arg_shape=shape, kernel=kern, center=(k // 2,) * 3, mode="constant")
        op._dense3d_ok = force
        for what, fn in (("apply", lambda: op.apply(x)), ("adjoint", lambda: op.adjoint(x)), ("apply - y", lambda: op._run_tiled(x, False, alpha=1.0, beta=-1.0, add=y))):
            if label == "gather" and what != "apply":
                continue
            ms = timeit(fn, 2 if label == "gather" else args.reps)
            rows.append({"taps": k, "kernel": f"dense {label}", "what": what, "ms": ms, "tflops": 2 * k**3 * N / ms / 1e9,
                         "GBps": (3 if what == "apply - y" else 2) * isz * N / ms / 1e6})
            print(f"{k}x{k}x{k} dense (full rank) {label:7s} {what:10s} {ms:9.3f} ms  {2 * k**3 * N / ms / 1e9:6.2f} TFLOP/s  {(3 if what == 'apply - y' else 2) * isz * N / ms / 1e6:7.0f} GB/s",
                  flush=True)
        if label == "march":
            assert op._march3d_ok is True, "the marching kernel declined"
print(json.dumps({"shape": f"{shape} {args.dtype}", "rows": rows}))
