#!/usr/bin/env python
"""A/B of one PD3O-TV iteration with a folding boundary mode: single-kernel form (MODES instances of pxb_pds_iter)
against the two-sweep form (pxb_pds_primal + pxb_pds_dual), and against the 'constant' single-kernel instance.
Usage: python tools/bench_modes.py [--size 512] [--size2d 8192] [--reps 10]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--size2d", type=int, default=8192)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--modes", default="constant,reflect,wrap,symmetric,edge")
ap.add_argument("--no-two-sweep", action="store_true")
args = ap.parse_args()
lib = K.lib()


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


for shape in [sh for sh in ((args.size,) * 3, (args.size2d,) * 2) if sh[0] > 0]:
    N, D = int(np.prod(shape)), len(shape)
    y = torch.rand(N, device="cuda")
    shift = -y
    P = K.PdsParams()
    P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
    P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
    f = K.FTerm()
    f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
    P.f = f
    P.hkind, P.lam = K.DUAL_L21, 0.08
    u0, u1, w, x = y.clone(), torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
    z0, z1 = torch.zeros(D * N, device="cuda"), torch.empty(D * N, device="cuda")
    bpv = 4 * (2 * D + 3)
    for mode in args.modes.split(","):
        d = pxo.Gradient(arg_shape=shape, mode=mode, dtype=np.float32)._desc(1, K.F32)

        def one():
            K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, None, None), "iter")
            K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u1.data_ptr(), z1.data_ptr(), u0.data_ptr(), z0.data_ptr(), None, None, None, None), "iter")

        def two():
            for _ in range(2):
                K.check(lib.pxb_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), None, x.data_ptr(), w.data_ptr(), None, None), "primal")
                K.check(lib.pxb_pds_dual(C.byref(d), C.byref(P), w.data_ptr(), z0.data_ptr(), None, None), "dual")

        t1 = timeit(one, args.reps) / 2
        t2 = float("nan") if args.no_two_sweep else timeit(two, args.reps) / 2
        print(f"{shape} fp32 mode={mode:9s}: single kernel {t1:7.3f} ms ({bpv * N / t1 / 1e6:5.0f} GB/s at {bpv} B/voxel)   two sweeps {t2:7.3f} ms", flush=True)
    del u0, u1, w, x, z0, z1, y, shift
