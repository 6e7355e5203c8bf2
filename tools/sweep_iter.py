#!/usr/bin/env python
"""Times the single-kernel PD3O-TV iteration (pxb_pds_iter_chunked) over chunk lengths / sizes on one GPU.
Usage: python tools/sweep_iter.py [--size 1024] [--chunks 16,32,64,128] [--reps 10] [--path 0|1|2]"""
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
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--chunks", default="16,32,64,128,256")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--path", type=int, default=0)
ap.add_argument("--write-x", action="store_true")
args = ap.parse_args()
n = args.size
lib = K.lib()
lib.pxb_set_iter_path(args.path)
Kop = pxo.Gradient(arg_shape=(n, n, n), dtype=np.float32)
d = Kop._desc(1, K.F32)
dev = "cuda"
y = torch.rand(n**3, device=dev)
u0, u1 = y.clone(), torch.empty_like(y)
z0, z1 = torch.zeros(3 * n**3, device=dev), torch.empty(3 * n**3, device=dev)
x = torch.empty_like(y) if args.write_x else None
shift = -y
P = K.PdsParams()
P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
f = K.FTerm()
f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
P.f = f
P.hkind, P.lam = K.DUAL_L21, 0.08
for chunk in [int(c) for c in args.chunks.split(",")]:
    def step(a, b, c, e):
        rc = lib.pxb_pds_iter_chunked(K.ALGO_PD3O, C.byref(d), C.byref(P), a.data_ptr(), b.data_ptr(), c.data_ptr(), e.data_ptr(),
                                      x.data_ptr() if x is not None else None, None, None, chunk, None)
        K.check(rc, "iter")
    for _ in range(3):
        step(u0, z0, u1, z1); step(u1, z1, u0, z0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step(u0, z0, u1, z1); step(u1, z1, u0, z0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (2 * args.reps)
    bpv = 40 if args.write_x else 36
    print(f"size {n} chunk {chunk:4d}: {ms:.3f} ms  {n**3/ms/1e6:.1f} Gvox/s  {bpv*n**3/ms/1e6:.0f} GB/s", flush=True)
