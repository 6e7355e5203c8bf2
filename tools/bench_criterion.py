#!/usr/bin/env python
"""What the default stopping criterion costs inside the single-kernel PD3O-TV iteration on ONE GPU, leg by leg:

    plain        read u, y, z0..2; write u, z0..2                                   36 B/voxel
    +x           ... and write x                                                    40 B/voxel
    +x+normz     ... and the RelError[z] sums (no extra traffic)                    40 B/voxel
    +x+norms     ... and the RelError[x] sums: the previous x is read back          44 B/voxel   (= Solver.fit() default)

    python tools/bench_criterion.py [--n0 1024 --n1 1024 --n2 1024] [--reps 10] [--only +x+norms]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K

ap = argparse.ArgumentParser()
ap.add_argument("--n0", type=int, default=1024)
ap.add_argument("--n1", type=int, default=1024)
ap.add_argument("--n2", type=int, default=1024)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default=None)
ap.add_argument("--dtype", default="f32")
args = ap.parse_args()
lib = K.lib()
shape = (args.n0, args.n1, args.n2)
N = int(np.prod(shape))
tdt, ndt, kdt, isz = (torch.float32, np.float32, K.F32, 4) if args.dtype == "f32" else (torch.float64, np.float64, K.F64, 8)
y = torch.rand(N, device="cuda", dtype=tdt)
shift = -y
P = K.PdsParams()
P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
f = K.FTerm()
f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
P.f = f
P.hkind, P.lam = K.DUAL_L21, 0.08
u0, u1 = y.clone(), torch.empty_like(y)
z0, z1 = torch.zeros(3 * N, device="cuda", dtype=tdt), torch.empty(3 * N, device="cuda", dtype=tdt)
x = y.clone()
nx, nz = torch.zeros(2, device="cuda", dtype=torch.float64), torch.zeros(2, device="cuda", dtype=torch.float64)
d = pxo.Gradient(arg_shape=shape, dtype=ndt)._desc(1, kdt)
legs = {
    "plain": (None, None, None, 9),
    "+x": (x.data_ptr(), None, None, 10),
    "+x+normz": (x.data_ptr(), None, nz.data_ptr(), 10),
    "+x+norms": (x.data_ptr(), nx.data_ptr(), nz.data_ptr(), 11),
}
rows = []
for name, (xp, nxp, nzp, words) in legs.items():
    if args.only and name != args.only:
        continue

    def pair():
        K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), xp, nxp, nzp, None), "iter")
        K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u1.data_ptr(), z1.data_ptr(), u0.data_ptr(), z0.data_ptr(), xp, nxp, nzp, None), "iter")

    for _ in range(2):
        pair()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        pair()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (2 * args.reps)
    rows.append({"leg": name, "ms": ms, "bytes_per_voxel": words * isz, "GBps": words * isz * N / ms / 1e6})
print(json.dumps({"shape": f"{shape} {args.dtype}, single-kernel PD3O-TV iteration, one GPU", "rows": rows}))
