#!/usr/bin/env python
"""How the single-kernel PD3O-TV iteration scales with the number of planes on ONE GPU: (n0, 1024, 1024) volumes for several n0.
A rank of an N-GPU run of the 1024^3 headline owns 1024/N planes; what this prints for n0 = 1024/N is the per-rank kernel time
with no neighbour at all -- the part of the N-GPU step time that is not communication (launch, ramp-up, tail, fixed costs).

    python tools/bench_slab_shape.py [--n1 1024 --n2 1024] [--reps 20]
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
ap.add_argument("--n1", type=int, default=1024)
ap.add_argument("--n2", type=int, default=1024)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--planes", default="64,128,256,512,1024")
ap.add_argument("--spin", type=int, default=0, help="extra untimed pairs first (keeps the GPUs of a box busy together before the timed ones)")
args = ap.parse_args()
# under torchrun: every rank times the same kernel on its own GPU AT THE SAME TIME, no communication -- how much the GPUs of one box
# differ under simultaneous load is the floor of what a ring of coupled ranks can run at
rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
lib = K.lib()
out = []
for n0 in [int(v) for v in args.planes.split(",")]:
    shape = (n0, args.n1, args.n2)
    N = int(np.prod(shape))
    y = torch.rand(N, device="cuda")
    shift = -y
    P = K.PdsParams()
    P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
    P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
    f = K.FTerm()
    f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
    P.f = f
    P.hkind, P.lam = K.DUAL_L21, 0.08
    u0, u1 = y.clone(), torch.empty_like(y)
    z0, z1 = torch.zeros(3 * N, device="cuda"), torch.empty(3 * N, device="cuda")
    d = pxo.Gradient(arg_shape=shape, dtype=np.float32)._desc(1, K.F32)

    def pair():
        K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, None, None), "iter")
        K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u1.data_ptr(), z1.data_ptr(), u0.data_ptr(), z0.data_ptr(), None, None, None, None), "iter")

    for _ in range(3 + args.spin):
        pair()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        pair()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (2 * args.reps)
    out.append({"planes": n0, "ms": ms, "us_per_plane": 1e3 * ms / n0, "GBps": 36 * N / ms / 1e6})
    del y, shift, u0, u1, z0, z1
    torch.cuda.empty_cache()
print(json.dumps({"shape": f"(n0, {args.n1}, {args.n2}) fp32, single-kernel PD3O-TV iteration, one GPU", "rank": rank, "rows": out}), flush=True)
