#!/usr/bin/env python
"""Is the single-kernel PD3O-TV iteration deterministic?  K iterations twice from the same state, bitwise comparison of (u, z), for
the TMA form (plain; with x + RelError sums), the direct-load form, and two volume sizes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K

lib = K.lib()
Kit = int(os.environ.get("K", 10))


def run(shape, path, norms, seed=0, use_x=None, use_nx=None, use_nz=None):
    use_x = norms if use_x is None else use_x
    use_nx = norms if use_nx is None else use_nx
    use_nz = norms if use_nz is None else use_nz
    N = int(np.prod(shape))
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = torch.rand(N, device="cuda", generator=g)
    shift = -y
    P = K.PdsParams()
    P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
    P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
    f = K.FTerm()
    f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
    P.f = f
    P.hkind, P.lam = K.DUAL_L21, 0.08
    d = pxo.Gradient(arg_shape=shape, dtype=np.float32)._desc(1, K.F32)
    lib.pxb_set_iter_path(path)
    outs = []
    for rep in range(2):
        u0, u1 = y.clone(), torch.zeros_like(y)
        z0, z1 = 0.01 * torch.randn(3 * N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)), torch.zeros(3 * N, device="cuda")
        x = y.clone()
        nx, nz = torch.zeros(2, device="cuda", dtype=torch.float64), torch.zeros(2, device="cuda", dtype=torch.float64)
        a, b = (u0, z0), (u1, z1)
        for _ in range(Kit):
            K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), a[0].data_ptr(), a[1].data_ptr(), b[0].data_ptr(), b[1].data_ptr(),
                                     x.data_ptr() if use_x else None, nx.data_ptr() if use_nx else None, nz.data_ptr() if use_nz else None, None), "iter")
            a, b = b, a
        torch.cuda.synchronize()
        outs.append((a[0].clone(), a[1].clone()))
        del u0, u1, z0, z1, x
    lib.pxb_set_iter_path(0)
    du = (outs[0][0] != outs[1][0])
    dz = (outs[0][1] != outs[1][1])
    nu, nzd = int(du.sum()), int(dz.sum())
    where = ""
    if nu:
        idx = torch.nonzero(du.reshape(shape))[:6].tolist()
        where = f" first differing (plane, row, col): {idx}"
    print(f"shape {shape} path {('auto(TMA)', 'direct', 'tma')[path]} x={use_x} nx={use_nx} nz={use_nz} K={Kit}: u differs in {nu} samples, z in {nzd}{where}", flush=True)


if os.environ.get("VARIANTS"):
    for rep in range(int(os.environ.get("REPS", 2))):
        for ux, unx, unz in ((True, False, False), (False, False, True), (True, False, True), (True, True, False), (True, True, True)):
            run((1024, 1024, 1024), 0, True, seed=rep, use_x=ux, use_nx=unx, use_nz=unz)
    sys.exit(0)
for shape in ((1024, 1024, 1024), (512, 1024, 1024)):
    for path, norms in ((0, False), (0, True), (1, False)):
        if path == 1 and shape[0] == 1024 and os.environ.get("SKIP_DIRECT_BIG"):
            continue
        run(shape, path, norms)
