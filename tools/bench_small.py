#!/usr/bin/env python
"""configs[0]-sized problems (512^2 fp64 PD3O-TV, the state lives in L2): n iterations through pxb_pds_iter_n -- one cooperative
launch with grid-wide barriers (k_tv_tile2d_loop) when every tile has a resident CTA -- against n single launches, and where the time of
a host-array fit() + solution() goes.   python tools/bench_small.py [--n 512] [--iters 200]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst
from pyxu_b200 import _cabi as K

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--dtype", default="f64")
args = ap.parse_args()
lib = K.lib()
n, it = args.n, args.iters
shape, N = (n, n), n * n
tdt, ndt, kdt = (torch.float64, np.float64, K.F64) if args.dtype == "f64" else (torch.float32, np.float32, K.F32)
y = torch.rand(N, device="cuda", dtype=tdt)
shift = -y
P = K.PdsParams()
P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
f = K.FTerm()
f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
P.f = f
P.hkind, P.lam = K.DUAL_L21, 0.1
d = pxo.Gradient(arg_shape=shape, dtype=ndt)._desc(1, kdt)
out = {}


def state():
    return y.clone(), torch.empty_like(y), torch.zeros(2 * N, device="cuda", dtype=tdt), torch.empty(2 * N, device="cuda", dtype=tdt)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


u0, u1, z0, z1 = state()
c0 = lib.pxb_launch_count()
K.check(lib.pxb_pds_iter_n(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, it, None, None, None), "iter_n")
out["launches_of_one_iter_n_call"] = int(lib.pxb_launch_count() - c0)
ms_n = timed(lambda: K.check(lib.pxb_pds_iter_n(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, it, None, None, None), "iter_n"))
ref_a = u0.clone()


def singles():
    a, b = (u0, z0), (u1, z1)
    for _ in range(it):
        K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), a[0].data_ptr(), a[1].data_ptr(), b[0].data_ptr(), b[1].data_ptr(), None, None, None, None), "iter")
        a, b = b, a


ms_1 = timed(singles)
out.update(iter_n_ms=ms_n, iter_n_us_per_iter=1e3 * ms_n / it, singles_ms=ms_1, singles_us_per_iter=1e3 * ms_1 / it)
# the two forms from the same start
u0, u1, z0, z1 = state()
K.check(lib.pxb_pds_iter_n(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, it, None, None, None), "iter_n")
ra = (u0 if it % 2 == 0 else u1).clone()
u0, u1, z0, z1 = state()
singles()
rb = (u0 if it % 2 == 0 else u1).clone()
out["bit_identical"] = bool(torch.equal(ra, rb))
# host-array fit + solution
yh = y.cpu().numpy()


def solve(crit):
    fobj = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yh)
    slv = pxs.PD3O(f=fobj, g=pxo.PositiveOrthant(dim=N), h=0.1 * pxo.L21Norm(arg_shape=(2, *shape), l2_axis=(0,)), K=pxo.Gradient(arg_shape=shape, dtype=ndt),
                   show_progress=False, final_writeback=False)
    t0 = time.perf_counter()
    slv.fit(x0=yh.copy(), stop_crit=crit())
    t1 = time.perf_counter()
    x = slv.solution()
    t2 = time.perf_counter()
    return dict(fit_ms=1e3 * (t1 - t0), solution_ms=1e3 * (t2 - t1), timing=slv._astate["timing"], n_hist=len(slv.stats()[1]))


for name, crit in (("maxiter", lambda: pxst.MaxIter(it)), ("maxiter_or_relerr", lambda: pxst.MaxIter(it) | pxst.RelError(eps=1e-30, var="x"))):
    solve(crit)
    out["fit_" + name] = solve(crit)
print(json.dumps(out))
if os.environ.get("PXB_PROFILE_HOST"):
    import cProfile
    import pstats

    def whole():
        fobj = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yh)
        slv = pxs.PD3O(f=fobj, g=pxo.PositiveOrthant(dim=N), h=0.1 * pxo.L21Norm(arg_shape=(2, *shape), l2_axis=(0,)), K=pxo.Gradient(arg_shape=shape, dtype=ndt),
                       show_progress=False, final_writeback=False)
        slv.fit(x0=yh.copy(), stop_crit=pxst.MaxIter(it))
        return slv.solution()

    whole()
    t0 = time.perf_counter()
    whole()
    print("whole solve ms:", 1e3 * (time.perf_counter() - t0))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        whole()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
