#!/usr/bin/env python
"""Run-to-run and path-to-path agreement of the PD3O-TV solve on one GPU (256 x 512 x 512 fp32, K iterations, MaxIter | RelError[x]):
device-resident fit twice, streamed host-array fit twice; bitwise comparisons of x and z."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst

shape = tuple(int(v) for v in os.environ.get("SHAPE", "256,512,512").split(","))
N = int(np.prod(shape))
K = int(os.environ.get("K", 20))
y = torch.rand(N, device="cuda")
if os.environ.get("PHANTOM"):  # bench.py's phantom (blocks of a 16^3 random field + noise) instead of white noise
    env = bench.Env()
    y = bench.local_phantom(env, shape[0]).reshape(-1)
yh = y.cpu().numpy()
if os.environ.get("RESERVE"):
    from pyxu_b200 import _array as A_

    A_.reserve_host_results(4 * N)
PDS = pxs.PD3O.__mro__[1]
crit = lambda: pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x")
res = {}
for name, floor, x0, sh in (("dev1", 1 << 62, y, -y), ("dev2", 1 << 62, y, -y), ("str1", 1, yh, -yh), ("str2", 1, yh, -yh)):
    PDS._STREAM_MIN_BYTES = floor
    slv = bench.tv_solver(shape, sh)
    slv.fit(x0=x0, stop_crit=crit())
    d, h = slv.stats()
    res[name] = (torch.as_tensor(d["x"]).cpu().numpy().copy(), torch.as_tensor(d["z"]).cpu().numpy().copy(), slv._slab is not None)
    del slv
for a, b in (("dev1", "dev2"), ("str1", "str2"), ("dev1", "str1")):
    xa, za, sa = res[a]
    xb, zb, sb = res[b]
    dx, dz = np.abs(xa - xb), np.abs(za - zb)
    print(f"{a} vs {b}: streamed {sa}/{sb}  x: {int((dx > 0).sum())} samples differ, max {dx.max():.3e}, rel L2 {np.linalg.norm(dx) / np.linalg.norm(xb):.3e};  "
          f"z: {int((dz > 0).sum())} differ, max {dz.max():.3e}")
    if (dx > 0).any():
        idx = np.argwhere(dx.reshape(shape) > 0)
        print("   planes with differences:", np.unique(idx[:, 0])[:40], " rows:", np.unique(idx[:, 1])[:12], " cols:", np.unique(idx[:, 2])[:12])
