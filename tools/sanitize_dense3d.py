#!/usr/bin/env python
"""Small launches of every instance of the dense marching stencil kernel (K = 3, 5, 7; fp32 / fp64; ragged tiles, several chunks,
epilogue operand) for compute-sanitizer:    compute-sanitizer --tool memcheck python tools/sanitize_dense3d.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200.operator.linop import stencil as st_mod

st_mod.DENSE3D_MARCH = True
rng = np.random.default_rng(0)
for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    for shape, ks, cen in (((21, 19, 140), (7, 7, 7), (3, 3, 3)), ((9, 35, 264), (7, 6, 7), (0, 5, 1)), ((12, 17, 72), (5, 5, 5), (2, 2, 2)), ((6, 20, 24), (3, 3, 3), (1, 1, 1))):
        op = pxo.Stencil(arg_shape=shape, kernel=rng.standard_normal(ks).astype(dt), center=cen, mode="constant")
        x = torch.randn(2, op.dim, device="cuda", dtype=tdt)
        y = torch.randn(op.dim, device="cuda", dtype=tdt)
        op.apply(x), op.adjoint(x), op._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y)
        assert op._march3d_ok is True
torch.cuda.synchronize()
print("dense3d: done")
