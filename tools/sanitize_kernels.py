#!/usr/bin/env python
"""One small launch of every hand-written kernel family (for compute-sanitizer memcheck / racecheck runs):
    compute-sanitizer --tool memcheck python tools/sanitize_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst
from pyxu_b200 import _cabi as K


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return (k / k.sum()).astype(np.float32)


rng = np.random.default_rng(0)
lib = K.lib()
for path in (2, 1):  # TMA forms, then direct-load forms of the single-kernel iteration
    lib.pxb_set_iter_path(path)
    for shape in ((19, 21, 136), (37, 136)):  # ragged tiles, several chunks
        N, D = int(np.prod(shape)), len(shape)
        y = torch.rand(N, device="cuda")
        f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)
        Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
        h = 0.08 * pxo.L21Norm(arg_shape=(D, *shape), l2_axis=(0,))
        for klass in (pxs.PD3O, pxs.CondatVu):
            slv = klass(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False, final_writeback=False)
            slv.fit(x0=y, stop_crit=pxst.MaxIter(3) | pxst.RelError(eps=1e-30, var="x") | pxst.RelError(eps=1e-30, var="z"))
            assert slv._plan.iter_ok is True
            slv.solution()
lib.pxb_set_iter_path(0)
# tiled stencils: separable (pipelined), dense, 3-D single pass, axis-0 streaming, two-pass FISTA, CondatVu with a blur data term
x2 = torch.rand(3, 70 * 520, device="cuda")
for kern, cen in (([gauss(9, 1.7), gauss(9, 1.7)], (4, 4)), (np.outer(gauss(5, 1.0), gauss(5, 1.0)) + 0.02 * np.eye(5, dtype=np.float32), (1, 3))):
    op = pxo.Stencil(arg_shape=(70, 520), kernel=kern, center=cen)
    op.apply(x2), op.adjoint(x2)
    assert op._tiled_ok is True
x3 = torch.rand(2, 21 * 19 * 136, device="cuda")
for k0 in (7, 4):  # 7 taps: single-pass 3-D kernel; 4 taps: axis-0 streaming + tiled in-plane
    op = pxo.Stencil(arg_shape=(21, 19, 136), kernel=[rng.standard_normal(k0).astype(np.float32), gauss(5, 1.0), gauss(7, 1.2)], center=(1, 2, 3))
    op.apply(x3), op.adjoint(x3)
    assert op._tiled_ok is True
shape, B = (70, 136), 3
N = shape[0] * shape[1]
for psf in (np.outer(gauss(5, 1.0), gauss(5, 1.0)), np.outer(gauss(5, 1.0), gauss(5, 1.0)) + 0.02 * np.eye(5, dtype=np.float32)):
    A = pxo.Stencil(arg_shape=shape, kernel=psf, center=(2, 2))
    yb = torch.rand(B, N, device="cuda")
    slv = pxs.PGD(f=(0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yb)) * A, g=0.02 * pxo.L1Norm(dim=N), show_progress=False, final_writeback=False)
    slv.fit(x0=yb, tau=1.0 / float(A.lipschitz) ** 2, stop_crit=pxst.MaxIter(3) | pxst.RelError(eps=1e-30, var="x"))
    assert slv._fused is not None
A = pxo.Stencil(arg_shape=shape, kernel=[gauss(9, 1.7), gauss(9, 1.7)], center=(4, 4))
y1 = torch.rand(N, device="cuda")
slv = pxs.CondatVu(f=(0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y1)) * A, g=None, h=0.05 * pxo.L21Norm(arg_shape=(2, *shape), l2_axis=(0,)),
                   K=pxo.Gradient(arg_shape=shape, dtype=np.float32), beta=float(A.lipschitz) ** 2, show_progress=False, final_writeback=False)
slv.fit(x0=y1, stop_crit=pxst.MaxIter(3))
G = pxo.Gradient(arg_shape=(9, 21, 136), dtype=np.float32, mode=("reflect", "constant", "wrap"))
G.adjoint(G.apply(x3[0, : G.dim].contiguous()))
# folding boundary modes: MODES instances of the three single-kernel forms (full and ragged tiles, wrap = the global rim
# evaluator), the two-sweep kernels, and the Pad -> tiled stencil -> Pad^T path when it is switched on
for path in (2, 1):
    lib.pxb_set_iter_path(path)
    for shape, mode in (((19, 21, 136), ("reflect", "symmetric", "wrap")), ((16, 16, 256), "wrap"), ((37, 136), ("edge", "reflect")), ((32, 256), "wrap")):
        N, D = int(np.prod(shape)), len(shape)
        y = torch.rand(N, device="cuda")
        slv = pxs.PD3O(f=0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y), g=pxo.PositiveOrthant(dim=N), h=0.08 * pxo.L21Norm(arg_shape=(D, *shape), l2_axis=(0,)),
                       K=pxo.Gradient(arg_shape=shape, dtype=np.float32, mode=mode), show_progress=False, final_writeback=False)
        slv.fit(x0=y, stop_crit=pxst.MaxIter(3) | pxst.RelError(eps=1e-30, var="x") | pxst.RelError(eps=1e-30, var="z"))
        assert slv._plan.iter_ok is True
        slv.solution()
lib.pxb_set_iter_path(0)
lib.pxb_set_iter_modes(0)
slv = pxs.PD3O(f=0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y), g=None, h=0.08 * pxo.L21Norm(arg_shape=(D, *shape), l2_axis=(0,)),
               K=pxo.Gradient(arg_shape=shape, dtype=np.float32, mode="reflect"), show_progress=False, final_writeback=False)
slv.fit(x0=y, stop_crit=pxst.MaxIter(3))
assert slv._plan.iter_ok is False  # two sweeps
lib.pxb_set_iter_modes(-1)
from pyxu_b200.operator.linop import stencil as _st

for kern, cen in (([gauss(9, 1.7), gauss(9, 1.7)], (4, 4)), (np.outer(gauss(5, 1.0), gauss(5, 1.0)) + 0.02 * np.eye(5, dtype=np.float32), (1, 3))):
    op = pxo.Stencil(arg_shape=(70, 520), kernel=kern, center=cen, mode=("reflect", "wrap"))
    op.apply(x2), op.adjoint(x2)
    assert op._padded_ok is (True if _st.PADDED_TILED else None)
op = pxo.Stencil(arg_shape=(21, 19, 136), kernel=[gauss(5, 1.0), gauss(5, 1.0), gauss(7, 1.2)], center=(1, 2, 3), mode=("reflect", "wrap", "symmetric"))
op.apply(x3), op.adjoint(x3)  # (padded path: streaming axis-0 pass with the boundary map + Pad -> tiled stencil / -> Pad^T)
torch.cuda.synchronize()
print("sanitize_kernels: all launches completed")
