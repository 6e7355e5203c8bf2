"""
CPU checks of the TMA-tiled 2-D stencil (pyxu_b200/csrc/pxb_stencil_tma.cuh): the device's per-thread bodies, with the
box load emulated as a zero-filled gather, against (a) the generic per-sample bodies, (b) fixtures from the real
reference for the 'constant'-mode cases, (c) the adjoint identity <Sx, y> == <x, S^T y>.
"""
import numpy as np
import pytest

import cases
import emu_util as E
import pyxu_b200.operator as pxo
from conftest import golden


def relerr(a, b):
    return float(np.linalg.norm((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).ravel()) / max(np.linalg.norm(np.asarray(b, dtype=np.float64).ravel()), 1e-300))


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return k / k.sum()


CASES = [
    # (arg_shape, kernel, center)
    ((70, 300), [gauss(9, 1.7), gauss(9, 1.7)], (4, 4)),            # separable 9x9 (config[1] blur), ragged tiles
    ((70, 300), np.outer(gauss(9, 1.7), gauss(9, 1.7)), (4, 4)),    # the same, dense
    ((33, 132), np.arange(1.0, 26.0).reshape(5, 5) / 10, (1, 3)),   # dense 5x5, off-centre (config[2] PSF shape)
    ((40, 64), [np.r_[1.0, -2, 1], np.r_[0.5, 0.25, 3.0, 1.0]], (0, 3)),
    ((3, 37, 72), [np.r_[1.0, 2.0, -1.0], gauss(7, 1.2), gauss(7, 1.2)], (1, 3, 3)),   # 3-D separable 7x7 in-plane + 3 taps along axis 0
    ((4, 37, 72), np.arange(1.0, 10.0).reshape(1, 3, 3), (0, 1, 1)),                    # dense 2-D kernel on a stack of planes
    ((132,), np.r_[1.0, 2, -3, 0.5, 7], (2,)),                                          # 1-D
    ((45, 48), [np.r_[2.0], gauss(13, 2.0)], (0, 4)),                                   # 13 column taps (fp32 window limit), scalar row factor
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_tiled_equals_generic(ci, dtype):
    shape, kern, cen = CASES[ci]
    k = [np.asarray(_, dtype=dtype) for _ in kern] if isinstance(kern, list) else np.asarray(kern, dtype=dtype)
    op = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode="constant")
    rng = np.random.default_rng(ci)
    x = rng.standard_normal((2, op.dim)).astype(dtype)
    tol = 1e-13 if dtype == np.float64 else 3e-6
    for adj in (False, True):
        ref = E.stencil_run(op, x, adj)
        if dtype == np.float64 and isinstance(kern, list) and max(np.size(_) for _ in kern) > 11:
            assert E.stencil_run_tiled(op, x, adj) is None  # 13 taps: outside the fp64 window -> generic kernel
            continue
        out = E.stencil_run_tiled(op, x, adj)
        assert out is not None and relerr(out, ref) < tol, (ci, adj, relerr(out, ref))


def test_tiled_epilogue_and_adjoint_identity():
    rng = np.random.default_rng(3)
    op = pxo.Stencil(arg_shape=(50, 140), kernel=[gauss(9, 1.7), gauss(5, 1.0)], center=(4, 1), mode="constant")
    x, y = rng.standard_normal(op.dim), rng.standard_normal(op.dim)
    lhs = np.dot(E.stencil_run_tiled(op, x, False), y)
    rhs = np.dot(x, E.stencil_run_tiled(op, y, True))
    assert abs(lhs - rhs) < 1e-10 * (1 + abs(lhs))
    # out = alpha * S x + beta * add  (the "A x - y" of a data term), add broadcast over a stack of 3 images
    xs = rng.standard_normal((3, op.dim))
    out = E.stencil_run_tiled(op, xs, False, alpha=0.5, beta=-1.0, add=y)
    assert relerr(out, 0.5 * E.stencil_run(op, xs, False) - y) < 1e-13


def test_tiled_not_applicable():
    op = pxo.Stencil(arg_shape=(20, 24), kernel=np.ones((3, 3)), center=(1, 1), mode="reflect")
    assert op._tiled_plan(False) is None
    op = pxo.Stencil(arg_shape=(6, 20, 24), kernel=np.ones((3, 3, 3)), center=(1, 1, 1), mode="constant")
    assert op._tiled_plan(False) is None  # dense 3-D kernels keep the generic path
    op = pxo.Stencil(arg_shape=(20, 22), kernel=np.ones((3, 3), dtype=np.float32), center=(1, 1), mode="constant")
    x = np.zeros(op.dim, dtype=np.float32)
    assert E.stencil_run_tiled(op, x, False) is None  # fp32: the last axis must be a multiple of 4 samples
