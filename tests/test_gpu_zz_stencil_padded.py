"""GPU parity of Stencil.apply / adjoint with FOLDING boundary modes through the tiled kernel (Stencil._run_padded:
pxb_pad2d -> pxb_stencil2d_apply on the padded array / pxb_stencil2d_apply onto the padded extent -> pxb_pad2d_adjoint)
against the gather kernels (pinned on the real reference by the golden-vector tests), the reference's fixtures, the adjoint
identity and the solver path that uses it (CondatVu deblurring with a reflect-mode blur).

The path is the default (PYXU_B200_STENCIL_PADDED=0 switches it off); the tests pin it on so that they compare what they
say they compare whatever the environment holds."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from conftest import golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
DEV = "cuda"  # tests/test_emu_device_solvers.py replays these functions on the emulated device with DEV = "cpu"


@pytest.fixture(autouse=True)
def _padded_on(monkeypatch):
    from pyxu_b200.operator.linop import stencil as st

    monkeypatch.setattr(st, "PADDED_TILED", True)


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return k / k.sum()


CASES = [
    ((517, 1028), [gauss(9, 1.7), gauss(9, 1.7)], (4, 4), "reflect"),
    ((517, 1028), np.outer(gauss(9, 1.7), gauss(9, 1.7)) + 0.01 * np.arange(81.0).reshape(9, 9), (4, 4), "symmetric"),
    ((333, 260), np.arange(1.0, 26.0).reshape(5, 5) / 10, (1, 3), ("wrap", "edge")),
    ((333, 260), [gauss(5, 1.0), gauss(7, 1.3)], (0, 6), ("constant", "wrap")),
    ((5, 130, 264), [np.r_[1.0, 2.0, -1.0], gauss(7, 1.2), gauss(7, 1.2)], (1, 3, 3), ("reflect", "symmetric", "wrap")),
    ((4, 130, 264), np.arange(1.0, 10.0).reshape(1, 3, 3), (0, 1, 1), ("constant", "edge", "reflect")),
    ((4100,), np.r_[1.0, 2, -3, 0.5, 7], (2,), "wrap"),
]


def rel(a, b):
    return float(torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_padded_vs_generic(ci, dtype):
    import pyxu_b200.operator as pxo
    from pyxu_b200.operator.linop import stencil as st

    assert st.PADDED_TILED
    shape, kern, cen, mode = CASES[ci]
    k = [np.asarray(_, dtype=dtype) for _ in kern] if isinstance(kern, list) else np.asarray(kern, dtype=dtype)
    op = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode=mode)
    generic = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode=mode)
    generic._padded_ok = False
    gen = torch.Generator(device=DEV).manual_seed(ci)
    x = torch.randn(3, op.dim, device=DEV, dtype=torch.float64 if dtype == np.float64 else torch.float32, generator=gen)
    y = torch.randn(3, op.dim, device=DEV, dtype=x.dtype, generator=gen)
    tol = 1e-13 if dtype == np.float64 else 3e-6
    for adj in (False, True):
        a = op.adjoint(x) if adj else op.apply(x)
        b = generic.adjoint(x) if adj else generic.apply(x)
        assert op._padded_ok is True, "the padded tiled path did not run"
        assert torch.isfinite(a).all() and rel(a, b) < tol, (ci, adj, rel(a, b))
    lhs, rhs = torch.sum(op.apply(x).double() * y.double()), torch.sum(x.double() * op.adjoint(y).double())
    assert abs(float(lhs - rhs)) < (1e-10 if dtype == np.float64 else 2e-4) * (1 + abs(float(lhs)))


@pytest.mark.parametrize("case", [c for c in cases.STENCIL_CASES if c["mode"] != "constant" and c["arg_shape"][-1] % 2 == 0],
                         ids=lambda c: c["name"])
def test_padded_golden(case):
    import types

    import pyxu_b200.operator as pxo

    g = golden("stencil.npz")
    n = case["name"]
    op = cases.make_stencil(types.SimpleNamespace(operator=pxo), case)
    a, b = op.apply(g[f"{n}/x"]), op.adjoint(g[f"{n}/y"])
    err = lambda u, v: np.linalg.norm(np.asarray(u) - v) / np.linalg.norm(v)
    assert err(a, g[f"{n}/apply"]) < 1e-13 and err(b, g[f"{n}/adjoint"]) < 1e-13


def test_cv_deblur_reflect_blur_uses_the_padded_path():
    """CondatVu deblurring with a reflect-mode 9x9 blur: grad f = A^T(Ax - y) runs through the padded tiled passes; same
    iterates as with the gather kernels (which test_gpu_solvers.py pins on the reference's fixtures)."""
    import types

    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    px = types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst)
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"].reshape(32, 40)
    g9 = gauss(9, 1.7)
    res = []
    for padded in (True, False):
        slv, Aop = cases.build_tv_deblur(px, y, (32, 40), np.outer(g9, g9), (4, 4), lam=0.02, blur_mode="reflect", positivity=True)
        if not padded:
            Aop._padded_ok = False
        slv.fit(x0=np.zeros(y.size), stop_crit=px.stop.MaxIter(30))
        assert slv._astate.get("error") is None, slv._astate.get("error")
        assert (Aop._padded_ok is True) == padded
        res.append(slv.stats()[0]["x"])
    assert np.linalg.norm(res[0] - res[1]) / np.linalg.norm(res[1]) < 1e-11
