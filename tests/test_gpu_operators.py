"""
GPU parity tests of the operators through the public API (=> through the C ABI):
Stencil / Convolve / Gradient apply + adjoint and the proximal maps, against fixtures produced by the
real reference, the CPU oracle on fresh seeded inputs, and size-independent properties at large sizes.
Tolerances: fp64 rel. L2 <= 1e-12 (single operator application); fp32 <= 5e-6.
"""
import types

import numpy as np
import pytest

import cases
from conftest import golden
from oracle import pyxu_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def px():
    import pyxu_b200.operator as pxo

    assert torch.cuda.is_available()
    return types.SimpleNamespace(operator=pxo)


def relerr(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    d = np.linalg.norm(np.asarray(b, dtype=np.float64).ravel())
    return np.linalg.norm((a.astype(np.float64) - b).ravel()) / (d if d else 1.0)


@pytest.mark.parametrize("case", cases.STENCIL_CASES, ids=lambda c: c["name"])
def test_stencil_golden(px, case):
    g, n = golden("stencil.npz"), case["name"]
    op = cases.make_stencil(px, case)
    out = op.apply(g[f"{n}/x"])
    assert isinstance(out, np.ndarray) and out.dtype == np.float64  # NumPy in -> NumPy out
    assert relerr(out, g[f"{n}/apply"]) < 1e-12
    assert relerr(op.adjoint(g[f"{n}/y"]), g[f"{n}/adjoint"]) < 1e-12
    # device buffers in -> device buffers out (zero-copy path), 1-D input
    xd = torch.from_numpy(g[f"{n}/x"][0]).cuda()
    yd = op(xd)
    assert yd.is_cuda and relerr(yd, g[f"{n}/apply"][0]) < 1e-12
    # fp32
    op32 = cases.make_stencil(px, case, dtype=np.float32)
    assert relerr(op32.apply(g[f"{n}/x"].astype(np.float32)), g[f"{n}/apply"]) < 5e-6
    assert relerr(op32.adjoint(g[f"{n}/y"].astype(np.float32)), g[f"{n}/adjoint"]) < 5e-6


@pytest.mark.parametrize("case", cases.GRADIENT_CASES, ids=lambda c: c["name"])
def test_gradient_golden(px, case):
    g, n = golden("gradient.npz"), case["name"]
    op = cases.make_gradient(px, case)
    assert relerr(op.apply(g[f"{n}/x"]), g[f"{n}/apply"]) < 1e-12
    assert relerr(op.adjoint(g[f"{n}/y"]), g[f"{n}/adjoint"]) < 1e-12
    assert relerr(op.apply(g[f"{n}/x"].astype(np.float32)), g[f"{n}/apply"]) < 5e-6
    assert abs(op.lipschitz - float(g[f"{n}/lipschitz"])) <= 1e-12 * op.lipschitz


def test_dlpack_input(px):
    """Any __dlpack__ exporter on the device is accepted zero-copy."""

    class Foreign:  # stands for a CuPy array / another framework's buffer
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None):
            return self._t.__dlpack__()

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    op = px.operator.Gradient(arg_shape=(10, 11))
    x = np.random.default_rng(0).standard_normal(110)
    out = op.apply(Foreign(torch.from_numpy(x).cuda()))
    assert out.is_cuda and relerr(out, orc.Gradient((10, 11)).apply(x)) < 1e-13


@pytest.mark.parametrize("shape,ks,modes", [
    ((37, 53), (5, 5), ("reflect", "wrap")),
    ((19, 23, 17), (3, 4, 5), ("symmetric", "edge", "constant")),
    ((301,), (9,), ("wrap",)),
    ((33, 65), (9, 9), ("constant", "constant")),
    ((12, 20, 36), (7, 7, 7), ("constant", "reflect", "symmetric")),
])
def test_stencil_vs_oracle_random(px, shape, ks, modes):
    rng = np.random.default_rng(abs(hash((shape, ks))) % 2**31)
    kern = rng.standard_normal(ks)
    cen = tuple(int(rng.integers(0, k)) for k in ks)
    op = px.operator.Stencil(arg_shape=shape, kernel=kern, center=cen, mode=modes)
    ref = orc.Stencil(shape, kern, cen, modes)
    x = rng.standard_normal((3, op.dim))
    assert relerr(op.apply(x), ref.apply(x)) < 1e-12
    assert relerr(op.adjoint(x), ref.adjoint(x)) < 1e-12


def test_funcs_golden(px):
    pxo = px.operator
    g = golden("funcs.npz")
    x = g["x"]
    N = 60
    for tau in (0.3, 1.7):
        t = f"{tau}"
        chk = lambda got, key: relerr(got, g[key]) < 1e-14 or pytest.fail(key)
        chk(pxo.L1Norm(dim=N).prox(x, tau), f"l1/prox/{t}")
        chk((0.4 * pxo.L1Norm(dim=N)).prox(x, tau), f"l1s/prox/{t}")
        chk(pxo.L1Norm(dim=N).fenchel_prox(x, tau), f"l1/fprox/{t}")
        chk(pxo.PositiveL1Norm(dim=N).prox(x, tau), f"posl1/prox/{t}")
        chk(pxo.PositiveOrthant(dim=N).prox(x, tau), f"pos/prox/{t}")
        chk(pxo.LInfinityBall(dim=N, radius=0.8).prox(x, tau), f"linfball/prox/{t}")
        chk(pxo.SquaredL2Norm(dim=N).prox(x, tau), f"sql2/prox/{t}")
        chk(pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(0,)).prox(x, tau), f"l21/prox/{t}")
        chk((0.7 * pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(0,))).fenchel_prox(x, tau), f"l21s/fprox/{t}")
        chk(pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(1, 2)).prox(x, tau), f"l21ax12/prox/{t}")
    assert relerr(pxo.L1Norm(dim=N).apply(x), g["l1/apply"]) < 1e-14
    assert relerr(pxo.L21Norm(arg_shape=(3, 4, 5)).apply(x), g["l21/apply"]) < 1e-14
    assert relerr(pxo.SquaredL2Norm(dim=N).apply(x), g["sql2/apply"]) < 1e-14
    assert relerr(pxo.SquaredL2Norm(dim=N).grad(x), g["sql2/grad"]) < 1e-14
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-x[0])
    assert relerr(f.grad(x), g["sql2shift/grad"]) < 1e-14
    assert relerr(f.apply(x), g["sql2shift/apply"]) < 1e-13
    assert f.diff_lipschitz == float(g["sql2shift/diff_lipschitz"])


def test_large_properties_fp32(px):
    """At sizes no CPU oracle finishes quickly: adjointness <Ax,y> = <x,A^T y>, linearity, constant nullspace."""
    pxo = px.operator
    shape = (256, 320, 384)  # 31 M voxels
    N = int(np.prod(shape))
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, device="cuda", dtype=torch.float32, generator=gen)
    for mode in ("constant", ("reflect", "wrap", "symmetric")):
        Kop = pxo.Gradient(arg_shape=shape, mode=mode)
        y = torch.randn(3 * N, device="cuda", dtype=torch.float32, generator=gen)
        lhs = torch.dot(Kop(x).double(), y.double()).item()
        rhs = torch.dot(x.double(), Kop.adjoint(y).double()).item()
        assert abs(lhs - rhs) < 1e-5 * (abs(lhs) + np.sqrt(3.0 * N))
        if mode != "constant":
            ones = torch.ones(N, device="cuda", dtype=torch.float32)
            assert float(Kop(ones).abs().max()) == 0.0  # derivative of a constant under non-zero boundary extension
    k = np.random.default_rng(1).standard_normal((3, 5, 5)).astype(np.float32)
    S = pxo.Stencil(arg_shape=shape, kernel=k, center=(1, 2, 2), mode=("edge", "reflect", "wrap"))
    y = torch.randn(N, device="cuda", dtype=torch.float32, generator=gen)
    lhs = torch.dot(S(x).double(), y.double()).item()
    rhs = torch.dot(x.double(), S.adjoint(y).double()).item()
    assert abs(lhs - rhs) < 1e-5 * (abs(lhs) + np.sqrt(float(N)) * np.abs(k).sum())
    a = S(2.0 * x + y)
    b = 2.0 * S(x) + S(y)
    assert float((a - b).abs().max()) < 1e-3
