"""
CPU checks of the single-kernel PD3O / CondatVu iteration (pyxu_b200/csrc/pxb_tv_iter.cuh).

The per-thread phases of the CUDA kernel are compiled for the host (tests/emu) and replayed CTA by CTA with the
shared-memory ring as a host array.  They must reproduce (a) the two-pass kernel bodies on random states for
every finite-difference scheme, tile / chunk / batch layout, and (b) the fixtures produced by the real reference.
"""
import ctypes as C

import numpy as np
import pytest

import emu_util as E
import pyxu_b200.operator as pxo
from conftest import golden
from pyxu_b200 import _cabi as K


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64).ravel() - np.asarray(b, dtype=np.float64).ravel()) / max(np.linalg.norm(np.asarray(b, dtype=np.float64).ravel()), 1e-300))


def two_pass(algo, d, P, u, z, x, nx=None, nz=None):
    """reference for the emulation: generic per-voxel bodies, primal pass then dual pass (in place)."""
    w = np.zeros_like(u)
    E.lib().emu_pds_primal(algo, C.byref(d), C.byref(P), E.p(u), E.p(z), None, E.p(x) if algo == K.ALGO_PD3O else None, E.p(w), E.p(nx))
    E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(z), E.p(nz))


def one_pass(algo, d, P, u, z, x, nx=None, nz=None, chunk=0, form="direct"):
    u2, z2 = np.full_like(u, np.nan), np.full_like(z, np.nan)
    fn = {"direct": E.lib().emu_tv_iter, "tma": E.lib().emu_tv_iter_tma, "tile2d": E.lib().emu_tv_tile2d}[form]
    rc = fn(algo, C.byref(d), C.byref(P), E.p(u), E.p(z), E.p(u2), E.p(z2), E.p(x) if algo == K.ALGO_PD3O else None,
                             E.p(nx), E.p(nz), chunk)
    assert rc == 0, rc
    return u2, z2


SCHEMES = ["forward", "backward", "central"]


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("algo", [K.ALGO_PD3O, K.ALGO_CV])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_equals_two_pass_3d(scheme, algo, dtype, form):
    rng = np.random.default_rng(5)
    vec = 2 if dtype == np.float64 else 4
    # > 1 tile along rows (8) and columns (32*vec), ragged in both, several chunks of planes
    shape = (7, 19, 32 * vec * 2 + 3 * vec)
    Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, sampling=(1.0, 0.5, 2.0))
    d = Kop._desc(1, E.dcode(np.zeros(1, dtype=dtype)))
    shift = rng.standard_normal(Kop.dim).astype(dtype)
    for hkind, gspec in ((K.DUAL_L21, (K.PROX_POS, 0.0, 0.0)), (K.DUAL_L1, (K.PROX_BOX, 0.2, 0.9))):
        P = E.pds_params(0.21, 0.19, 0.9, gspec=gspec, fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=hkind, lam=0.3)
        for chunk in (0, 3, 1):
            u = rng.standard_normal(Kop.dim).astype(dtype)
            x = rng.standard_normal(Kop.dim).astype(dtype)
            z = rng.standard_normal(Kop.codim).astype(dtype)
            ua, za, xa = u.copy(), z.copy(), x.copy()
            nxa, nza, nxb, nzb = np.zeros(2), np.zeros(2), np.zeros(2), np.zeros(2)
            two_pass(algo, d, P, ua, za, xa, nxa, nza)
            xb = x.copy()
            ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=chunk, form=form)
            tol = 1e-13 if dtype == np.float64 else 2e-6
            assert relerr(ub, ua) < tol and relerr(zb, za) < tol, (scheme, algo, chunk)
            if algo == K.ALGO_PD3O:
                assert relerr(xb, xa) < tol
            assert np.allclose(nxa, nxb, rtol=1e-5 if dtype == np.float32 else 1e-10) and np.allclose(nza, nzb, rtol=1e-5 if dtype == np.float32 else 1e-10)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("width", [40, 300, 1100])
@pytest.mark.parametrize("form", ["direct", "tile2d"])
def test_iter_equals_two_pass_2d_batched(scheme, width, form):
    """2-D images (marching along the rows), a batch of them, narrow and wide tile variants."""
    rng = np.random.default_rng(7)
    shape, batch = (13, width), 3
    Kop = pxo.Gradient(arg_shape=shape, scheme=scheme)
    d = Kop._desc(batch, K.F32)
    shift = rng.standard_normal(Kop.dim).astype(np.float32)  # broadcast over the batch (period = one image)
    P = E.pds_params(0.3, 0.25, 1.0, gspec=(K.PROX_L1, 0.05, 0.0), fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=0.2)
    u = rng.standard_normal((batch, Kop.dim)).astype(np.float32)
    x = rng.standard_normal((batch, Kop.dim)).astype(np.float32)
    z = rng.standard_normal((batch, Kop.codim)).astype(np.float32)
    for algo in (K.ALGO_PD3O, K.ALGO_CV):
        for chunk in (0, 4):
            ua, za, xa = u.copy(), z.copy(), x.copy()
            nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
            two_pass(algo, d, P, ua, za, xa, nxa, nza)
            xb = x.copy()
            ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=chunk, form=form)
            assert relerr(ub, ua) < 2e-6 and relerr(zb, za) < 2e-6
            assert np.allclose(nxa, nxb, rtol=1e-5) and np.allclose(nza, nzb, rtol=1e-5)


def test_iter_cv_gradarr_and_stacked_2d():
    """CondatVu with a precomputed grad f array; 2-D gradient over the last two axes of a 3-D arg_shape."""
    rng = np.random.default_rng(9)
    shape = (3, 11, 24)
    Kop = pxo.Gradient(arg_shape=shape, directions=(1, 2))
    d = Kop._desc(2, K.F64)
    garr = rng.standard_normal((2, Kop.dim))
    P = E.pds_params(0.3, 0.25, 0.8, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_GRADARR, garr=garr, hkind=K.DUAL_L21, lam=0.2)
    u, z = rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.codim))
    ua, za = u.copy(), z.copy()
    two_pass(K.ALGO_CV, d, P, ua, za, None)
    ub, zb = one_pass(K.ALGO_CV, d, P, u, z, None, chunk=5)
    assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13
    ub, zb = one_pass(K.ALGO_CV, d, P, u, z, None, form="tile2d")
    assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13


def test_iter_not_eligible():
    Kop = pxo.Gradient(arg_shape=(8, 16), mode="reflect")
    d = Kop._desc(1, K.F64)
    P = E.pds_params(0.3, 0.25, 1.0, hkind=K.DUAL_L21, lam=0.2)
    a = np.zeros(Kop.dim)
    z = np.zeros(Kop.codim)
    E.lib().emu_set_iter_modes(0)  # pxb_set_iter_modes(0): folding modes are declined (callers take the two-sweep form)
    try:
        for fn in (E.lib().emu_tv_iter, E.lib().emu_tv_tile2d):
            assert fn(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(a), E.p(z), E.p(a.copy()), E.p(z.copy()), None, None, None, 0) == -104
    finally:
        E.lib().emu_set_iter_modes(1)
    Kop = pxo.Gradient(arg_shape=(8, 15))  # last axis not a multiple of the vector width
    d = Kop._desc(1, K.F64)
    assert E.lib().emu_tv_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(a), E.p(z), E.p(a.copy()), E.p(z.copy()), None, None, None, 0) == -105


def _solve(name, algo, shape, n_iter, lam, gspec, y, x0, dtype=np.float64):
    g = golden("solvers.npz")
    tau, sigma, rho = (float(g[f"{name}/{k}"]) for k in ("tau", "sigma", "rho"))
    Kop = pxo.Gradient(arg_shape=shape)
    shift = np.ascontiguousarray(-y.reshape(-1), dtype=dtype)
    P = E.pds_params(tau, sigma, rho, gspec=gspec, fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=lam)
    d = Kop._desc(1, E.dcode(shift))
    x = np.ascontiguousarray(x0, dtype=dtype).copy()
    z = E.gradient_run(Kop, x, False)
    u = x.copy()
    nx, nz = np.zeros(2), np.zeros(2)
    hist = []
    for _ in range(n_iter):
        nx[:] = 0
        nz[:] = 0
        if algo == K.ALGO_PD3O:
            u, z = one_pass(algo, d, P, u, z, x, nx, nz)
        else:
            x, z = one_pass(algo, d, P, x, z, None, nx, nz)
        hist.append((nx.copy(), nz.copy()))
    return x, z, hist, g


POS = (K.PROX_POS, 0.0, 0.0)


def test_iter_golden_pd3o_2d_3d_cv():
    """N iterations of the single-kernel form against the real reference's NumPy float64 solver (<= 1e-10)."""
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _solve("pd3o_tv2d/s1", K.ALGO_PD3O, (32, 40), 60, 0.1, POS, y, y.reshape(-1))
    assert relerr(x, g["pd3o_tv2d/s1/x"]) < 1e-10 and relerr(z, g["pd3o_tv2d/s1/z"]) < 1e-10
    x, z, *_ = _solve("cv_tv2d", K.ALGO_CV, (32, 40), 60, 0.1, POS, y, y.reshape(-1))
    assert relerr(x, g["cv_tv2d/x"]) < 1e-10 and relerr(z, g["cv_tv2d/z"]) < 1e-10
    y3 = g["pd3o_tv3d/y"]
    x, z, hist, _ = _solve("pd3o_tv3d", K.ALGO_PD3O, (10, 12, 14), 50, 0.08, POS, y3, y3.reshape(-1))
    assert relerr(x, g["pd3o_tv3d/x"]) < 1e-10 and relerr(z, g["pd3o_tv3d/z"]) < 1e-10
    # fused RelError sums of the last iteration == recomputed from iterates 49 -> 50
    xa, za, *_ = _solve("pd3o_tv3d", K.ALGO_PD3O, (10, 12, 14), 49, 0.08, POS, y3, y3.reshape(-1))
    nx, nz = hist[-1]
    assert abs(nx[0] - np.sum((x - xa) ** 2)) < 1e-12 * (1 + nx[0]) and abs(nx[1] - np.sum(xa**2)) < 1e-9 * nx[1]
    assert abs(nz[0] - np.sum((z - za) ** 2)) < 1e-12 * (1 + nz[0]) and abs(nz[1] - np.sum(za**2)) < 1e-9 * nz[1]


def test_iter_golden_f32():
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _solve("pd3o_tv2d/s1", K.ALGO_PD3O, (32, 40), 60, 0.1, POS, y, y.reshape(-1), dtype=np.float32)
    assert relerr(x, g["pd3o_tv2d/s1/x"]) < 1e-4


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tma_form_batched_shift_modes_gradarr(scheme):
    """TMA-staged form: batch > 1 with a broadcast shift (one volume) and a per-item shift, CondatVu with grad f array."""
    rng = np.random.default_rng(11)
    shape, batch = (5, 11, 24), 2
    Kop = pxo.Gradient(arg_shape=shape, scheme=scheme)
    d = Kop._desc(batch, K.F64)
    u, x = rng.standard_normal((batch, Kop.dim)), rng.standard_normal((batch, Kop.dim))
    z = rng.standard_normal((batch, Kop.codim))
    for shift in (rng.standard_normal(Kop.dim), rng.standard_normal((batch, Kop.dim)), np.r_[0.3]):
        P = E.pds_params(0.21, 0.19, 0.9, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L21, lam=0.3)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            ua, za, xa = u.copy(), z.copy(), x.copy()
            nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
            two_pass(algo, d, P, ua, za, xa, nxa, nza)
            xb = x.copy()
            ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=2, form="tma")
            assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13, (scheme, algo, shift.shape)
            assert np.allclose(nxa, nxb, rtol=1e-10) and np.allclose(nza, nzb, rtol=1e-10)
    garr = rng.standard_normal((batch, Kop.dim))
    P = E.pds_params(0.3, 0.25, 0.8, gspec=(K.PROX_L1, 0.1, 0.0), fkind=K.F_GRADARR, garr=garr, hkind=K.DUAL_L1, lam=0.2)
    ua, za = u.copy(), z.copy()
    two_pass(K.ALGO_CV, d, P, ua, za, None)
    ub, zb = one_pass(K.ALGO_CV, d, P, u, z, None, chunk=0, form="tma")
    assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13
    for gk in (K.PROX_POS, K.PROX_NONE):  # the instances specialised for CondatVu + grad f array (forward scheme) / the generic one
        P = E.pds_params(0.3, 0.25, 0.8, gspec=(gk, 0.0, 0.0), fkind=K.F_GRADARR, garr=garr, hkind=K.DUAL_L21, lam=0.2)
        ua, za = u.copy(), z.copy()
        nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
        two_pass(K.ALGO_CV, d, P, ua, za, None, nxa, nza)
        ub, zb = one_pass(K.ALGO_CV, d, P, u, z, None, nxb, nzb, chunk=3, form="tma")
        assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13
        assert np.allclose(nxa, nxb, rtol=1e-10) and np.allclose(nza, nzb, rtol=1e-10)


def test_tma_form_golden_3d():
    g = golden("solvers.npz")
    y3 = g["pd3o_tv3d/y"]
    tau, sigma, rho = (float(g[f"pd3o_tv3d/{k}"]) for k in ("tau", "sigma", "rho"))
    Kop = pxo.Gradient(arg_shape=(10, 12, 14))
    shift = np.ascontiguousarray(-y3.reshape(-1))
    P = E.pds_params(tau, sigma, rho, gspec=POS, fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=0.08)
    d = Kop._desc(1, K.F64)
    x = y3.reshape(-1).copy()
    z = E.gradient_run(Kop, x, False)
    u = x.copy()
    for _ in range(50):
        u, z = one_pass(K.ALGO_PD3O, d, P, u, z, x, form="tma")
    assert relerr(x, g["pd3o_tv3d/x"]) < 1e-10 and relerr(z, g["pd3o_tv3d/z"]) < 1e-10


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tile2d_form_fp64_shift_modes_and_golden(scheme):
    """TMA-tiled 2-D form: per-item / broadcast / scalar shifts, L1 and L21 duals, PD3O and CondatVu; then the reference's fixture."""
    rng = np.random.default_rng(13)
    shape, batch = (37, 70), 2
    Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, sampling=(0.5, 2.0))
    d = Kop._desc(batch, K.F64)
    u, x = rng.standard_normal((batch, Kop.dim)), rng.standard_normal((batch, Kop.dim))
    z = rng.standard_normal((batch, Kop.codim))
    for shift in (rng.standard_normal(Kop.dim), rng.standard_normal((batch, Kop.dim)), np.r_[0.3]):
        for hkind, gspec in ((K.DUAL_L21, (K.PROX_POS, 0.0, 0.0)), (K.DUAL_L1, (K.PROX_BOX, -0.2, 0.9))):
            P = E.pds_params(0.21, 0.19, 0.9, gspec=gspec, fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=hkind, lam=0.3)
            for algo in (K.ALGO_PD3O, K.ALGO_CV):
                ua, za, xa = u.copy(), z.copy(), x.copy()
                nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
                two_pass(algo, d, P, ua, za, xa, nxa, nza)
                xb = x.copy()
                ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, form="tile2d")
                assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13, (scheme, algo, shift.shape)
                if algo == K.ALGO_PD3O:
                    assert relerr(xb, xa) < 1e-13
                assert np.allclose(nxa, nxb, rtol=1e-10) and np.allclose(nza, nzb, rtol=1e-10)
    if scheme == "forward":
        g = golden("solvers.npz")
        y = g["pd3o_tv2d/y"]
        tau, sigma, rho = (float(g[f"pd3o_tv2d/s1/{k}"]) for k in ("tau", "sigma", "rho"))
        Kop = pxo.Gradient(arg_shape=(32, 40))
        shift = np.ascontiguousarray(-y.reshape(-1))  # (kept alive: the parameter block only holds its address)
        P = E.pds_params(tau, sigma, rho, gspec=POS, fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=0.1)
        d = Kop._desc(1, K.F64)
        x = y.reshape(-1).copy()
        z = E.gradient_run(Kop, x, False)
        u = x.copy()
        for _ in range(60):
            u, z = one_pass(K.ALGO_PD3O, d, P, u, z, x, form="tile2d")
        assert relerr(x, g["pd3o_tv2d/s1/x"]) < 1e-10 and relerr(z, g["pd3o_tv2d/s1/z"]) < 1e-10


# ---------------------------------------------------------------------------------------------------------
# Folding boundary modes (numpy.pad wrap / reflect / symmetric / edge, pad.py:252-302) inside the single-kernel forms:
# the MODES instances recompute the two-sample band along each folding face and the out-of-domain rim of the w tiles
# through the per-sample boundary map / pre-image gather.  Reference: the generic two-pass bodies (themselves checked
# against fixtures of the real reference for every mode, test_emu_kernels.py).
# ---------------------------------------------------------------------------------------------------------
MODES = ["wrap", "reflect", "symmetric", "edge"]


@pytest.mark.parametrize("mode", MODES + [("constant", "reflect", "wrap"), ("edge", "constant", "symmetric")])
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_modes_3d(mode, scheme, form):
    rng = np.random.default_rng(21)
    # ragged tiles; full tiles (the folded rim of an edge tile is then served from its own staged boxes); a tiny volume
    for dtype, shape in ((np.float64, (7, 19, 2 * 64 + 6)), (np.float32, (3, 9, 4 * 32 + 8)), (np.float64, (9, 16, 128)), (np.float32, (4, 8, 256)), (np.float64, (3, 3, 4))):
        Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=mode, sampling=(1.0, 0.5, 2.0))
        d = Kop._desc(1, E.dcode(np.zeros(1, dtype=dtype)))
        shift = rng.standard_normal(Kop.dim).astype(dtype)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            P = E.pds_params(0.21, 0.19, 0.9, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L21, lam=0.3)
            for chunk in (0, 3):
                u = rng.standard_normal(Kop.dim).astype(dtype)
                x = rng.standard_normal(Kop.dim).astype(dtype)
                z = rng.standard_normal(Kop.codim).astype(dtype)
                ua, za, xa = u.copy(), z.copy(), x.copy()
                nxa, nza, nxb, nzb = np.zeros(2), np.zeros(2), np.zeros(2), np.zeros(2)
                two_pass(algo, d, P, ua, za, xa, nxa, nza)
                xb = x.copy()
                ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=chunk, form=form)
                tol = 1e-13 if dtype == np.float64 else 2e-6
                assert relerr(ub, ua) < tol and relerr(zb, za) < tol, (mode, scheme, algo, chunk, shape)
                if algo == K.ALGO_PD3O:
                    assert relerr(xb, xa) < tol
                rt = 1e-5 if dtype == np.float32 else 1e-10
                assert np.allclose(nxa, nxb, rtol=rt) and np.allclose(nza, nzb, rtol=rt)


@pytest.mark.parametrize("mode", MODES + [("reflect", "wrap"), ("constant", "edge")])
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("form", ["direct", "tile2d"])
def test_iter_modes_2d_batched(mode, scheme, form):
    rng = np.random.default_rng(23)
    for dtype, shape, batch in ((np.float32, (37, 300), 2), (np.float64, (16, 64), 1), (np.float32, (32, 256), 2), (np.float64, (3, 4), 3)):
        Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=mode)
        d = Kop._desc(batch, E.dcode(np.zeros(1, dtype=dtype)))
        shift = rng.standard_normal((batch, Kop.dim)).astype(dtype)
        garr = rng.standard_normal((batch, Kop.dim)).astype(dtype)
        u = rng.standard_normal((batch, Kop.dim)).astype(dtype)
        x = rng.standard_normal((batch, Kop.dim)).astype(dtype)
        z = rng.standard_normal((batch, Kop.codim)).astype(dtype)
        cases = [(K.ALGO_PD3O, dict(fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21)),
                 (K.ALGO_CV, dict(fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L1)),
                 (K.ALGO_CV, dict(fkind=K.F_GRADARR, garr=garr, hkind=K.DUAL_L21))]
        for algo, kw in cases:
            P = E.pds_params(0.3, 0.25, 0.95, gspec=(K.PROX_L1, 0.05, 0.0), lam=0.2, **kw)
            ua, za, xa = u.copy(), z.copy(), x.copy()
            nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
            two_pass(algo, d, P, ua, za, xa, nxa, nza)
            xb = x.copy()
            ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, form=form)
            tol = 1e-13 if dtype == np.float64 else 2e-6
            assert relerr(ub, ua) < tol and relerr(zb, za) < tol, (mode, scheme, algo, shape)
            if algo == K.ALGO_PD3O:
                assert relerr(xb, xa) < tol
            rt = 1e-5 if dtype == np.float32 else 1e-10
            assert np.allclose(nxa, nxb, rtol=rt) and np.allclose(nza, nzb, rtol=rt)


def test_iter_modes_stacked_2d_in_3d_shape():
    """2-D gradient over the last two axes of a 3-D arg_shape: axis 0 enumerates images, its mode is never consulted."""
    rng = np.random.default_rng(25)
    Kop = pxo.Gradient(arg_shape=(3, 11, 24), directions=(1, 2), mode=("constant", "reflect", "wrap"))
    d = Kop._desc(2, K.F64)
    shift = np.r_[0.2]  # (kept alive: the parameter block only holds its address)
    P = E.pds_params(0.3, 0.25, 0.8, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.4, shift=shift, hkind=K.DUAL_L21, lam=0.2)
    u, x, z = rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.codim))
    for form in ("direct", "tile2d"):
        ua, za, xa = u.copy(), z.copy(), x.copy()
        two_pass(K.ALGO_PD3O, d, P, ua, za, xa)
        xb = x.copy()
        ub, zb = one_pass(K.ALGO_PD3O, d, P, u, z, xb, form=form)
        assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13 and relerr(xb, xa) < 1e-13


def _solve_modes(prefix, shape, mode, n_iter, lam, gspec, y, x0, form):
    g = golden("solvers.npz")
    tau, sigma, rho = (float(g[f"{prefix}/{k}"]) for k in ("tau", "sigma", "rho"))
    Kop = pxo.Gradient(arg_shape=shape, mode=mode)
    shift = np.ascontiguousarray(-y.reshape(-1))
    P = E.pds_params(tau, sigma, rho, gspec=gspec, fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=lam)
    d = Kop._desc(1, K.F64)
    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    z = E.gradient_run(Kop, x, False)
    u = x.copy()
    for _ in range(n_iter):
        u, z = one_pass(K.ALGO_PD3O, d, P, u, z, x, form=form)
    return x, z, g


@pytest.mark.parametrize("mode", ["reflect", "wrap", "symmetric", "edge"])
@pytest.mark.parametrize("form", ["direct", "tile2d"])
def test_iter_modes_golden_2d(mode, form):
    """40 PD3O iterations with a folding boundary mode against the real reference's NumPy float64 solver (<= 1e-10)."""
    y = golden("solvers.npz")["pd3o_tv2d/y"]
    x, z, g = _solve_modes(f"pd3o_tv2d/{mode}", (32, 40), mode, 40, 0.15, (K.PROX_NONE, 0.0, 0.0), y, np.zeros(y.size), form)
    assert relerr(x, g[f"pd3o_tv2d/{mode}/x"]) < 1e-10 and relerr(z, g[f"pd3o_tv2d/{mode}/z"]) < 1e-10


@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_modes_golden_3d_mixed(form):
    y3 = golden("solvers.npz")["pd3o_tv3d/y"]
    x, z, g = _solve_modes("pd3o_tv3d/mixed", (10, 12, 14), ("reflect", "wrap", "constant"), 30, 0.08, POS, y3, y3.reshape(-1), form)
    assert relerr(x, g["pd3o_tv3d/mixed/x"]) < 1e-10 and relerr(z, g["pd3o_tv3d/mixed/z"]) < 1e-10


def test_iter_modes_folded_rim_is_served_from_the_tile():
    """On full tiles the staged forms evaluate a folded rim row / column from the tile's own boxes (pxb_rim_src): the
    global-memory evaluator only runs for the plane past the last one (3-D) and for 'wrap', whose fold lands in another tile."""
    rng = np.random.default_rng(31)
    shape3, shape2 = (8, 16, 256), (32, 256)  # fp32: 2 x 2 tiles per plane; 2 x 2 tiles
    counts = {}
    for mode in ("reflect", "symmetric", "edge", "wrap"):
        for form, shape in (("tma", shape3), ("tile2d", shape2)):
            Kop = pxo.Gradient(arg_shape=shape, mode=mode, dtype=np.float32)
            d = Kop._desc(1, K.F32)
            shift = rng.standard_normal(Kop.dim).astype(np.float32)
            P = E.pds_params(0.21, 0.19, 0.9, gspec=POS, fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L21, lam=0.3)
            u, x, z = (rng.standard_normal(n).astype(np.float32) for n in (Kop.dim, Kop.dim, Kop.codim))
            ua, za, xa = u.copy(), z.copy(), x.copy()
            two_pass(K.ALGO_PD3O, d, P, ua, za, xa)
            E.lib().emu_w_global_cells(1)
            ub, zb = one_pass(K.ALGO_PD3O, d, P, u, z, x.copy(), form=form)
            counts[mode, form] = E.lib().emu_w_global_cells(1)
            assert relerr(ub, ua) < 2e-6 and relerr(zb, za) < 2e-6
    plane = shape3[1] * shape3[2]
    for mode in ("reflect", "symmetric", "edge"):
        assert counts[mode, "tile2d"] == 0
        assert counts[mode, "tma"] == plane  # forward differences: the plane past the last one, once per tile
    assert counts["wrap", "tile2d"] == shape2[0] + shape2[1]          # one rim column cell per row + one rim row
    assert counts["wrap", "tma"] == plane + shape3[0] * (shape3[1] + shape3[2])


@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_modes_3d_batched_and_degenerate_axes(form):
    """Folding modes with a batch of volumes (per-item and broadcast shifts), and axes of length 1 / 2 where both faces of
    an axis fold onto the same or neighbouring samples."""
    rng = np.random.default_rng(41)
    for shape, mode, batch in (((5, 11, 24), ("reflect", "wrap", "symmetric"), 2), ((1, 9, 8), ("symmetric", "edge", "wrap"), 3),
                               ((2, 2, 8), ("wrap", "symmetric", "edge"), 2), ((4, 1, 4), ("edge", "wrap", "reflect"), 1)):
        for scheme in SCHEMES:
            if scheme == "central" and min(shape) == 1:
                continue  # a 3-tap kernel pads by 2 > the axis length: refused at construction, as in the reference (pad.py:217-229)
            Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=mode)
            d = Kop._desc(batch, K.F64)
            u, x = rng.standard_normal((batch, Kop.dim)), rng.standard_normal((batch, Kop.dim))
            z = rng.standard_normal((batch, Kop.codim))
            for shift in (rng.standard_normal(Kop.dim), rng.standard_normal((batch, Kop.dim))):
                P = E.pds_params(0.21, 0.19, 0.9, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L21, lam=0.3)
                for algo in (K.ALGO_PD3O, K.ALGO_CV):
                    ua, za, xa = u.copy(), z.copy(), x.copy()
                    nxa, nza, nxb, nzb = (np.zeros(2 * batch) for _ in range(4))
                    two_pass(algo, d, P, ua, za, xa, nxa, nza)
                    xb = x.copy()
                    ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=2, form=form)
                    assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13, (shape, mode, scheme, algo, shift.shape)
                    assert np.allclose(nxa, nxb, rtol=1e-10) and np.allclose(nza, nzb, rtol=1e-10)


def test_iter_modes_periodic_shift_direct_form():
    """A shift that repeats with a period of its own (one image of a stack: PXB_SHIFT_MOD) is declined by the staged forms and
    served by the direct-load one, folding modes included (the out-of-line evaluator addresses the shift the same way)."""
    rng = np.random.default_rng(43)
    Kop = pxo.Gradient(arg_shape=(3, 11, 24), directions=(1, 2), mode=("constant", "symmetric", "wrap"))
    d = Kop._desc(2, K.F64)
    shift = rng.standard_normal(11 * 24)  # one image, broadcast over the 3 images of each of the 2 batch items
    P = E.pds_params(0.3, 0.25, 0.8, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.4, shift=shift, hkind=K.DUAL_L21, lam=0.2)
    u, x, z = rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.codim))
    a, zz = np.zeros_like(u), np.zeros_like(z)
    assert E.lib().emu_tv_tile2d(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(u), E.p(z), E.p(a), E.p(zz), E.p(x.copy()), None, None, 0) == -121
    for algo in (K.ALGO_PD3O, K.ALGO_CV):
        ua, za, xa = u.copy(), z.copy(), x.copy()
        two_pass(algo, d, P, ua, za, xa)
        xb = x.copy()
        ub, zb = one_pass(algo, d, P, u, z, xb, form="direct")
        assert relerr(ub, ua) < 1e-13 and relerr(zb, za) < 1e-13


def test_iter_modes_random_geometries():
    """Random shapes around the tile sizes (8 rows x 32 vectors in 3-D, 16 x 32 vectors in 2-D), modes, schemes, batches and
    chunk lengths: every single-kernel form against the two-pass generic bodies."""
    rng = np.random.default_rng(2024)
    names = ["constant", "wrap", "reflect", "symmetric", "edge"]
    for trial in range(36):
        ndim = 3 if trial % 2 == 0 else 2
        dtype = np.float64 if trial % 3 else np.float32
        vec = 2 if dtype == np.float64 else 4
        t2 = 32 * vec
        cols = int(rng.choice([vec * int(rng.integers(1, 6)), t2, t2 + vec * int(rng.integers(1, 4)), 2 * t2]))
        if ndim == 3:
            shape = (int(rng.integers(1, 12)), int(rng.choice([3, 7, 8, 9, 16, 17])), cols)
        else:
            shape = (int(rng.choice([3, 15, 16, 17, 32, 33])), cols)
        scheme = str(rng.choice(SCHEMES))
        p = 2 if scheme == "central" else 1
        mode = []
        for n in shape:
            ok = [m for m in names if m in ("constant", "edge") or p <= (n - 1 if m == "reflect" else n)]
            ok = [m for m in ok if not (m == "reflect" and n <= 2)]
            mode.append(str(rng.choice(ok)))
        batch = int(rng.integers(1, 3))
        Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=tuple(mode))
        d = Kop._desc(batch, E.dcode(np.zeros(1, dtype=dtype)))
        shift = rng.standard_normal((batch, Kop.dim)).astype(dtype)
        hk, gs = ((K.DUAL_L21, (K.PROX_POS, 0.0, 0.0)), (K.DUAL_L1, (K.PROX_BOX, -0.3, 0.8)))[trial % 2]
        P = E.pds_params(0.21, 0.19, 0.9, gspec=gs, fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=hk, lam=0.3)
        u, x = (rng.standard_normal((batch, Kop.dim)).astype(dtype) for _ in range(2))
        z = rng.standard_normal((batch, Kop.codim)).astype(dtype)
        algo = K.ALGO_PD3O if trial % 4 < 2 else K.ALGO_CV
        ua, za, xa = u.copy(), z.copy(), x.copy()
        nxa, nza = np.zeros(2 * batch), np.zeros(2 * batch)
        two_pass(algo, d, P, ua, za, xa, nxa, nza)
        tol = 1e-13 if dtype == np.float64 else 3e-6
        for form in (("direct", "tma") if ndim == 3 else ("direct", "tile2d")):
            nxb, nzb = np.zeros(2 * batch), np.zeros(2 * batch)
            xb = x.copy()
            ub, zb = one_pass(algo, d, P, u, z, xb, nxb, nzb, chunk=int(rng.choice([0, 1, 3])) if form != "tile2d" else 0, form=form)
            assert relerr(ub, ua) < tol and relerr(zb, za) < tol, (trial, shape, mode, scheme, form, algo)
            if algo == K.ALGO_PD3O:
                assert relerr(xb, xa) < tol
            rt = 1e-10 if dtype == np.float64 else 1e-4
            assert np.allclose(nxa, nxb, rtol=rt) and np.allclose(nza, nzb, rtol=rt), (trial, shape, mode, scheme, form)
