"""
CPU checks of the CUDA kernels' per-voxel bodies (pyxu_b200/csrc/pxb_core.cuh compiled for the host,
see tests/emu) driven by the descriptors the Python host layer builds, against fixtures produced by
the real reference.  This validates boundary maps, adjoint pre-image gathering, the fused PD3O / CV
half-steps and the host->C-ABI translation without a GPU.  (The GPU tests repeat all of this through
the real library.)
"""
import ctypes as C
import types

import numpy as np
import pytest

import cases
import emu_util as E
import pyxu_b200.operator as pxo
from conftest import golden
from pyxu_b200 import _cabi as K

ns = types.SimpleNamespace(operator=pxo)


def relerr(a, b):
    d = np.linalg.norm(np.asarray(b, dtype=np.float64).ravel())
    return np.linalg.norm((np.asarray(a, dtype=np.float64) - b).ravel()) / (d if d else 1.0)


@pytest.mark.parametrize("case", cases.STENCIL_CASES, ids=lambda c: c["name"])
def test_stencil_bodies(case):
    g, n = golden("stencil.npz"), case["name"]
    op = cases.make_stencil(ns, case)
    assert relerr(E.stencil_run(op, g[f"{n}/x"], False), g[f"{n}/apply"]) < 1e-13
    assert relerr(E.stencil_run(op, g[f"{n}/y"], True), g[f"{n}/adjoint"]) < 1e-13
    assert abs(op.lipschitz - float(g[f"{n}/lipschitz"])) <= 1e-12 * op.lipschitz


@pytest.mark.parametrize("case", cases.STENCIL_CASES[:8], ids=lambda c: c["name"])
def test_stencil_bodies_f32(case):
    g, n = golden("stencil.npz"), case["name"]
    op = cases.make_stencil(ns, case, dtype=np.float32)
    assert relerr(E.stencil_run(op, g[f"{n}/x"].astype(np.float32), False), g[f"{n}/apply"]) < 2e-6
    assert relerr(E.stencil_run(op, g[f"{n}/y"].astype(np.float32), True), g[f"{n}/adjoint"]) < 2e-6


@pytest.mark.parametrize("case", cases.GRADIENT_CASES, ids=lambda c: c["name"])
def test_gradient_bodies(case):
    g, n = golden("gradient.npz"), case["name"]
    op = cases.make_gradient(ns, case)
    assert relerr(E.gradient_run(op, g[f"{n}/x"], False), g[f"{n}/apply"]) < 1e-13
    assert relerr(E.gradient_run(op, g[f"{n}/y"], True), g[f"{n}/adjoint"]) < 1e-13
    assert abs(op.lipschitz - float(g[f"{n}/lipschitz"])) <= 1e-12 * op.lipschitz


def test_adjoint_is_transpose_random_modes():
    """<Ax, y> == <x, A^T y> for every mode combination, kernel bigger than needed, tiny arrays."""
    rng = np.random.default_rng(5)
    for trial in range(60):
        D = rng.integers(1, 4)
        shape = tuple(int(rng.integers(3, 7)) for _ in range(D))
        modes = tuple(rng.choice(list(K.MODES)) for _ in range(D))
        ks = []
        for n, m in zip(shape, modes):
            lim = {"constant": 5, "edge": 5, "wrap": n + 1, "symmetric": n + 1, "reflect": n}[m]
            ks.append(int(rng.integers(1, min(5, lim) + 1)))
        kern = rng.standard_normal(ks)
        cen = tuple(int(rng.integers(0, k)) for k in ks)
        op = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode=modes)
        x, y = rng.standard_normal(op.dim), rng.standard_normal(op.dim)
        lhs = np.dot(E.stencil_run(op, x, False), y)
        rhs = np.dot(x, E.stencil_run(op, y, True))
        assert abs(lhs - rhs) < 1e-10 * (1 + abs(lhs)), (shape, modes, ks, cen)


def _fused_run(name, algo, shape, n_iter, lam, gspec, y, x0, mode="constant", dtype=np.float64, vec=0):
    """vec = 0: generic per-voxel bodies; vec in {1, 2, 4}: vectorised fast bodies (pxb_tv_fast.cuh)."""
    g = golden("solvers.npz")
    tau, sigma, rho = (float(g[f"{name}/{k}"]) for k in ("tau", "sigma", "rho"))
    Kop = pxo.Gradient(arg_shape=shape, mode=mode)
    D, N = len(shape), int(np.prod(shape))
    shift = np.ascontiguousarray(-y.reshape(-1), dtype=dtype)
    P = E.pds_params(tau, sigma, rho, gspec=gspec, fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=lam)
    d = Kop._desc(1, K.F32 if dtype == np.float32 else K.F64)
    x = np.ascontiguousarray(x0, dtype=dtype).copy()
    z = E.gradient_run(Kop, x, False)
    u, w = x.copy(), np.empty_like(x)
    nx, nz = np.zeros(2), np.zeros(2)
    for _ in range(n_iter):
        nx[:] = 0
        nz[:] = 0
        xu, xo = (u, x) if algo == K.ALGO_PD3O else (x, None)
        if vec:
            assert E.lib().emu_tv_fast(vec, 0, algo, C.byref(d), C.byref(P), E.p(xu), E.p(z), E.p(xo), E.p(w), E.p(nx)) == 0
            assert E.lib().emu_tv_fast(vec, 1, algo, C.byref(d), C.byref(P), None, E.p(z), None, E.p(w), E.p(nz)) == 0
        else:
            E.lib().emu_pds_primal(algo, C.byref(d), C.byref(P), E.p(xu), E.p(z), None, E.p(xo), E.p(w), E.p(nx))
            E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(z), E.p(nz))
    return x, z, nx, nz, g


POS = (K.PROX_POS, 0.0, 0.0)
NONE = (K.PROX_NONE, 0.0, 0.0)


@pytest.mark.parametrize("strat", [1, 2, 3])
def test_fused_pd3o_tv2d(strat):
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _fused_run(f"pd3o_tv2d/s{strat}", K.ALGO_PD3O, (32, 40), 60, 0.1, POS, y, y.reshape(-1))
    assert relerr(x, g[f"pd3o_tv2d/s{strat}/x"]) < 1e-10
    assert relerr(z, g[f"pd3o_tv2d/s{strat}/z"]) < 1e-10


@pytest.mark.parametrize("mode", ["reflect", "wrap", "symmetric", "edge"])
def test_fused_pd3o_tv2d_modes(mode):
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _fused_run(f"pd3o_tv2d/{mode}", K.ALGO_PD3O, (32, 40), 40, 0.15, NONE, y, np.zeros(y.size), mode=mode)
    assert relerr(x, g[f"pd3o_tv2d/{mode}/x"]) < 1e-10
    assert relerr(z, g[f"pd3o_tv2d/{mode}/z"]) < 1e-10


def test_fused_cv_tv2d():
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _fused_run("cv_tv2d", K.ALGO_CV, (32, 40), 60, 0.1, POS, y, y.reshape(-1))
    assert relerr(x, g["cv_tv2d/x"]) < 1e-10
    assert relerr(z, g["cv_tv2d/z"]) < 1e-10


def test_fused_pd3o_tv3d_and_norms():
    g = golden("solvers.npz")
    y = g["pd3o_tv3d/y"]
    x, z, nx, nz, _ = _fused_run("pd3o_tv3d", K.ALGO_PD3O, (10, 12, 14), 50, 0.08, POS, y, y.reshape(-1))
    assert relerr(x, g["pd3o_tv3d/x"]) < 1e-10
    assert relerr(z, g["pd3o_tv3d/z"]) < 1e-10
    x2, z2, *_ = _fused_run("pd3o_tv3d/mixed", K.ALGO_PD3O, (10, 12, 14), 30, 0.08, POS, y, y.reshape(-1),
                            mode=("reflect", "wrap", "constant"))
    assert relerr(x2, g["pd3o_tv3d/mixed/x"]) < 1e-10
    # fused RelError norms of the last iteration == norms recomputed from iterates 49 -> 50
    xa, za, *_ = _fused_run("pd3o_tv3d", K.ALGO_PD3O, (10, 12, 14), 49, 0.08, POS, y, y.reshape(-1))
    assert abs(nx[0] - np.sum((x - xa) ** 2)) < 1e-12 * (1 + nx[0]) and abs(nx[1] - np.sum(xa**2)) < 1e-9 * nx[1]
    assert abs(nz[0] - np.sum((z - za) ** 2)) < 1e-12 * (1 + nz[0]) and abs(nz[1] - np.sum(za**2)) < 1e-9 * nz[1]


def test_fused_pd3o_f32_tolerance():
    """fp32 kernels vs the reference's float64 solver: rel. L2 error <= 1e-4 (north-star tolerance)."""
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    x, z, *_ = _fused_run("pd3o_tv2d/s1", K.ALGO_PD3O, (32, 40), 60, 0.1, POS, y, y.reshape(-1), dtype=np.float32)
    assert relerr(x, g["pd3o_tv2d/s1/x"]) < 1e-4


def test_funcs_bodies():
    g = golden("funcs.npz")
    x = g["x"]
    lib = E.lib()

    def prox(spec, tau):
        out = np.empty_like(x)
        s = K.ProxSpec(spec[0], 0, spec[1], spec[2])
        lib.emu_prox_lincomb(K.F64, C.byref(s), tau, x.size, E.p(out), 1.0, E.p(x), 0.0, None, 0, 0.0, None, 0)
        return out

    for tau in (0.3, 1.7):
        t = f"{tau}"
        assert relerr(prox((K.PROX_L1, 1.0, 0), tau), g[f"l1/prox/{t}"]) < 1e-14
        assert relerr(prox((K.PROX_L1, 0.4, 0), tau), g[f"l1s/prox/{t}"]) < 1e-14
        assert relerr(prox((K.PROX_POSL1, 1.0, 0), tau), g[f"posl1/prox/{t}"]) < 1e-14
        assert relerr(prox((K.PROX_POS, 0, 0), tau), g[f"pos/prox/{t}"]) < 1e-14
        assert relerr(prox((K.PROX_BOX, -0.8, 0.8), tau), g[f"linfball/prox/{t}"]) < 1e-14
        assert relerr(prox((K.PROX_SQL2, 1.0, 0), tau), g[f"sql2/prox/{t}"]) < 1e-14
        out = np.empty_like(x)
        lib.emu_prox_l21(K.F64, 3, 3, 20, 1.0, tau, E.p(x), E.p(out))
        assert relerr(out, g[f"l21/prox/{t}"]) < 1e-14
        lib.emu_prox_l21(K.F64, 9, 20, 1, 1.0, tau, E.p(x), E.p(out))
        assert relerr(out, g[f"l21ax12/prox/{t}"]) < 1e-14
        # fenchel prox of 0.7*L21 at x == dual_update with z=0... use z = x, t = 0, rho = 1: prox_{sigma h*}(x)
        z, tt = x.copy(), np.zeros_like(x)
        lib.emu_dual_update(K.F64, K.DUAL_L21, 3, 3, 20, 0.7, tau, 1.0, E.p(z), E.p(tt), None)
        assert relerr(z, g[f"l21s/fprox/{t}"]) < 1e-14
        z = x.copy()
        lib.emu_dual_update(K.F64, K.DUAL_L1, 3, 1, 60, 1.0, tau, 1.0, E.p(z), E.p(tt), None)
        assert relerr(z, g[f"l1/fprox/{t}"]) < 1e-14


# ---- vectorised fast bodies (what the B200 runs for first-order-FD TV problems) ------------------------
@pytest.mark.parametrize("vec", [1, 2, 4])
@pytest.mark.parametrize("mode", ["constant", "reflect", "wrap", "symmetric", "edge"])
def test_fast_bodies_pd3o_2d_all_modes(vec, mode):
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    if mode == "constant":
        name, lam, gspec, x0, it = "pd3o_tv2d/s3", 0.1, POS, y.reshape(-1), 60
    else:
        name, lam, gspec, x0, it = f"pd3o_tv2d/{mode}", 0.15, NONE, np.zeros(y.size), 40
    x, z, nx, nz, _ = _fused_run(name, K.ALGO_PD3O, (32, 40), it, lam, gspec, y, x0, mode=mode, vec=vec)
    assert relerr(x, g[f"{name}/x"]) < 1e-10 and relerr(z, g[f"{name}/z"]) < 1e-10
    # identical to the generic bodies up to rounding, including the fused norms
    x2, z2, nx2, nz2, _ = _fused_run(name, K.ALGO_PD3O, (32, 40), it, lam, gspec, y, x0, mode=mode, vec=0)
    assert relerr(x, x2) < 1e-13 and relerr(z, z2) < 1e-13
    assert np.allclose(nx, nx2, rtol=1e-9, atol=1e-30) and np.allclose(nz, nz2, rtol=1e-9, atol=1e-30)


@pytest.mark.parametrize("vec", [1, 2])
def test_fast_bodies_3d_and_cv(vec):
    g = golden("solvers.npz")
    y = g["pd3o_tv3d/y"]
    x, z, *_ = _fused_run("pd3o_tv3d", K.ALGO_PD3O, (10, 12, 14), 50, 0.08, POS, y, y.reshape(-1), vec=vec)
    assert relerr(x, g["pd3o_tv3d/x"]) < 1e-10 and relerr(z, g["pd3o_tv3d/z"]) < 1e-10
    x, z, *_ = _fused_run("pd3o_tv3d/mixed", K.ALGO_PD3O, (10, 12, 14), 30, 0.08, POS, y, y.reshape(-1),
                          mode=("reflect", "wrap", "constant"), vec=vec)
    assert relerr(x, g["pd3o_tv3d/mixed/x"]) < 1e-10 and relerr(z, g["pd3o_tv3d/mixed/z"]) < 1e-10
    y2 = g["pd3o_tv2d/y"]
    x, z, *_ = _fused_run("cv_tv2d", K.ALGO_CV, (32, 40), 60, 0.1, POS, y2, y2.reshape(-1), vec=2 * vec)
    assert relerr(x, g["cv_tv2d/x"]) < 1e-10 and relerr(z, g["cv_tv2d/z"]) < 1e-10


@pytest.mark.parametrize("scheme", ["backward", "central"])
def test_fast_bodies_other_schemes_equal_generic(scheme):
    """backward / central differences (taps at -1 / +-1) under every mode: fast == generic bodies."""
    rng = np.random.default_rng(2)
    shape = (6, 8, 12)
    for mode in ("constant", ("reflect", "symmetric", "wrap"), "edge"):
        Kop = pxo.Gradient(arg_shape=shape, mode=mode, scheme=scheme, sampling=(1.0, 0.5, 2.0))
        d = Kop._desc(1, K.F64)
        shift = rng.standard_normal(Kop.dim)
        P = E.pds_params(0.21, 0.19, 1.2, gspec=(K.PROX_L1, 0.05, 0.0), fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L1, lam=0.3)
        outs = []
        for vec in (0, 1, 2, 4):
            r = np.random.default_rng(3)
            u, x, w = r.standard_normal(Kop.dim), r.standard_normal(Kop.dim), np.zeros(Kop.dim)
            z = r.standard_normal(Kop.codim)
            for _ in range(3):
                if vec:
                    assert E.lib().emu_tv_fast(vec, 0, K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(u), E.p(z), E.p(x), E.p(w), None) == 0
                    assert E.lib().emu_tv_fast(vec, 1, 0, C.byref(d), C.byref(P), None, E.p(z), None, E.p(w), None) == 0
                else:
                    E.lib().emu_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(u), E.p(z), None, E.p(x), E.p(w), None)
                    E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(z), None)
            outs.append((u.copy(), x.copy(), z.copy()))
        for o in outs[1:]:
            for a, b in zip(o, outs[0]):
                assert relerr(a, b) < 1e-13, (scheme, mode)


@pytest.mark.parametrize("scheme", ["forward", "backward", "central"])
@pytest.mark.parametrize("vec", [1, 2, 4])
def test_fast_gradient_bodies_equal_generic(scheme, vec):
    """Vectorised Gradient apply / adjoint bodies (what pxb_gradient_apply / adjoint launch for first-order stacks) == the generic
    per-voxel bodies (pinned on the reference's fixtures above), every mode, 1-D / 2-D / 3-D, stacked inputs."""
    rng = np.random.default_rng(5)
    for shape, mode in (((6, 8, 12), "constant"), ((6, 8, 12), ("reflect", "symmetric", "wrap")), ((5, 8, 12), "edge"), ((9, 16), ("wrap", "reflect")), ((24,), "symmetric")):
        Kop = pxo.Gradient(arg_shape=shape, mode=mode, scheme=scheme, sampling=0.5)
        d = Kop._desc(2, K.F64)
        x, z = rng.standard_normal((2, Kop.dim)), rng.standard_normal((2, Kop.codim))
        for adj, src, n_out in ((0, x, Kop.codim), (1, z, Kop.dim)):
            a, b = np.full((2, n_out), np.nan), np.full((2, n_out), np.nan)
            E.lib().emu_gradient(C.byref(d), adj, E.p(src), E.p(a))
            assert E.lib().emu_tv_grad(vec, adj, C.byref(d), E.p(src), E.p(b)) == 0
            assert relerr(b, a) < 1e-13, (scheme, vec, shape, mode, adj)


def test_fast_bodies_fold_terms_random_small_geometries():
    """The vectorised bodies have no per-voxel fallback: folding boundary modes are the 'constant' arithmetic plus the fold
    terms (pxb_tv_fold_kz / pxb_tv_nbr).  Random small geometries -- including lines as short as the vector, where the two
    faces of an axis fold onto the same or neighbouring samples -- against the generic bodies."""
    rng = np.random.default_rng(77)
    modes_all = ["constant", "wrap", "reflect", "symmetric", "edge"]
    for trial in range(60):
        D = int(rng.integers(1, 4))
        vec = int(rng.choice([1, 2, 4]))
        shape = tuple(int(rng.integers(3, 7)) for _ in range(D - 1)) + (vec * int(rng.integers(1, 4)) if vec > 2 else max(3, vec * int(rng.integers(1, 4))),)
        if shape[-1] % vec:
            vec = 1
        mode = tuple(str(rng.choice(modes_all)) for _ in range(D))
        scheme = str(rng.choice(["forward", "backward", "central"]))
        Kop = pxo.Gradient(arg_shape=shape, mode=mode, scheme=scheme)
        batch = int(rng.integers(1, 3))
        d = Kop._desc(batch, K.F64)
        x, z = rng.standard_normal((batch, Kop.dim)), rng.standard_normal((batch, Kop.codim))
        for adj, src, n_out in ((0, x, Kop.codim), (1, z, Kop.dim)):
            a, b = np.full((batch, n_out), np.nan), np.full((batch, n_out), np.nan)
            E.lib().emu_gradient(C.byref(d), adj, E.p(src), E.p(a))
            assert E.lib().emu_tv_grad(vec, adj, C.byref(d), E.p(src), E.p(b)) == 0
            assert relerr(b, a) < 1e-13, (trial, shape, mode, scheme, vec, adj)
        shift = rng.standard_normal((batch, Kop.dim))
        P = E.pds_params(0.21, 0.19, 1.1, gspec=(K.PROX_POS, 0.0, 0.0), fkind=K.F_SQL2, alpha=0.7, shift=shift, hkind=K.DUAL_L21, lam=0.3)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            res = []
            for v in (0, vec):
                u, xx, w, zz = x.copy(), x[::-1].copy(), np.zeros_like(x), z.copy()
                if v:
                    assert E.lib().emu_tv_fast(v, 0, algo, C.byref(d), C.byref(P), E.p(u), E.p(zz), E.p(xx) if algo == K.ALGO_PD3O else None, E.p(w), None) == 0
                    assert E.lib().emu_tv_fast(v, 1, algo, C.byref(d), C.byref(P), None, E.p(zz), None, E.p(w), None) == 0
                else:
                    E.lib().emu_pds_primal(algo, C.byref(d), C.byref(P), E.p(u), E.p(zz), None, E.p(xx) if algo == K.ALGO_PD3O else None, E.p(w), None)
                    E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(zz), None)
                res.append((u, w, zz))
            for a_, b_ in zip(res[1], res[0]):
                assert relerr(a_, b_) < 1e-13, (trial, shape, mode, scheme, vec, algo)
