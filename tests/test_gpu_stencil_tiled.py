"""GPU parity of the TMA-tiled 2-D stencil (pxb_stencil2d_apply) through Stencil.apply / adjoint: against the generic
gather kernels on the same inputs (which the golden-vector tests pin on the real reference), larger and ragged sizes,
adjoint identity, epilogue."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return k / k.sum()


CASES = [
    ((517, 1028), [gauss(9, 1.7), gauss(9, 1.7)], (4, 4)),
    ((517, 1028), np.outer(gauss(9, 1.7), gauss(9, 1.7)), (4, 4)),
    ((333, 260), np.arange(1.0, 26.0).reshape(5, 5) / 10, (1, 3)),
    ((5, 130, 264), [np.r_[1.0, 2.0, -1.0], gauss(7, 1.2), gauss(7, 1.2)], (1, 3, 3)),
    ((4, 130, 264), np.arange(1.0, 10.0).reshape(1, 3, 3), (0, 1, 1)),
    ((4100,), np.r_[1.0, 2, -3, 0.5, 7], (2,)),
    ((45, 48), [np.r_[2.0], gauss(11, 2.0)], (0, 4)),
]


def rel(a, b):
    return float(torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()))


def adjoint_gap(op, x, z):
    """|<S x, z> - <x, S^T z>| relative to ||S x|| ||z|| + ||x|| ||S^T z|| (the scale rounding errors of the two sides live on: the
    inner products themselves cancel to almost nothing for random vectors and a kernel with coefficients of both signs)."""
    sx, stz = op.apply(x).double(), op.adjoint(z).double()
    lhs, rhs = torch.sum(sx * z.double()), torch.sum(x.double() * stz)
    scale = torch.linalg.vector_norm(sx) * torch.linalg.vector_norm(z.double()) + torch.linalg.vector_norm(x.double()) * torch.linalg.vector_norm(stz)
    return float(abs(lhs - rhs) / scale)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_tiled_vs_generic(ci, dtype):
    import pyxu_b200.operator as pxo

    shape, kern, cen = CASES[ci]
    k = [np.asarray(_, dtype=dtype) for _ in kern] if isinstance(kern, list) else np.asarray(kern, dtype=dtype)
    tiled = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode="constant")
    generic = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode="constant")
    generic._tiled_ok = False
    gen = torch.Generator(device="cuda").manual_seed(ci)
    x = torch.randn(3, tiled.dim, device="cuda", dtype=torch.float64 if dtype == np.float64 else torch.float32, generator=gen)
    tol = 1e-13 if dtype == np.float64 else 3e-6
    for adj in (False, True):
        a = tiled.adjoint(x) if adj else tiled.apply(x)
        b = generic.adjoint(x) if adj else generic.apply(x)
        assert tiled._tiled_ok is True, "the tiled kernel did not run"
        assert rel(a, b) < tol, (ci, adj, rel(a, b))


def test_tiled_adjoint_identity_and_epilogue():
    import pyxu_b200.operator as pxo

    op = pxo.Stencil(arg_shape=(1000, 1024), kernel=[gauss(9, 1.7), gauss(5, 1.0)], center=(4, 1), mode="constant")
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(op.dim, device="cuda", dtype=torch.float64, generator=gen)
    y = torch.randn(op.dim, device="cuda", dtype=torch.float64, generator=gen)
    lhs, rhs = torch.dot(op.apply(x), y), torch.dot(x, op.adjoint(y))
    assert abs(float(lhs - rhs)) < 1e-10 * (1 + abs(float(lhs)))
    xs = torch.randn(3, op.dim, device="cuda", dtype=torch.float64, generator=gen)
    out = op._run_tiled(xs, False, alpha=0.5, beta=-1.0, add=y)
    assert rel(out, 0.5 * op.apply(xs) - y) < 1e-13


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_pgd_two_pass_fista_equals_generic_path(dtype, tol):
    """PGD on 1/2||A x - y||^2 + lam||x||_1 (batch of images): the two tiled passes per iteration (pxb_stencil2d_fista) against the
    generic five-pass iteration, same iterates and same iteration count under the default RelError criterion."""
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    shape, batch = (96, 128), 4
    N = shape[0] * shape[1]
    rng = np.random.default_rng(0)
    k = (np.outer(gauss(5, 1.0), gauss(5, 1.2)) + 0.02 * np.eye(5)).astype(dtype)  # full rank: the dense instances
    y = rng.random((batch, N)).astype(dtype)

    def build():
        A_ = pxo.Stencil(arg_shape=shape, kernel=k, center=(2, 2), mode="constant")
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)) * A_
        return pxs.PGD(f=f, g=0.02 * pxo.L1Norm(dim=N), show_progress=False, final_writeback=False), A_

    outs = []
    for fused in (True, False):
        slv, A_ = build()
        tau = 1.0 / float(A_.lipschitz) ** 2
        if not fused:
            A_._tiled_ok = False  # the whole operator on the generic gather kernels: no tiled descriptor, no fused step
        slv.fit(x0=np.zeros((batch, N), dtype=dtype), tau=tau, stop_crit=pxst.MaxIter(40) | pxst.RelError(eps=1e-3, var="x"))
        assert (slv._fused is not None) == fused
        assert slv._astate.get("error") is None, slv._astate.get("error")
        data, hist = slv.stats()
        outs.append((np.asarray(data["x"], dtype=np.float64), hist))
    (xa, ha), (xb, hb) = outs
    assert len(ha) == len(hb) and len(ha) > 5
    assert np.linalg.norm(xa - xb) / np.linalg.norm(xb) < tol
    assert np.allclose(ha["RelError[x]_max"] if "RelError[x]_max" in ha.dtype.names else ha["RelError[x]"],
                       hb["RelError[x]_max"] if "RelError[x]_max" in hb.dtype.names else hb["RelError[x]"], rtol=1e-3 if dtype == np.float32 else 1e-8)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 5e-6)])
def test_single_pass_3d_and_fast_gradient(dtype, tol):
    """pxb_stencil3d_apply (separable 3-D stencil in one marching pass) and the vectorised Gradient kernels against the generic ones."""
    import pyxu_b200.operator as pxo

    shape = (70, 45, 264)
    rng = np.random.default_rng(1)
    kern = [gauss(7, 1.2).astype(dtype), rng.standard_normal(5).astype(dtype), gauss(9, 1.5).astype(dtype)]
    fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=(3, 1, 4), mode="constant")
    slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=(3, 1, 4), mode="constant")
    slow._tiled_ok = slow._tiled3d_ok = False
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
    for adj in (False, True):
        a = fast.adjoint(x) if adj else fast.apply(x)
        b = slow.adjoint(x) if adj else slow.apply(x)
        assert fast._tiled3d_ok is True, "the single-pass 3-D kernel did not run"
        assert rel(a, b) < tol, (adj, rel(a, b))
    y = torch.randn(fast.dim, device="cuda", dtype=tdt)
    assert rel(fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y), 0.5 * slow.apply(x) - y) < tol
    # Gradient: first-order stacks take the vectorised kernels; a second-order-accurate stack keeps the generic ones
    for mode in ("constant", ("reflect", "wrap", "edge")):
        G = pxo.Gradient(arg_shape=shape, mode=mode, dtype=dtype, sampling=(1.0, 0.5, 2.0))
        parts = [pxo.PartialDerivative.finite_difference(arg_shape=shape, order=tuple(1 if a == k else 0 for a in range(3)), mode=mode,
                                                         sampling=(1.0, 0.5, 2.0), dtype=dtype) for k in range(3)]
        xs = x[0]
        zg = G.apply(xs)
        ref = torch.cat([p.apply(xs) for p in parts])
        assert rel(zg, ref) < tol
        z = torch.randn(G.codim, device="cuda", dtype=tdt)
        lhs, rhs = torch.dot(zg.double(), z.double()), torch.dot(xs.double(), G.adjoint(z).double())
        assert abs(float(lhs - rhs)) < (1e-10 if dtype == np.float64 else 2e-4) * (1 + abs(float(lhs)))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 5e-6)])
def test_dense_3d_psf_that_is_an_outer_product_takes_the_single_pass(dtype, tol):
    """BASELINE configs[4] names a "7x7x7 Stencil PSF": handed over as a dense array, a Gaussian PSF is recognised as
    a (x) b (x) c and served by pxb_stencil3d_apply (21 taps per sample); a PSF that is not an outer product keeps the gather kernel."""
    import pyxu_b200.operator as pxo

    shape = (40, 45, 264)
    g = [gauss(7, 1.2), gauss(5, 1.0), gauss(7, 1.5)]
    dense = np.einsum("i,j,k->ijk", *g).astype(dtype)
    fast = pxo.Stencil(arg_shape=shape, kernel=dense, center=(3, 2, 3), mode="constant")
    sep = pxo.Stencil(arg_shape=shape, kernel=[k.astype(dtype) for k in g], center=(3, 2, 3), mode="constant")
    slow = pxo.Stencil(arg_shape=shape, kernel=dense, center=(3, 2, 3), mode="constant")
    slow._tiled_ok = slow._tiled3d_ok = False
    x = torch.randn(2, fast.dim, device="cuda", dtype=torch.float64 if dtype == np.float64 else torch.float32)
    for adj in (False, True):
        a, b, c = ((o.adjoint(x) if adj else o.apply(x)) for o in (fast, sep, slow))
        assert fast._tiled3d_ok is True, "the single-pass 3-D kernel did not run"
        assert rel(a, b) < tol and rel(a, c) < tol, (adj, rel(a, b), rel(a, c))
    bumped = dense.copy()
    bumped[0, 0, 0] += 0.01
    assert pxo.Stencil(arg_shape=shape, kernel=bumped, center=(3, 2, 3), mode="constant")._tiled_plan(False) is None


def test_tiled_stencils_random_geometries():
    """Random kernel extents / centres / shapes (degenerate ones included) for every tiled path -- 2-D separable and dense,
    3-D single pass, axis-0 streaming + tiled -- apply and adjoint against the gather kernels."""
    import pyxu_b200.operator as pxo

    rng = np.random.default_rng(7)
    for trial in range(40):
        dtype = np.float64 if trial % 2 else np.float32
        vec = 2 if dtype == np.float64 else 4
        D = 3 if trial % 3 == 0 else 2
        shape = tuple(int(rng.integers(1, 40)) for _ in range(D - 1)) + (vec * int(rng.integers(1, 60)),)
        kmax = 9 if dtype == np.float32 else 7
        ks = [int(rng.integers(1, min(kmax, n) + 1)) if a < D - 1 else int(rng.integers(1, kmax + 1)) for a, n in enumerate(shape)]
        cen = tuple(int(rng.integers(0, k)) for k in ks)
        if trial % 4 == 1 and D == 2:
            kern = rng.standard_normal(ks).astype(dtype)  # dense
        else:
            kern = [rng.standard_normal(k).astype(dtype) for k in ks]
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._tiled_ok = slow._tiled3d_ok = False
        x = torch.randn(int(rng.integers(1, 3)), fast.dim, device="cuda", dtype=torch.float64 if dtype == np.float64 else torch.float32)
        tol = 1e-12 if dtype == np.float64 else 1e-5
        for adj in (False, True):
            a = fast.adjoint(x) if adj else fast.apply(x)
            b = slow.adjoint(x) if adj else slow.apply(x)
            den = float(torch.linalg.vector_norm(b.double()))
            err = float(torch.linalg.vector_norm(a.double() - b.double())) / max(den, 1e-30)
            assert err < tol, (trial, shape, ks, cen, adj, err, fast._tiled_ok, fast._tiled3d_ok)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 5e-6)])
@pytest.mark.parametrize("K", [3, 5, 7, 9])
def test_single_pass_3d_cubic_instances(dtype, tol, K):
    """K x K x K PSFs centred along the rows take the fully unrolled instances (k_stencil3d_fast); against the gather kernels and
    against the general marching kernel (pxb_set_stencil3d_path(1)): apply, adjoint, epilogue operand, ragged tiles, stacks,
    off-centre taps along axes 0 / 1, chunked volumes."""
    import pyxu_b200.operator as pxo
    from pyxu_b200 import _cabi as Kc

    rng = np.random.default_rng(K)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    lib = Kc.lib()
    for shape, cen in (((70, 45, 264), (K // 2, K // 2, K // 2)), ((150, 16, 128), (K - 1, 0, K // 2)), ((K, 37, 8), (0, K - 1, K // 2))):
        kern = [rng.standard_normal(K).astype(dtype) for _ in range(3)]
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._tiled_ok = slow._tiled3d_ok = False
        x = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        y = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        for adj in (False, True):
            n0 = lib.pxb_launch_count()
            a = fast.adjoint(x) if adj else fast.apply(x)
            assert fast._tiled3d_ok is True and lib.pxb_launch_count() - n0 == 1, "one marching pass"
            b = slow.adjoint(x) if adj else slow.apply(x)
            assert rel(a, b) < tol, (shape, cen, adj, rel(a, b))
            Kc.check(lib.pxb_set_stencil3d_path(1), "pxb_set_stencil3d_path")
            try:
                c = fast.adjoint(x) if adj else fast.apply(x)
            finally:
                lib.pxb_set_stencil3d_path(0)
            assert rel(a, c) < tol, (shape, cen, adj, rel(a, c))
        assert rel(fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y), 0.5 * slow.apply(x) - y) < tol


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 1e-5)])
def test_dense_3d_kernel_of_full_rank(dtype, tol, monkeypatch):
    """A dense 3-D PSF that is not an outer product ('constant' boundaries) as one tiled dense 2-D pass per plane of the kernel,
    accumulated in place through the epilogue operand (Stencil._run_dense3d: what serves the kernels the marching kernel declines,
    selected here with PYXU_B200_DENSE3D_MARCH = 0), against the per-sample gather kernel."""
    import pyxu_b200.operator as pxo
    from pyxu_b200 import _cabi as Kc
    from pyxu_b200.operator.linop import stencil as st_mod

    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", False)

    rng = np.random.default_rng(5)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    lib = Kc.lib()
    for shape, ks, cen in (((40, 45, 264), (7, 7, 7), (3, 3, 3)), ((9, 37, 64), (5, 3, 4), (0, 2, 3)), ((4, 50, 8), (7, 2, 3), (6, 0, 1))):
        kern = rng.standard_normal(ks).astype(dtype)
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._dense3d_ok = False
        x = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        y = torch.randn(fast.dim, device="cuda", dtype=tdt)
        for adj in (False, True):
            a = fast.adjoint(x) if adj else fast.apply(x)
            assert fast._dense3d_ok is True
            b = slow.adjoint(x) if adj else slow.apply(x)
            assert slow._dense3d_ok is False and rel(a, b) < tol, (shape, ks, adj, rel(a, b))
        assert rel(fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y), 0.5 * slow.apply(x) - y) < tol
        z = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        assert adjoint_gap(fast, x, z) < (1e-12 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-5)])
def test_dense_3d_marching_kernel(dtype, tol, monkeypatch):
    """A dense 3-D PSF of full rank in ONE marching pass (pxb_stencil3d_dense_apply, selected by PYXU_B200_DENSE3D_MARCH) against the
    per-sample gather kernels (which the golden-vector tests pin on the real reference) and the per-plane tiled passes: cubes and
    embedded kernels, off-centre entries, ragged tiles, several chunks and waves of thread blocks, stacks, epilogue, adjoint identity."""
    import pyxu_b200.operator as pxo
    from pyxu_b200.operator.linop import stencil as st_mod

    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", True)
    rng = np.random.default_rng(6)
    torch.manual_seed(6)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    cases = (((40, 45, 264), (7, 7, 7), (3, 3, 3)), ((70, 130, 520), (7, 7, 7), (0, 6, 1)), ((33, 37, 64), (5, 5, 5), (2, 2, 2)), ((19, 50, 136), (5, 4, 5), (4, 0, 3)),
             ((64, 64, 128), (3, 3, 3), (1, 1, 1)), ((3, 9, 8), (7, 6, 5), (6, 2, 0)))
    for shape, ks, cen in cases:
        kern = rng.standard_normal(ks).astype(dtype)
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._dense3d_ok = False
        x = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        y = torch.randn(fast.dim, device="cuda", dtype=tdt)
        for adj in (False, True):
            a = fast.adjoint(x) if adj else fast.apply(x)
            assert fast._march3d_ok is True
            b = slow.adjoint(x) if adj else slow.apply(x)
            assert slow._dense3d_ok is False and rel(a, b) < tol, (shape, ks, adj, rel(a, b))
        assert rel(fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y), 0.5 * slow.apply(x) - y) < tol
        z = torch.randn(2, fast.dim, device="cuda", dtype=tdt)
        assert adjoint_gap(fast, x, z) < (1e-12 if dtype == np.float64 else 1e-5)
    # against the per-plane tiled passes at a size with many tiles, chunks and waves
    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", False)
    shape = (96, 256, 512)
    kern = rng.standard_normal((7, 7, 7)).astype(dtype)
    planes = pxo.Stencil(arg_shape=shape, kernel=kern, center=(3, 3, 3), mode="constant")
    x = torch.randn(planes.dim, device="cuda", dtype=tdt)
    ref_f, ref_a = planes.apply(x), planes.adjoint(x)
    assert planes._dense3d_ok is True and planes._march3d_ok is None
    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", True)
    march = pxo.Stencil(arg_shape=shape, kernel=kern, center=(3, 3, 3), mode="constant")
    got_f, got_a = march.apply(x), march.adjoint(x)
    assert march._march3d_ok is True and rel(got_f, ref_f) < tol and rel(got_a, ref_a) < tol
    assert torch.equal(march.apply(x), got_f)  # run-to-run identical
