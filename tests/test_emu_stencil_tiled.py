"""
CPU checks of the TMA-tiled 2-D stencil (pyxu_b200/csrc/pxb_stencil_tma.cuh): the device's per-thread bodies, with the
box load emulated as a zero-filled gather, against (a) the generic per-sample bodies, (b) fixtures from the real
reference for the 'constant'-mode cases, (c) the adjoint identity <Sx, y> == <x, S^T y>.
"""
import ctypes as C

import numpy as np
import pytest

import cases
import emu_util as E
import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K
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


def test_rank_one_dense_kernels_take_the_separable_passes():
    g9 = gauss(9, 1.7)
    op = pxo.Stencil(arg_shape=(40, 64), kernel=np.outer(g9, g9), center=(4, 4), mode="constant")
    assert op._tiled_plan(False)[1][0] == "sep" and op._tiled_plan(True)[1][0] == "sep"
    x = np.random.default_rng(0).standard_normal(op.dim)
    for adj in (False, True):
        assert relerr(E.stencil_run_tiled(op, x, adj), E.stencil_run(op, x, adj)) < 1e-13
    full = pxo.Stencil(arg_shape=(40, 64), kernel=np.outer(g9, g9) + 0.01 * np.eye(9), center=(4, 4), mode="constant")
    assert full._tiled_plan(False)[1][0] == "dense"
    f32 = pxo.Stencil(arg_shape=(40, 64), kernel=np.outer(g9, g9).astype(np.float32), center=(4, 4), mode="constant")
    assert f32._tiled_plan(False)[1][0] == "sep"  # rank 1 at fp32 resolution


def test_rank_one_dense_3d_kernels_are_split_into_three_factors():
    g = [gauss(7, 1.2), gauss(5, 1.0), gauss(7, 1.5)]
    for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        dense = np.einsum("i,j,k->ijk", *g).astype(dt)
        op = pxo.Stencil(arg_shape=(12, 20, 32), kernel=dense, center=(3, 2, 3), mode="constant")
        for adj in (False, True):
            plan = op._tiled_plan(adj)
            assert plan is not None and plan[0] is not None and plan[1][0] == "sep"  # a factor along axis 0 + separable in-plane part
            assert op._desc3d(K.F64 if dt == np.float64 else K.F32, adj, 1) is not None
        # the three factors multiply back to the kernel
        fac = [k3.reshape(-1) for k3, _ in op._rank1_split(*op._passes(False)[0])]
        assert relerr(np.einsum("i,j,k->ijk", *fac), dense) < tol
        bumped = dense.copy()
        bumped[1, 2, 3] *= 1.05
        assert pxo.Stencil(arg_shape=(12, 20, 32), kernel=bumped, center=(3, 2, 3), mode="constant")._tiled_plan(False) is None
        pair = np.einsum("j,k->jk", g[1], g[2])[None].astype(dt)  # (1, 5, 7): no factor along axis 0, the 2-D rule applies
        assert pxo.Stencil(arg_shape=(12, 20, 32), kernel=pair, center=(0, 2, 3), mode="constant")._tiled_plan(False)[0] is None


def test_tiled_not_applicable():
    op = pxo.Stencil(arg_shape=(20, 24), kernel=np.ones((3, 3)), center=(1, 1), mode="reflect")
    assert op._tiled_plan(False) is None
    full = np.ones((3, 3, 3))
    full[0, 1, 2] = 2.0  # not an outer product
    op = pxo.Stencil(arg_shape=(6, 20, 24), kernel=full, center=(1, 1, 1), mode="constant")
    assert op._tiled_plan(False) is None  # dense 3-D kernels of full rank keep the generic path ...
    op = pxo.Stencil(arg_shape=(6, 20, 24), kernel=np.ones((3, 3, 3)), center=(1, 1, 1), mode="constant")
    assert op._tiled_plan(False) is not None  # ... an outer product handed over as an array takes the separable single pass
    op = pxo.Stencil(arg_shape=(20, 22), kernel=np.ones((3, 3), dtype=np.float32), center=(1, 1), mode="constant")
    x = np.zeros(op.dim, dtype=np.float32)
    assert E.stencil_run_tiled(op, x, False) is None  # fp32: the last axis must be a multiple of 4 samples


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fista_two_pass_form(dense, dtype):
    """pxb_stencil2d_fista (which = 0, 1) == extrapolate, A y + shift, A^T r, soft-threshold, RelError sums -- step by step."""
    import ctypes as C

    from pyxu_b200 import _cabi as K

    rng = np.random.default_rng(4)
    shape, batch = (37, 72), 3
    g1, g2 = gauss(5, 1.0), gauss(5, 1.3)
    kern = (np.outer(g1, g2) + 0.03 * np.eye(5)).astype(dtype) if dense else [g1.astype(dtype), g2.astype(dtype)]  # full rank: dense instance
    op = pxo.Stencil(arg_shape=shape, kernel=kern, center=(2, 1), mode="constant")
    x = rng.standard_normal((batch, op.dim)).astype(dtype)
    xp = rng.standard_normal((batch, op.dim)).astype(dtype)
    shift = rng.standard_normal((batch, op.dim)).astype(dtype)
    alpha, tau, lam = 0.5, 0.7, 0.05
    tol = 1e-13 if dtype == np.float64 else 3e-6
    for a in (0.0, 0.3):
        y = (1 + a) * x - a * xp
        r_ref = 2 * alpha * (E.stencil_run(op, y, False) + shift)
        v = y - tau * E.stencil_run(op, r_ref, True)
        x_ref = np.sign(v) * np.maximum(np.abs(v) - lam * tau, 0)
        # descriptors exactly as PGD._setup_fista_fused builds them (host arrays instead of device tensors)
        fw = _desc(op, x, False, 2 * alpha, 2 * alpha, shift)
        bw = _desc(op, x, True, -tau, 0.0, None)
        st = K.FistaStep()
        r = np.empty_like(x)
        nrm = np.zeros(2 * batch)
        st.x, st.x_prev, st.r, st.a, st.tau = x.ctypes.data, xp.ctypes.data, r.ctypes.data, a, tau
        st.g = K.ProxSpec(K.PROX_L1, 0, lam, 0.0)
        st.norms, st.imgs_per_row = nrm.ctypes.data, 1
        assert E.lib().emu_stencil2d_fista(C.byref(fw), C.byref(st), 0, E.p(r)) == 0
        assert relerr(r, r_ref) < tol
        out = xp.copy()  # x_new overwrites the x_prev buffer
        st.x_prev = out.ctypes.data
        assert E.lib().emu_stencil2d_fista(C.byref(bw), C.byref(st), 1, E.p(out)) == 0
        assert relerr(out, x_ref) < tol
        num = ((x_ref.astype(np.float64) - x) ** 2).sum(axis=1)
        den = (x.astype(np.float64) ** 2).sum(axis=1)
        assert np.allclose(nrm[0::2], num, rtol=1e-4 if dtype == np.float32 else 1e-10) and np.allclose(nrm[1::2], den, rtol=1e-6)


def _desc(op, like, adjoint, alpha, beta, add):
    """pxb_stencil2d descriptor on host arrays, mirroring Stencil._tiled_desc."""
    from pyxu_b200 import _cabi as K

    axis0, inplane, scale = op._tiled_plan(adjoint)
    assert axis0 is None
    D = len(op._arg_shape)
    shape3 = (1,) * (3 - D) + op._arg_shape
    d = K.Stencil2D()
    d.dtype, d.nimg = E.dcode(like), max(1, like.size // op.dim) * shape3[0]
    d.shape[0], d.shape[1] = shape3[1], shape3[2]
    if inplane[0] == "dense":
        _, k2d, c1, c2 = inplane
        d._keep = np.ascontiguousarray(k2d.reshape(-1), dtype=like.dtype)
        d.dense, d.coef = 1, d._keep.ctypes.data
        d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k2d.shape[0], k2d.shape[1], c1, c2
    else:
        _, t1, c1, t2, c2 = inplane
        d.ksize[0], d.ksize[1], d.center[0], d.center[1] = t1.size, t2.size, c1, c2
        for i, v in enumerate(t1):
            d.coef1[i] = float(v)
        for i, v in enumerate(t2):
            d.coef2[i] = float(v)
    d.alpha, d.beta = alpha * scale, beta
    if add is not None:
        d.add, d.add_period = add.ctypes.data, add.size
    return d


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k0,c0", [(3, 1), (5, 0), (7, 3), (9, 8)])
def test_single_pass_3d_separable(dtype, k0, c0):
    """pxb_stencil3d (marching, register ring) == the generic chain of 1-D stencils; apply, adjoint, epilogue, several chunks."""
    rng = np.random.default_rng(k0)
    shape = (21, 19, 140)  # ragged tiles in both in-plane directions
    kern = [rng.standard_normal(k0).astype(dtype), gauss(5, 1.0).astype(dtype), gauss(7, 1.3).astype(dtype)]
    op = pxo.Stencil(arg_shape=shape, kernel=kern, center=(c0, 1, 3), mode="constant")
    x = rng.standard_normal((2, op.dim)).astype(dtype)
    y = rng.standard_normal((2, op.dim)).astype(dtype)
    tol = 1e-13 if dtype == np.float64 else 5e-6
    for adj in (False, True):
        ref = E.stencil_run(op, x, adj)
        out = E.stencil3d_run(op, x, adj)
        assert out is not None and relerr(out, ref) < tol, (adj, relerr(out, ref))
    out = E.stencil3d_run(op, x, False, alpha=0.5, beta=-1.0, add=y)
    assert relerr(out, 0.5 * E.stencil_run(op, x, False) - y) < tol


def test_single_pass_3d_slab_cuts():
    """z-slab cuts: every slab filtered on its own (ghost planes = the neighbours' planes) == the whole volume."""
    from pyxu_b200 import _cabi as K

    rng = np.random.default_rng(2)
    shape = (23, 17, 72)
    kern = [gauss(7, 1.2), gauss(5, 1.0), gauss(7, 1.3)]
    op = pxo.Stencil(arg_shape=shape, kernel=kern, center=(3, 2, 3), mode="constant")
    x = rng.standard_normal(shape)
    plane = shape[1] * shape[2]
    H = 3
    for adj in (False, True):
        ref = E.stencil_run(op, x.reshape(-1), adj).reshape(shape)
        parts, got = [(0, 8), (8, 16), (16, 23)], []
        for r, (a, b) in enumerate(parts):
            n0 = b - a
            buf = np.zeros((n0 + 2 * H,) + shape[1:])
            lo, hi = max(0, a - H), min(shape[0], b + H)
            buf[H - (a - lo) : H + n0 + (hi - b)] = x[lo:hi]
            out = np.full_like(buf, np.nan)
            slab = K.Slab(1 if r > 0 else 0, 1 if r < len(parts) - 1 else 0, H, n0 + 2 * H)
            ptr = lambda t: C.c_void_p(t.ctypes.data + 8 * H * plane)
            import ctypes as C
            assert E.stencil3d_run(op, buf, adj, slab=slab, shape0=n0, raw_ptrs=(ptr(buf), ptr(out))) == 0
            got.append(out[H : H + n0])
        assert relerr(np.concatenate(got, axis=0), ref) < 1e-13, adj


# ---------------------------------------------------------------------------------------------------------
# Dense (full-rank) 3-D kernels in one marching pass (pxb_stencil3d_dense.cuh): every staged input plane scattered into the
# register accumulators of the K output planes it contributes to.  Against the gather kernels' bodies and the NumPy oracle.
# ---------------------------------------------------------------------------------------------------------
DENSE3D_CASES = [
    # (arg_shape, kernel extents, center)
    ((21, 19, 140), (7, 7, 7), (3, 3, 3)),   # the "7x7x7 PSF" of BASELINE configs[4]; ragged tiles in both in-plane directions
    ((9, 35, 264), (7, 7, 7), (0, 6, 1)),    # off-centre everywhere (windows start at any column); three tiles along the rows
    ((12, 17, 72), (5, 5, 5), (2, 2, 2)),
    ((11, 18, 16), (5, 4, 5), (4, 0, 3)),    # embedded in the 5-cube (zero taps behind the kernel's own)
    ((6, 20, 24), (3, 3, 3), (1, 1, 1)),
    ((3, 9, 8), (7, 6, 5), (6, 2, 0)),       # fewer planes than taps: most kernel planes never meet a source plane
    ((10, 16, 12), (2, 3, 3), (1, 0, 2)),    # embedded in the 3-cube
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(DENSE3D_CASES)))
def test_dense_3d_marching_kernel(dtype, case):
    from oracle import pyxu_oracle as orc

    shape, ks, cen = DENSE3D_CASES[case]
    rng = np.random.default_rng(100 + case)
    kern = rng.standard_normal(ks).astype(dtype)
    op = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
    x = rng.standard_normal((2, op.dim)).astype(dtype)
    y = rng.standard_normal(op.dim).astype(dtype)
    tol = 1e-13 if dtype == np.float64 else 2e-5
    oref = orc.Stencil(shape, kern.astype(np.float64), cen, "constant")
    for adj in (False, True):
        ref = E.stencil_run(op, x, adj)
        for chunk in (0, 4):  # the launcher's chunk length (one chunk at these sizes) / several chunks, the last one ragged
            out = E.stencil3d_dense_run(op, x, adj, chunk=chunk)
            assert out is not None and not np.isnan(out).any() and relerr(out, ref) < tol, (adj, chunk, relerr(out, ref))
        oo = oref.adjoint(x.astype(np.float64)) if adj else oref.apply(x.astype(np.float64))
        assert relerr(out, oo) < tol
    out = E.stencil3d_dense_run(op, x, False, alpha=0.5, beta=-1.0, add=y, chunk=5)  # the operand of one item serves the stack
    assert relerr(out, 0.5 * E.stencil_run(op, x, False) - y) < tol
    # adjoint identity <S x, z> = <x, S^T z> through the marching kernel alone
    z = rng.standard_normal((2, op.dim)).astype(dtype)
    lhs = np.sum(E.stencil3d_dense_run(op, x, False).astype(np.float64) * z)
    rhs = np.sum(x * E.stencil3d_dense_run(op, z, True).astype(np.float64))
    assert abs(lhs - rhs) < (1e-11 if dtype == np.float64 else 2e-4) * (1 + abs(lhs))


def test_dense_3d_marching_kernel_envelope():
    """Outside the envelope the descriptor is not even built: folding modes, more than 7 taps, kernels that leave most of the cube empty."""
    rng = np.random.default_rng(0)
    mk = lambda ks, mode="constant": pxo.Stencil(arg_shape=(12, 16, 16), kernel=rng.standard_normal(ks), center=(0, 0, 0), mode=mode)
    assert mk((7, 7, 7))._desc3d_dense(1, False, 1) is not None
    assert mk((3, 3, 3), "reflect")._desc3d_dense(1, False, 1) is None
    assert mk((9, 3, 3))._desc3d_dense(1, False, 1) is None
    assert mk((7, 2, 3))._desc3d_dense(1, False, 1) is None     # 42 of 343 taps
    assert mk((1, 3, 3))._desc3d_dense(1, False, 1) is None     # a 2-D kernel: the tiled 2-D pass
    # last axis not a multiple of the vector width: the launcher declines (PXB_ENOSUP), callers fall back
    op = pxo.Stencil(arg_shape=(5, 6, 7), kernel=rng.standard_normal((3, 3, 3)), center=(1, 1, 1), mode="constant")
    assert E.stencil3d_dense_run(op, rng.standard_normal(op.dim), False) is None


def test_dense_3d_marching_kernel_slab_cuts():
    """z-slab cuts: every slab filtered on its own (ghost planes = the neighbours' planes) == the whole volume."""
    import ctypes as C

    from pyxu_b200 import _cabi as K

    rng = np.random.default_rng(3)
    shape = (23, 17, 72)
    op = pxo.Stencil(arg_shape=shape, kernel=rng.standard_normal((7, 5, 6)), center=(2, 2, 3), mode="constant")
    x = rng.standard_normal(shape)
    plane = shape[1] * shape[2]
    H = 4
    for adj in (False, True):
        ref = E.stencil_run(op, x.reshape(-1), adj).reshape(shape)
        parts, got = [(0, 8), (8, 16), (16, 23)], []
        for r, (a, b) in enumerate(parts):
            n0 = b - a
            buf = np.zeros((n0 + 2 * H,) + shape[1:])
            lo, hi = max(0, a - H), min(shape[0], b + H)
            buf[H - (a - lo) : H + n0 + (hi - b)] = x[lo:hi]
            out = np.full_like(buf, np.nan)
            slab = K.Slab(1 if r > 0 else 0, 1 if r < len(parts) - 1 else 0, H, n0 + 2 * H)
            ptr = lambda t: C.c_void_p(t.ctypes.data + 8 * H * plane)
            assert E.stencil3d_dense_run(op, buf, adj, chunk=3, slab=slab, shape0=n0, raw_ptrs=(ptr(buf), ptr(out))) == 0
            got.append(out[H : H + n0])
        assert relerr(np.concatenate(got, axis=0), ref) < 1e-13, adj


# ---------------------------------------------------------------------------------------------------------
# Folding boundary modes through the tiled kernel (Stencil._run_padded): Pad (pxb_pad2d) -> tiled S0 on the padded array /
# tiled S0^T onto the padded extent -> Pad^T (pxb_pad2d_adjoint).  Against the gather kernels, the fixtures of the real
# reference and the adjoint identity.
# ---------------------------------------------------------------------------------------------------------
PADDED_CASES = [
    # (arg_shape, kernel, center, mode)
    ((70, 300), [gauss(9, 1.7), gauss(9, 1.7)], (4, 4), "reflect"),                       # config[1]'s blur with a folding mode
    ((70, 300), np.outer(gauss(9, 1.7), gauss(9, 1.7)) + 0.01 * np.arange(81.0).reshape(9, 9), (4, 4), "symmetric"),  # dense 9x9
    ((40, 72), [gauss(5, 1.0), gauss(7, 1.3)], (0, 6), ("wrap", "symmetric")),           # one-sided pads
    ((33, 64), np.arange(1.0, 16.0).reshape(3, 5) / 7, (2, 0), ("constant", "wrap")),     # a 'constant' axis next to a folding one
    ((33, 64), np.arange(1.0, 16.0).reshape(3, 5) / 7, (1, 4), ("edge", "reflect")),
    ((5, 9, 64), [gauss(3, 1.0), gauss(5, 1.0), gauss(9, 1.5)], (1, 2, 4), ("reflect", "symmetric", "wrap")),   # folding factor along axis 0
    ((5, 9, 64), [gauss(3, 1.0), gauss(5, 1.0), gauss(9, 1.5)], (2, 0, 8), ("constant", "edge", "reflect")),
    ((4, 37, 72), np.arange(1.0, 10.0).reshape(1, 3, 3), (0, 1, 1), ("constant", "wrap", "edge")),             # dense 2-D kernel on a stack of planes
    ((128,), [gauss(9, 2.0)], (4,), "symmetric"),
    ((4, 8), np.arange(1.0, 13.0).reshape(3, 4), (2, 3), "wrap"),                          # pads as wide as the mode allows on a tiny image
    ((6, 9, 64), [gauss(3, 1.0), gauss(5, 1.0), gauss(9, 1.5)], (0, 2, 4), ("symmetric", "constant", "constant")),  # only axis 0 folds
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ci", range(len(PADDED_CASES)))
def test_padded_tiled_equals_generic(ci, dtype):
    shape, kern, cen, mode = PADDED_CASES[ci]
    k = [np.asarray(_, dtype=dtype) for _ in kern] if isinstance(kern, list) else np.asarray(kern, dtype=dtype)
    op = pxo.Stencil(arg_shape=shape, kernel=k, center=cen, mode=mode)
    rng = np.random.default_rng(100 + ci)
    x = rng.standard_normal((2, op.dim)).astype(dtype)
    tol = 1e-13 if dtype == np.float64 else 3e-6
    for adj in (False, True):
        ref = E.stencil_run(op, x, adj)
        out = E.stencil_run_padded(op, x, adj)
        assert out is not None and np.isfinite(out).all() and relerr(out, ref) < tol, (ci, adj, relerr(out, ref))
    y = rng.standard_normal((2, op.dim)).astype(dtype)
    lhs, rhs = np.vdot(E.stencil_run_padded(op, x, False).astype(np.float64), y), np.vdot(x, E.stencil_run_padded(op, y, True).astype(np.float64))
    assert abs(lhs - rhs) < (1e-11 if dtype == np.float64 else 2e-4) * (1 + abs(lhs))


def test_padded_tiled_epilogue():
    """alpha * S(x) + beta * add in both directions (the '- y' of a data term rides in the stencil pass / in the fold)."""
    rng = np.random.default_rng(5)
    op = pxo.Stencil(arg_shape=(50, 140), kernel=[gauss(9, 1.7), gauss(5, 1.0)], center=(4, 1), mode=("reflect", "wrap"))
    x, y = rng.standard_normal((3, op.dim)), rng.standard_normal(op.dim)  # `add` broadcast over the stack
    for adj in (False, True):
        out = E.stencil_run_padded(op, x, adj, alpha=0.5, beta=-1.5, add=y)
        assert relerr(out, 0.5 * E.stencil_run(op, x, adj) - 1.5 * y) < 1e-13


@pytest.mark.parametrize("case", [c for c in cases.STENCIL_CASES if c["mode"] != "constant"], ids=lambda c: c["name"])
def test_padded_tiled_golden(case):
    """Fixtures produced by the real reference (tests/golden/make_golden.py); cases whose last axis is not a multiple of
    the vector width stay on the gather kernels."""
    g = golden("stencil.npz")
    n = case["name"]
    op = cases.make_stencil(type("ns", (), {"operator": pxo}), case)
    got = [E.stencil_run_padded(op, g[f"{n}/x"], False), E.stencil_run_padded(op, g[f"{n}/y"], True)]
    if case["arg_shape"][-1] % 2 or len(case["arg_shape"]) == 3 and not isinstance(case["kernel"], list):
        assert got[0] is None and got[1] is None
        return
    assert relerr(got[0], g[f"{n}/apply"]) < 1e-13 and relerr(got[1], g[f"{n}/adjoint"]) < 1e-13


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pad2d_is_numpy_pad_and_its_transpose(dtype):
    """pxb_pad2d reproduces numpy.pad (what the reference's Pad calls, pad.py:252-302) bit for bit for every mode and pad width the
    reference admits (pad.py:217-229), filler columns included; pxb_pad2d_adjoint is its exact transpose."""
    rng = np.random.default_rng(17)
    names = ["constant", "wrap", "reflect", "symmetric", "edge"]
    for trial in range(40):
        n1, n2, nimg = int(rng.integers(1, 9)), int(rng.integers(1, 12)), int(rng.integers(1, 4))
        modes = [str(rng.choice(names)) for _ in range(2)]
        lo, hi = [0, 0], [0, 0]
        for a, n in enumerate((n1, n2)):
            lim = {"constant": 5, "edge": 5, "wrap": n, "symmetric": n, "reflect": n - 1}[modes[a]]
            lo[a], hi[a] = int(rng.integers(0, lim + 1)), int(rng.integers(0, lim + 1))
        org = [lo[0] + int(rng.integers(0, 2)), lo[1] + int(rng.integers(0, 4))]
        ext = [org[0] + n1 + hi[0] + int(rng.integers(0, 2)), org[1] + n2 + hi[1] + int(rng.integers(0, 4))]
        d = E.K.Pad2D()
        d.dtype, d.nimg = E.dcode(np.zeros(1, dtype=dtype)), nimg
        d.shape[0], d.shape[1], d.ext_shape[0], d.ext_shape[1] = n1, n2, ext[0], ext[1]
        for a in (0, 1):
            d.org[a], d.lo[a], d.hi[a], d.mode[a] = org[a], lo[a], hi[a], E.K.MODES[modes[a]]
        x = rng.standard_normal((nimg, n1, n2)).astype(dtype)
        got = np.full((nimg, ext[0], ext[1]), np.nan, dtype=dtype)
        E.lib().emu_pad2d(C.byref(d), E.p(x), E.p(got))
        want = np.zeros_like(got)
        ref = x
        for a in (0, 1):  # numpy.pad axis by axis, as Pad.apply does for mixed modes (pad.py:270-302)
            pw = [(0, 0)] * 3
            pw[1 + a] = (lo[a], hi[a])
            ref = np.pad(ref, pw, mode=modes[a])
        want[:, org[0] - lo[0] : org[0] + n1 + hi[0], org[1] - lo[1] : org[1] + n2 + hi[1]] = ref
        assert np.array_equal(got, want), (trial, modes, lo, hi)
        y = rng.standard_normal(got.shape).astype(dtype)
        back = np.full_like(x, np.nan)
        E.lib().emu_pad2d_adjoint(C.byref(d), E.p(y), E.p(back), 1.0, 0.0, None, 0)
        yin = np.zeros_like(y)  # only the padded extent takes part
        yin[:, org[0] - lo[0] : org[0] + n1 + hi[0], org[1] - lo[1] : org[1] + n2 + hi[1]] = y[:, org[0] - lo[0] : org[0] + n1 + hi[0], org[1] - lo[1] : org[1] + n2 + hi[1]]
        lhs, rhs = np.vdot(got.astype(np.float64), yin), np.vdot(x.astype(np.float64), back)
        assert abs(lhs - rhs) < (1e-12 if dtype == np.float64 else 1e-4) * (1 + abs(lhs)), (trial, modes, lo, hi)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_axis0_streaming_kernel_with_folding_modes(dtype):
    """pxb_stencil_axis0_fold (streaming register-ring kernel, folding boundary mode along axis 0): apply = S o Pad, adjoint =
    Pad^T o S0^T, against the gather kernels for every mode, tap count, centre and chunk length."""
    rng = np.random.default_rng(23)
    vec = 2 if dtype == np.float64 else 4
    for trial in range(60):
        k0 = int(rng.integers(2, 10))
        mode = str(rng.choice(["wrap", "reflect", "symmetric", "edge", "constant"]))
        nmin = {"wrap": k0 - 1, "symmetric": k0 - 1, "reflect": k0, "edge": 1, "constant": 1}[mode]
        n0 = int(rng.integers(max(nmin, 2), max(nmin, 2) + 9))
        shape = (n0, int(rng.integers(1, 4)), vec * int(rng.integers(1, 4)))
        c0 = int(rng.integers(0, k0))
        taps = rng.standard_normal(k0)
        op = pxo.Stencil(arg_shape=shape, kernel=[taps.astype(dtype), np.ones(1, dtype=dtype), np.ones(1, dtype=dtype)], center=(c0, 0, 0),
                         mode=(mode, "constant", "constant"))
        batch = int(rng.integers(1, 3))
        x = rng.standard_normal((batch, op.dim)).astype(dtype)
        for adj in (False, True):
            ref = E.stencil_run(op, x, adj)
            k, c = (taps[::-1].copy(), k0 - 1 - c0) if adj else (taps, c0)  # the transposed form takes the reversed taps
            out = np.full_like(x, np.nan)
            rc = E.lib().emu_stencil_axis0_fold(E.dcode(x), batch, (C.c_int64 * 3)(*shape), k0, c, (C.c_double * k0)(*k), E.K.MODES[mode], int(adj),
                                                E.p(x), E.p(out), int(rng.choice([0, 1, 2, 5])))
            assert rc == 0
            assert relerr(out, ref) < (1e-13 if dtype == np.float64 else 3e-6), (trial, shape, k0, c0, mode, adj)
