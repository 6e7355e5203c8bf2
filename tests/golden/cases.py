"""
Problem definitions shared by the golden-vector generator (run against the real reference) and by
the parity tests (run against the oracle and against pyxu_b200 on the GPU).

Builders take a namespace `ns` exposing `.operator`, `.solver`, `.stop` with the reference's names,
so the same code drives `pyxu` (reference) and `pyxu_b200` (this repo).
"""
import numpy as np


def gaussian_1d(size, sigma):
    t = np.arange(size) - (size - 1) / 2
    k = np.exp(-0.5 * (t / sigma) ** 2)
    return k / k.sum()


# ------------------------------------------------------------------------------------------------
# Stencil cases: the 11 parameterisations of the reference's own test-suite
# (reference: src/pyxu_tests/operator/linop/test_stencil.py:19-84) + 3-D and Convolve cases.
# ------------------------------------------------------------------------------------------------
STENCIL_CASES = [
    dict(name="1d_constant", arg_shape=(10,), kernel=np.arange(1, 7.0), center=(0,), mode="constant"),
    dict(name="1d_edge", arg_shape=(10,), kernel=np.arange(1, 7.0), center=(1,), mode="edge"),
    dict(name="1d_wrap", arg_shape=(10,), kernel=np.arange(1, 7.0), center=(2,), mode="wrap"),
    dict(name="1d_reflect", arg_shape=(10,), kernel=np.arange(1, 7.0), center=(3,), mode="reflect"),
    dict(name="1d_symmetric", arg_shape=(10,), kernel=np.arange(1, 7.0), center=(4,), mode="symmetric"),
    dict(name="2d_constant", arg_shape=(10, 11), kernel=np.arange(1, 9.0).reshape(2, 4), center=(0, 3), mode="constant"),
    dict(name="2d_wrap_reflect", arg_shape=(10, 11), kernel=np.arange(1, 9.0).reshape(2, 4), center=(1, 2), mode=("wrap", "reflect")),
    dict(name="2d_edge_symmetric", arg_shape=(10, 11), kernel=np.arange(1, 9.0).reshape(2, 4), center=(1, 1), mode=("edge", "symmetric")),
    dict(name="2d_sep_constant", arg_shape=(10, 11), kernel=[np.arange(1, 7.0), np.arange(2, 5.0)], center=(3, 0), mode="constant"),
    dict(name="2d_sep_edge_wrap", arg_shape=(10, 11), kernel=[np.arange(1, 7.0), np.arange(2, 5.0)], center=(2, 1), mode=("edge", "wrap")),
    dict(name="2d_sep_reflect_symmetric", arg_shape=(10, 11), kernel=[np.arange(1, 7.0), np.arange(2, 5.0)], center=(3, 2), mode=("reflect", "symmetric")),
    # extra coverage -----------------------------------------------------------------------------
    dict(name="3d_constant", arg_shape=(6, 7, 9), kernel=np.arange(1, 19.0).reshape(2, 3, 3) / 7, center=(1, 1, 0), mode="constant"),
    dict(name="3d_mixed", arg_shape=(6, 7, 9), kernel=np.cos(np.arange(27.0)).reshape(3, 3, 3), center=(1, 2, 0), mode=("reflect", "wrap", "edge")),
    dict(name="3d_sep_mixed", arg_shape=(5, 8, 7), kernel=[np.r_[1.0, -2, 1], np.r_[0.5, 0.25], np.r_[3.0, 1, 2, 4]], center=(1, 0, 3), mode=("symmetric", "constant", "wrap")),
    dict(name="2d_gauss9", arg_shape=(16, 20), kernel=np.outer(gaussian_1d(9, 1.7), gaussian_1d(9, 1.7)), center=(4, 4), mode="reflect"),
    dict(name="2d_sep_gauss9", arg_shape=(16, 20), kernel=[gaussian_1d(9, 1.7), gaussian_1d(9, 1.7)], center=(4, 4), mode="symmetric"),
    dict(name="2d_wrap_full", arg_shape=(4, 5), kernel=np.arange(1, 21.0).reshape(4, 5) / 10, center=(3, 0), mode="wrap"),  # pad == N
    dict(name="1d_conv", arg_shape=(12,), kernel=np.r_[1.0, 2, -3, 0.5], center=(1,), mode="constant", conv=True),
    dict(name="2d_conv_reflect", arg_shape=(9, 8), kernel=np.arange(1, 13.0).reshape(3, 4), center=(2, 1), mode="reflect", conv=True),
]

GRADIENT_CASES = [
    dict(name="g2_fwd_constant", arg_shape=(10, 11), mode="constant", diff_kwargs={}),
    dict(name="g2_bwd_reflect", arg_shape=(10, 11), mode="reflect", diff_kwargs=dict(scheme="backward")),
    dict(name="g2_ctr_wrap", arg_shape=(9, 12), mode="wrap", diff_kwargs=dict(scheme="central")),
    dict(name="g2_fwd_sym_edge", arg_shape=(7, 8), mode=("symmetric", "edge"), diff_kwargs=dict(sampling=0.5)),
    dict(name="g3_fwd_constant", arg_shape=(6, 7, 8), mode="constant", diff_kwargs={}),
    dict(name="g3_fwd_mixed", arg_shape=(6, 7, 8), mode=("wrap", "reflect", "edge"), diff_kwargs=dict(sampling=(1.0, 2.0, 0.5))),
    dict(name="g3_ctr_acc2", arg_shape=(6, 7, 8), mode="symmetric", diff_kwargs=dict(scheme="central", accuracy=2)),
    dict(name="g1_fwd_acc2", arg_shape=(13,), mode="constant", diff_kwargs=dict(scheme="forward", accuracy=2)),
]


def make_stencil(ns, case, dtype=np.float64):
    k = case["kernel"]
    k = [np.asarray(_, dtype=dtype) for _ in k] if isinstance(k, list) else np.asarray(k, dtype=dtype)
    klass = ns.operator.Convolve if case.get("conv") else ns.operator.Stencil
    return klass(arg_shape=case["arg_shape"], kernel=k, center=case["center"], mode=case["mode"])


def make_gradient(ns, case, dtype=np.float64, **kw):
    return ns.operator.Gradient(arg_shape=case["arg_shape"], mode=case["mode"], dtype=dtype, **case["diff_kwargs"], **kw)


# ------------------------------------------------------------------------------------------------
# Solver cases.
# ------------------------------------------------------------------------------------------------
def phantom(shape, seed=0, noise=0.1, dtype=np.float64):
    """Piecewise-constant blocks + noise (synthetic phantom)."""
    rng = np.random.default_rng(seed)
    x = np.zeros(shape)
    for _ in range(6):
        lo = [rng.integers(0, max(1, n - 2)) for n in shape]
        hi = [rng.integers(l + 1, n + 1) for l, n in zip(lo, shape)]
        x[tuple(slice(l, h) for l, h in zip(lo, hi))] += rng.uniform(0.2, 1.0)
    y = x + noise * rng.standard_normal(shape)
    return x.astype(dtype), y.astype(dtype)


def build_tv_denoise(ns, y, arg_shape, lam=0.1, mode="constant", positivity=True, solver="PD3O", dtype=np.float64, **kw):
    """config[0] / config[3] of BASELINE.json: f = 1/2||x-y||^2, g = i_+, h = lam*L21, K = Gradient."""
    pxo, N, D = ns.operator, int(np.prod(arg_shape)), len(arg_shape)
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-np.asarray(y, dtype=dtype).reshape(-1))
    K = pxo.Gradient(arg_shape=arg_shape, mode=mode, dtype=dtype)
    h = lam * pxo.L21Norm(arg_shape=(D, *arg_shape), l2_axis=(0,))
    g = pxo.PositiveOrthant(dim=N) if positivity else None
    klass = getattr(ns.solver, solver)
    return klass(f=f, g=g, h=h, K=K, show_progress=False, **kw)


def build_tv_deblur(ns, y, arg_shape, blur_kernel, blur_center, lam=0.05, mode="constant", blur_mode="constant",
                    positivity=False, solver="CondatVu", dtype=np.float64, **kw):
    """config[1] / config[4]: f = 1/2||A x - y||^2 with A a Stencil blur, h = lam*L21 o Gradient."""
    pxo, N, D = ns.operator, int(np.prod(arg_shape)), len(arg_shape)
    bk = [np.asarray(_, dtype=dtype) for _ in blur_kernel] if isinstance(blur_kernel, list) else np.asarray(blur_kernel, dtype=dtype)
    A = pxo.Stencil(arg_shape=arg_shape, kernel=bk, center=blur_center, mode=blur_mode)
    f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-np.asarray(y, dtype=dtype).reshape(-1))) * A
    K = pxo.Gradient(arg_shape=arg_shape, mode=mode, dtype=dtype)
    h = lam * pxo.L21Norm(arg_shape=(D, *arg_shape), l2_axis=(0,))
    g = pxo.PositiveOrthant(dim=N) if positivity else None
    klass = getattr(ns.solver, solver)
    # diff-Lipschitz constant of f: ||A||^2 <= A.lipschitz^2 (the reference cannot infer it for Quadratic o LinOp
    # without an SVD run, so it is handed over explicitly -- same value for both implementations).
    return klass(f=f, g=g, h=h, K=K, beta=float(A.lipschitz) ** 2, show_progress=False, **kw), A


def build_l1_deconv(ns, y, arg_shape, blur_kernel, blur_center, lam=0.02, blur_mode="constant", dtype=np.float64, **kw):
    """config[2]: PGD/FISTA, f = 1/2||A x - y||^2 (A = 5x5 Stencil), g = lam*L1."""
    pxo, N = ns.operator, int(np.prod(arg_shape))
    A = pxo.Stencil(arg_shape=arg_shape, kernel=np.asarray(blur_kernel, dtype=dtype), center=blur_center, mode=blur_mode)
    f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-np.asarray(y, dtype=dtype))) * A
    g = lam * pxo.L1Norm(dim=N)
    # fit() must be given tau = 1 / A.lipschitz^2 (see build_tv_deblur).
    return ns.solver.PGD(f=f, g=g, show_progress=False, **kw), A
