"""
Generate golden vectors by running the REAL reference (AdriaJ/pyxu at /root/reference) in the build
container.  Usage:  python tests/golden/make_golden.py      (writes tests/golden/*.npz)

The committed .npz files are what the tests read; this script is kept so the vectors can be
re-derived.  /root/reference is never needed at test time.
"""
import os
import sys
import warnings
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import _ref_import  # noqa: E402
import cases  # noqa: E402

warnings.filterwarnings("ignore")


def stencils(ns):
    out = {}
    for case in cases.STENCIL_CASES:
        op = cases.make_stencil(ns, case)
        rng = np.random.default_rng(zlib.crc32(case["name"].encode()))
        N = int(np.prod(case["arg_shape"]))
        x = rng.standard_normal((2, N))
        y = rng.standard_normal((2, N))
        n = case["name"]
        out[f"{n}/x"], out[f"{n}/y"] = x, y
        out[f"{n}/apply"] = op.apply(x)
        out[f"{n}/adjoint"] = op.adjoint(y)
        out[f"{n}/lipschitz"] = np.float64(op.lipschitz)
    np.savez_compressed(os.path.join(HERE, "stencil.npz"), **out)
    print("stencil.npz", len(out))


def gradients(ns):
    out = {}
    for case in cases.GRADIENT_CASES:
        op = cases.make_gradient(ns, case)
        rng = np.random.default_rng(zlib.crc32(case["name"].encode()))
        N = int(np.prod(case["arg_shape"]))
        D = len(case["arg_shape"])
        x = rng.standard_normal((2, N))
        y = rng.standard_normal((2, D * N))
        n = case["name"]
        out[f"{n}/x"], out[f"{n}/y"] = x, y
        out[f"{n}/apply"] = op.apply(x)
        out[f"{n}/adjoint"] = op.adjoint(y)
        out[f"{n}/lipschitz"] = np.float64(op.lipschitz)
    np.savez_compressed(os.path.join(HERE, "gradient.npz"), **out)
    print("gradient.npz", len(out))


def funcs(ns):
    pxo = ns.operator
    rng = np.random.default_rng(7)
    N = 60
    x = rng.standard_normal((3, N)) * 2
    out = {"x": x}
    for tau in (0.3, 1.7):
        t = f"{tau}"
        out[f"l1/prox/{t}"] = pxo.L1Norm(dim=N).prox(x, tau)
        out[f"l1s/prox/{t}"] = (0.4 * pxo.L1Norm(dim=N)).prox(x, tau)
        out[f"l1/fprox/{t}"] = pxo.L1Norm(dim=N).fenchel_prox(x, tau)
        out[f"posl1/prox/{t}"] = pxo.PositiveL1Norm(dim=N).prox(x, tau)
        out[f"pos/prox/{t}"] = pxo.PositiveOrthant(dim=N).prox(x, tau)
        out[f"linfball/prox/{t}"] = pxo.LInfinityBall(dim=N, radius=0.8).prox(x, tau)
        out[f"sql2/prox/{t}"] = pxo.SquaredL2Norm(dim=N).prox(x, tau)
        out[f"sql2shift/prox/{t}"] = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-x[0])).prox(x, tau)
        out[f"l21/prox/{t}"] = pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(0,)).prox(x, tau)
        out[f"l21s/fprox/{t}"] = (0.7 * pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(0,))).fenchel_prox(x, tau)
        out[f"l21ax12/prox/{t}"] = pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(1, 2)).prox(x, tau)
    out["l1/apply"] = pxo.L1Norm(dim=N).apply(x)
    out["l21/apply"] = pxo.L21Norm(arg_shape=(3, 4, 5), l2_axis=(0,)).apply(x)
    out["sql2/apply"] = pxo.SquaredL2Norm(dim=N).apply(x)
    out["sql2/grad"] = pxo.SquaredL2Norm(dim=N).grad(x)
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-x[0])
    out["sql2shift/grad"] = f.grad(x)
    out["sql2shift/apply"] = f.apply(x)
    out["sql2shift/diff_lipschitz"] = np.float64(f.diff_lipschitz)
    out["pos/apply"] = pxo.PositiveOrthant(dim=N).apply(np.abs(x))
    np.savez_compressed(os.path.join(HERE, "funcs.npz"), **out)
    print("funcs.npz", len(out))


def _record(slv, prefix, out, keys=("x", "z")):
    data, hist = slv.stats()
    for k in keys:
        if data.get(k) is not None:
            out[f"{prefix}/{k}"] = np.asarray(data[k])
    for k in ("tau", "sigma", "rho"):
        if k in slv._mstate:
            out[f"{prefix}/{k}"] = np.float64(slv._mstate[k])
    out[f"{prefix}/n_hist"] = np.int64(len(hist))
    names = hist.dtype.names
    out[f"{prefix}/hist_last"] = np.array([float(hist[-1][n]) for n in names])
    out[f"{prefix}/hist_names"] = np.array(names)


def solvers(ns):
    stop = ns.stop
    out = {}

    # --- PD3O TV denoising, 2-D, positivity (config[0] in miniature) ---------------------------
    shape = (32, 40)
    _, y = cases.phantom(shape, seed=1)
    out["pd3o_tv2d/y"] = y
    for strat in (1, 2, 3):
        slv = cases.build_tv_denoise(ns, y, shape, lam=0.1)
        slv.fit(x0=y.reshape(-1).copy(), stop_crit=stop.MaxIter(60), tuning_strategy=strat)
        _record(slv, f"pd3o_tv2d/s{strat}", out)

    # default stopping criterion (RelError on x & z): iteration count must match too
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.1)
    slv.fit(x0=y.reshape(-1).copy())
    _record(slv, "pd3o_tv2d/default_stop", out)

    # other boundary modes / no positivity
    for mode in ("reflect", "wrap", "symmetric", "edge"):
        slv = cases.build_tv_denoise(ns, y, shape, lam=0.15, mode=mode, positivity=False)
        slv.fit(x0=np.zeros(y.size), stop_crit=stop.MaxIter(40))
        _record(slv, f"pd3o_tv2d/{mode}", out)

    # --- CondatVu TV denoise (same problem through CV) -------------------------------------------
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.1, solver="CondatVu")
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=stop.MaxIter(60))
    _record(slv, "cv_tv2d", out)

    # --- PD3O TV denoising, 3-D (config[3] in miniature) ---------------------------------------
    shape3 = (10, 12, 14)
    _, y3 = cases.phantom(shape3, seed=2)
    out["pd3o_tv3d/y"] = y3
    slv = cases.build_tv_denoise(ns, y3, shape3, lam=0.08)
    slv.fit(x0=y3.reshape(-1).copy(), stop_crit=stop.MaxIter(50))
    _record(slv, "pd3o_tv3d", out)
    slv = cases.build_tv_denoise(ns, y3, shape3, lam=0.08, mode=("reflect", "wrap", "constant"))
    slv.fit(x0=y3.reshape(-1).copy(), stop_crit=stop.MaxIter(30), tuning_strategy=3)
    _record(slv, "pd3o_tv3d/mixed", out)

    # --- CondatVu TV deblurring with a 9x9 Gaussian Stencil (config[1] in miniature) -----------
    shape = (28, 24)
    x_true, _ = cases.phantom(shape, seed=3)
    g9 = cases.gaussian_1d(9, 1.5)
    for tag, kern in (("dense", np.outer(g9, g9)), ("sep", [g9, g9])):
        slv, A = cases.build_tv_deblur(ns, np.zeros(shape), shape, kern, (4, 4))
        yb = A.apply(x_true.reshape(-1)) + 0.02 * np.random.default_rng(4).standard_normal(x_true.size)
        out[f"cv_deblur2d/{tag}/y"] = yb
        slv, A = cases.build_tv_deblur(ns, yb, shape, kern, (4, 4), lam=0.02)
        slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(40))
        _record(slv, f"cv_deblur2d/{tag}", out)
    # PD3O on the deblurring problem, reflect boundaries on the blur, positivity
    slv, A = cases.build_tv_deblur(ns, yb, shape, np.outer(g9, g9), (4, 4), lam=0.02, blur_mode="reflect",
                                   positivity=True, solver="PD3O")
    slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(40))
    _record(slv, "pd3o_deblur2d", out)

    # --- CondatVu 3-D deblurring, 3x3x3 PSF + positivity (config[4] in miniature; 7^3 in bench) --
    shape3 = (9, 10, 11)
    x3, _ = cases.phantom(shape3, seed=5)
    g3 = cases.gaussian_1d(3, 0.8)
    psf = np.einsum("i,j,k->ijk", g3, g3, g3)
    slv, A = cases.build_tv_deblur(ns, np.zeros(shape3), shape3, psf, (1, 1, 1), positivity=True)
    yb3 = A.apply(x3.reshape(-1))
    out["cv_deblur3d/y"] = yb3
    slv, A = cases.build_tv_deblur(ns, yb3, shape3, psf, (1, 1, 1), lam=0.01, positivity=True)
    slv.fit(x0=np.zeros(yb3.size), stop_crit=stop.MaxIter(30))
    _record(slv, "cv_deblur3d", out)

    # --- PGD / FISTA L1 deconvolution with a 5x5 Stencil over a batch (config[2] in miniature) ---
    B, shape = 3, (20, 22)
    rng = np.random.default_rng(6)
    xs = (rng.random((B,) + shape) > 0.93) * rng.uniform(0.5, 2, (B,) + shape)
    k5 = np.outer(cases.gaussian_1d(5, 1.0), cases.gaussian_1d(5, 1.0))[None]  # (1,5,5): per-image blur
    slv, A = cases.build_l1_deconv(ns, np.zeros(xs.size), (B,) + shape, k5, (0, 2, 2))
    yb = A.apply(xs.reshape(-1)) + 0.01 * rng.standard_normal(xs.size)
    out["pgd_l1/y"] = yb
    for acc in (True, False):
        slv, A = cases.build_l1_deconv(ns, yb, (B,) + shape, k5, (0, 2, 2), lam=0.02)
        slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(50), acceleration=acc, tau=1 / A.lipschitz**2)
        _record(slv, f"pgd_l1/acc{int(acc)}", out, keys=("x",))
    slv, A = cases.build_l1_deconv(ns, yb, (B,) + shape, k5, (0, 2, 2), lam=0.02)
    slv.fit(x0=np.zeros(yb.size), tau=1 / A.lipschitz**2)  # default RelError stop
    _record(slv, "pgd_l1/default_stop", out, keys=("x",))

    np.savez_compressed(os.path.join(HERE, "solvers.npz"), **out)
    print("solvers.npz", len(out))


def slabs(ns):
    """Volumes tall enough to be cut into 2, 3, 4 or 8 z-slabs (tests of the multi-GPU path compare with these, i.e. with
    the reference itself, not with the repo's single-domain solver)."""
    stop = ns.stop
    out = {}
    shape = (32, 12, 16)
    _, y = cases.phantom(shape, seed=21)
    out["y"] = y
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.08)
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=stop.MaxIter(25), rho=1.2)
    _record(slv, "pd3o_tv3d", out)
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.3)
    slv.fit(x0=y.reshape(-1).copy())  # default criterion: the iteration count must match too
    _record(slv, "pd3o_tv3d/default_stop", out)
    for tag, mode in (("ring", ("wrap", "reflect", "edge")), ("fold", ("reflect", "symmetric", "wrap")), ("edge", ("edge", "constant", "symmetric"))):
        slv = cases.build_tv_denoise(ns, y, shape, lam=0.08, mode=mode)
        slv.fit(x0=y.reshape(-1).copy(), stop_crit=stop.MaxIter(20), tuning_strategy=3)
        _record(slv, f"pd3o_tv3d/{tag}", out)
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.08, solver="CondatVu", positivity=False)
    slv.fit(x0=np.zeros(y.size), stop_crit=stop.MaxIter(25))
    _record(slv, "cv_tv3d", out)
    # CondatVu deblurring, separable 7x5x7 PSF (reaches 3 planes across a cut) + positivity; then the same PSF given dense
    x3, _ = cases.phantom(shape, seed=22)
    psf = [cases.gaussian_1d(7, 1.2), cases.gaussian_1d(5, 1.0), cases.gaussian_1d(7, 1.5)]
    slv, A = cases.build_tv_deblur(ns, np.zeros(shape), shape, psf, (3, 2, 3), positivity=True)
    yb = A.apply(x3.reshape(-1)) + 0.01 * np.random.default_rng(23).standard_normal(x3.size)
    out["cv_deblur3d/y"] = yb
    slv, A = cases.build_tv_deblur(ns, yb, shape, psf, (3, 2, 3), lam=0.02, positivity=True)
    slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(15), rho=0.9)
    _record(slv, "cv_deblur3d", out)
    psf6 = [cases.gaussian_1d(6, 1.2), cases.gaussian_1d(5, 1.0), cases.gaussian_1d(7, 1.5)]  # even tap count along z, off-centre
    slv, A = cases.build_tv_deblur(ns, yb, shape, psf6, (2, 2, 3), lam=0.02, positivity=True)
    slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(15))
    _record(slv, "cv_deblur3d/even", out)
    np.savez_compressed(os.path.join(HERE, "slabs.npz"), **out)
    print("slabs.npz", len(out))


def dense_psf(ks, seed, skew=0.6):
    """A dense PSF of FULL RANK (not an outer product): a sheared anisotropic Gaussian plus a small irregular part, normalised to sum 1."""
    ax = [np.arange(k) - (k - 1) / 2 for k in ks]
    a, b, c = np.meshgrid(*ax, indexing="ij")
    q = (a / (0.35 * ks[0])) ** 2 + ((b - skew * a) / (0.3 * ks[1])) ** 2 + ((c + skew * b) / (0.4 * ks[2])) ** 2
    k = np.exp(-0.5 * q) * (1 + 0.1 * np.random.default_rng(seed).standard_normal(ks))
    return k / k.sum()


def dense3d(ns):
    """Dense 3-D kernels of full rank (the reference takes any dense kernel: stencil.py:356-461): Stencil apply / adjoint, and CondatVu
    TV deblurring with such a PSF on a volume tall enough to be cut into z-slabs."""
    stop = ns.stop
    out = {}
    rng = np.random.default_rng(31)
    for i, (shape, ks, cen) in enumerate((((12, 19, 24), (7, 7, 7), (3, 3, 3)), ((11, 18, 16), (5, 4, 5), (4, 0, 3)), ((6, 20, 24), (3, 3, 3), (1, 1, 1)),
                                          ((9, 21, 28), (7, 7, 7), (0, 6, 1)))):
        kern = rng.standard_normal(ks)
        op = ns.operator.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        x = rng.standard_normal((2, int(np.prod(shape))))
        out[f"st{i}/shape"], out[f"st{i}/center"], out[f"st{i}/kernel"], out[f"st{i}/x"] = np.array(shape), np.array(cen), kern, x
        out[f"st{i}/apply"], out[f"st{i}/adjoint"] = np.asarray(op.apply(x)), np.asarray(op.adjoint(x))
    shape = (32, 12, 16)
    x3, _ = cases.phantom(shape, seed=32)
    psf = dense_psf((5, 5, 5), 33)
    assert np.linalg.svd(psf.reshape(5, 25), compute_uv=False)[1] > 1e-3  # full rank: no separable pass applies
    slv, A = cases.build_tv_deblur(ns, np.zeros(shape), shape, psf, (2, 2, 2), positivity=True)
    yb = np.asarray(A.apply(x3.reshape(-1))) + 0.01 * np.random.default_rng(34).standard_normal(x3.size)
    out["cv_deblur3d_dense/y"], out["cv_deblur3d_dense/psf"] = yb, psf
    slv, A = cases.build_tv_deblur(ns, yb, shape, psf, (2, 2, 2), lam=0.02, positivity=True)
    slv.fit(x0=np.zeros(yb.size), stop_crit=stop.MaxIter(15), rho=0.9)
    _record(slv, "cv_deblur3d_dense", out)
    np.savez_compressed(os.path.join(HERE, "dense3d.npz"), **out)
    print("dense3d.npz", len(out))


def config0(ns):
    """BASELINE.json configs[0] at full size: 512x512 float64 PD3O TV denoising, 200 iterations.

    Only a strided subsample + norms are stored (fixture stays small)."""
    shape = (512, 512)
    _, y = cases.phantom(shape, seed=11, noise=0.15)
    slv = cases.build_tv_denoise(ns, y, shape, lam=0.1)
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=ns.stop.MaxIter(200))
    d, hist = slv.stats()
    x, z = np.asarray(d["x"]), np.asarray(d["z"])
    out = dict(
        x_sub=x[::37], z_sub=z[::41], x_norm=np.linalg.norm(x), z_norm=np.linalg.norm(z), x_sum=x.sum(),
        tau=np.float64(slv._mstate["tau"]), sigma=np.float64(slv._mstate["sigma"]), rho=np.float64(slv._mstate["rho"]),
    )
    np.savez_compressed(os.path.join(HERE, "config0.npz"), **out)
    print("config0.npz")


if __name__ == "__main__":
    ns = _ref_import.load()
    which = sys.argv[1:] or ["stencils", "gradients", "funcs", "solvers", "slabs", "dense3d", "config0"]
    for w in which:
        globals()[w](ns)
