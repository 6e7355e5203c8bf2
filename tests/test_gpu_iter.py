"""
GPU parity of the single-kernel PD3O / CondatVu iteration (pxb_pds_iter) called through the C ABI:

* against the two-sweep kernels (pxb_pds_primal + pxb_pds_dual) on random states -- every finite-difference scheme,
  2-D / 3-D, ragged tiles, several chunks, batches, fp32 / fp64;
* against the NumPy oracle (oracle/pyxu_oracle.py, pinned on the real reference) after N iterations through
  Solver.fit(), including the lazily materialised x and the iteration count under the default RelError criterion;
* at a large size through size-independent properties (two-sweep agreement on a 256^3 volume, chunk invariance).
"""
import ctypes as C

import numpy as np
import pytest

import cases
from conftest import golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

# Where the test arrays live.  "cuda" on a GPU box; tests/test_emu_device_solvers.py replays some of these functions on the
# emulated device (tests/emu_device.py) with DEV = "cpu".
DEV = "cuda"


def _sync():
    if DEV == "cuda":
        torch.cuda.synchronize()


@pytest.fixture(scope="module")
def env():
    import types

    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst
    from pyxu_b200 import _array as A
    from pyxu_b200 import _cabi as K

    assert torch.cuda.is_available()
    return types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst, A=A, K=K, lib=K.lib())


@pytest.fixture
def iter_path(env):
    """select the implementation behind pxb_pds_iter for one test (1 direct loads, 2 TMA), restore 'auto' after"""
    yield lambda p: env.K.check(env.lib.pxb_set_iter_path(p), "pxb_set_iter_path")
    env.lib.pxb_set_iter_path(0)


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-300))


def params(K, tau, sigma, rho, gspec, fkind, alpha, shift, garr, hkind, lam):
    P = K.PdsParams()
    P.tau, P.sigma, P.rho = tau, sigma, rho
    P.g = K.ProxSpec(gspec[0], 0, gspec[1], gspec[2])
    f = K.FTerm()
    f.kind, f.alpha = fkind, alpha
    if shift is not None:
        f.shift, f.shift_period = shift.data_ptr(), shift.numel()
    if garr is not None:
        f.garr = garr.data_ptr()
    P.f = f
    P.hkind, P.lam = hkind, lam
    return P


def both_forms(env, algo, Kop, batch, dtype, P, seed=0, chunk=0, norms=True):
    """Runs one iteration in both forms from the same random state; returns ((u, z, x, nx, nz) two-sweep, same single-kernel)."""
    K, lib, A = env.K, env.lib, env.A
    gen = torch.Generator(device=DEV).manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, device=DEV, dtype=dtype, generator=gen)
    d = Kop._desc(batch, K.F32 if dtype == torch.float32 else K.F64)
    u, x, z = rnd(batch, Kop.dim), rnd(batch, Kop.dim), rnd(batch, Kop.codim)
    st = A.stream()
    # two sweeps (in place)
    ua, xa, za, w = u.clone(), x.clone(), z.clone(), torch.empty_like(u)
    na = torch.zeros((2, batch, 2), dtype=torch.float64, device=DEV)
    pd3o = algo == K.ALGO_PD3O
    K.check(lib.pxb_pds_primal(algo, C.byref(d), C.byref(P), ua.data_ptr(), za.data_ptr(), None, xa.data_ptr() if pd3o else None, w.data_ptr(),
                               na[0].data_ptr() if norms else None, st), "primal")
    K.check(lib.pxb_pds_dual(C.byref(d), C.byref(P), w.data_ptr(), za.data_ptr(), na[1].data_ptr() if norms else None, st), "dual")
    # one sweep (out of place)
    ub, zb, xb = torch.full_like(u, float("nan")), torch.full_like(z, float("nan")), x.clone()
    nb = torch.zeros((2, batch, 2), dtype=torch.float64, device=DEV)
    args = (algo, C.byref(d), C.byref(P), u.data_ptr(), z.data_ptr(), ub.data_ptr(), zb.data_ptr(), xb.data_ptr() if pd3o else None,
            nb[0].data_ptr() if norms else None, nb[1].data_ptr() if norms else None)
    rc = lib.pxb_pds_iter_chunked(*args, chunk, st) if chunk else lib.pxb_pds_iter(*args, st)
    K.check(rc, "pxb_pds_iter")
    _sync()
    return (ua, za, xa, na), (ub, zb, xb, nb)


def assert_same(env, algo, a, b, tol):
    (ua, za, xa, na), (ub, zb, xb, nb) = a, b
    assert torch.isfinite(ub).all() and torch.isfinite(zb).all()
    assert rel(ub, ua) < tol and rel(zb, za) < tol, (rel(ub, ua), rel(zb, za))
    if algo == env.K.ALGO_PD3O:
        assert rel(xb, xa) < tol
    assert torch.allclose(na, nb, rtol=1e-9 if tol < 1e-9 else 1e-5), (na, nb)


@pytest.mark.parametrize("scheme", ["forward", "backward", "central"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_vs_two_sweeps_3d(env, scheme, dtype, form, iter_path):
    K = env.K
    iter_path(1 if form == "direct" else 2)
    vec = 4 if dtype == torch.float32 else 2
    shape = (21, 19, 32 * vec * 2 + 3 * vec)  # 3 row tiles (ragged), 3 column tiles (ragged)
    Kop = env.operator.Gradient(arg_shape=shape, scheme=scheme, sampling=(1.0, 0.5, 2.0), dtype=np.float32 if dtype == torch.float32 else np.float64)
    shift = torch.randn(Kop.dim, device=DEV, dtype=dtype)
    tol = 1e-13 if dtype == torch.float64 else 2e-6
    for algo in (K.ALGO_PD3O, K.ALGO_CV):
        for hkind, gspec in ((K.DUAL_L21, (K.PROX_POS, 0.0, 0.0)), (K.DUAL_L1, (K.PROX_BOX, 0.2, 0.9))):
            P = params(K, 0.21, 0.19, 0.9, gspec, K.F_SQL2, 0.7, shift, None, hkind, 0.3)
            for chunk in (0, 5, 1):
                a, b = both_forms(env, algo, Kop, 1, dtype, P, seed=chunk, chunk=chunk)
                assert_same(env, algo, a, b, tol)


@pytest.mark.parametrize("scheme", ["forward", "backward", "central"])
@pytest.mark.parametrize("width", [40, 300, 1100])
@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_vs_two_sweeps_2d_batched(env, scheme, width, form, iter_path):
    K = env.K
    iter_path(1 if form == "direct" else 2)  # 2: TMA-staged 2-D tiles (pxb_tv_tile2d.cu)
    Kop = env.operator.Gradient(arg_shape=(37, width), scheme=scheme, dtype=np.float32)
    shift = torch.randn(Kop.dim, device=DEV, dtype=torch.float32)  # one image: broadcast over the batch
    P = params(K, 0.3, 0.25, 1.0, (K.PROX_L1, 0.05, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.2)
    for algo in (K.ALGO_PD3O, K.ALGO_CV):
        for chunk in ((0, 4) if form == "direct" else (0,)):
            a, b = both_forms(env, algo, Kop, 5, torch.float32, P, seed=width + chunk, chunk=chunk)
            assert_same(env, algo, a, b, 2e-6)


@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_cv_gradarr_stacked(env, form, iter_path):
    K = env.K
    iter_path(1 if form == "direct" else 2)
    Kop = env.operator.Gradient(arg_shape=(3, 33, 24), directions=(1, 2))
    garr = torch.randn(2, Kop.dim, device=DEV, dtype=torch.float64)
    P = params(K, 0.3, 0.25, 0.8, (K.PROX_POS, 0.0, 0.0), K.F_GRADARR, 0.0, None, garr, K.DUAL_L21, 0.2)
    a, b = both_forms(env, K.ALGO_CV, Kop, 2, torch.float64, P, chunk=7 if form == "direct" else 0)
    assert_same(env, K.ALGO_CV, a, b, 1e-13)
    # a volume-shaped shift broadcast over the batch (PD3O), fp64
    shift = torch.randn(Kop.dim, device=DEV, dtype=torch.float64)
    P = params(K, 0.3, 0.25, 0.8, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L1, 0.2)
    a, b = both_forms(env, K.ALGO_PD3O, Kop, 2, torch.float64, P)
    assert_same(env, K.ALGO_PD3O, a, b, 1e-13)


def test_iter_envelope_errors(env):
    K, lib = env.K, env.lib
    Kop = env.operator.Gradient(arg_shape=(8, 16), mode="reflect")
    d = Kop._desc(1, K.F64)
    P = params(K, 0.3, 0.25, 1.0, (K.PROX_NONE, 0.0, 0.0), K.F_NONE, 0.0, None, None, K.DUAL_L21, 0.2)
    u, z = torch.zeros(Kop.dim, device=DEV, dtype=torch.float64), torch.zeros(Kop.codim, device=DEV, dtype=torch.float64)
    u2, z2 = u.clone(), z.clone()
    n0 = lib.pxb_launch_count()
    K.check(lib.pxb_set_iter_modes(0), "pxb_set_iter_modes")  # folding modes declined: callers take the two-sweep form
    try:
        assert lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u.data_ptr(), z.data_ptr(), u2.data_ptr(), z2.data_ptr(), None, None, None, None) == -3
        assert b"reason 4" in lib.pxb_last_error()
    finally:
        lib.pxb_set_iter_modes(-1)
    assert lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u.data_ptr(), z.data_ptr(), u.data_ptr(), z2.data_ptr(), None, None, None, None) == -1
    assert lib.pxb_launch_count() == n0  # nothing was launched


@pytest.mark.parametrize("scheme", ["forward", "central"])
def test_tma_form_batched_shift_modes(env, scheme, iter_path):
    """TMA form with batch > 1: shift broadcast over the batch (one volume), per-item shift, scalar shift; CV + grad f array."""
    K = env.K
    iter_path(2)
    Kop = env.operator.Gradient(arg_shape=(9, 21, 136), scheme=scheme, dtype=np.float32)
    gen = torch.Generator(device=DEV).manual_seed(3)
    for shape in ((Kop.dim,), (3, Kop.dim), (1,)):
        shift = torch.randn(*shape, device=DEV, dtype=torch.float32, generator=gen)
        P = params(K, 0.21, 0.19, 0.9, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.7, shift, None, K.DUAL_L21, 0.3)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            a, b = both_forms(env, algo, Kop, 3, torch.float32, P, seed=len(shape), chunk=4)
            assert_same(env, algo, a, b, 2e-6)
    garr = torch.randn(3, Kop.dim, device=DEV, dtype=torch.float32, generator=gen)
    P = params(K, 0.3, 0.25, 0.8, (K.PROX_L1, 0.1, 0.0), K.F_GRADARR, 0.0, None, garr, K.DUAL_L1, 0.2)
    a, b = both_forms(env, K.ALGO_CV, Kop, 3, torch.float32, P)
    assert_same(env, K.ALGO_CV, a, b, 2e-6)


@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_large_volume_properties(env, form, iter_path):
    """256^3 fp32 (every SM busy, many chunks): agreement with the two-sweep form and chunk invariance."""
    K = env.K
    iter_path(1 if form == "direct" else 2)
    Kop = env.operator.Gradient(arg_shape=(256, 256, 256), dtype=np.float32)
    shift = torch.randn(Kop.dim, device=DEV, dtype=torch.float32)
    P = params(K, 0.28, 0.28, 1.0, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.08)
    a, b = both_forms(env, K.ALGO_PD3O, Kop, 1, torch.float32, P, seed=1)
    assert_same(env, K.ALGO_PD3O, a, b, 2e-6)
    _, c = both_forms(env, K.ALGO_PD3O, Kop, 1, torch.float32, P, seed=1, chunk=37)
    assert torch.equal(b[0], c[0]) and torch.equal(b[1], c[1]) and torch.equal(b[2], c[2])  # bitwise: same arithmetic per voxel


# ---- through Solver.fit(): the oracle / the real reference's fixtures ------------------------------------
def relnp(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    return np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel())


def test_solver_uses_single_kernel_and_lazy_x(env):
    g = golden("solvers.npz")
    y = g["pd3o_tv3d/y"]
    slv = cases.build_tv_denoise(env, y, (10, 12, 14), lam=0.08, final_writeback=False)  # the final dump would materialise x
    n0 = env.lib.pxb_launch_count()
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=env.stop.MaxIter(50))
    assert slv._plan.kind == "fused" and slv._plan.iter_ok is True
    assert slv._x_stale  # MaxIter never reads x: no iteration wrote it
    launches = env.lib.pxb_launch_count() - n0
    data, _ = slv.stats()  # materialises x from the previous iterate
    assert not slv._x_stale
    assert relnp(data["x"], g["pd3o_tv3d/x"]) < 1e-10 and relnp(data["z"], g["pd3o_tv3d/z"]) < 1e-10
    assert launches <= 50 + 8, launches  # one kernel per iteration (+ set-up: K x0, clones)


def test_solver_default_stop_fused_norms_single_kernel(env):
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    slv = cases.build_tv_denoise(env, y, (32, 40), lam=0.1)
    slv.fit(x0=y.reshape(-1).copy())
    assert slv._plan.iter_ok is True and "_fused_norms" in slv._mstate and not slv._x_stale
    data, hist = slv.stats()
    assert len(hist) == int(g["pd3o_tv2d/default_stop/n_hist"])
    assert relnp(data["x"], g["pd3o_tv2d/default_stop/x"]) < 1e-9
    last = np.array([float(hist[-1][n]) for n in hist.dtype.names])
    assert np.allclose(last, g["pd3o_tv2d/default_stop/hist_last"], rtol=1e-6)


def test_solver_cv_single_kernel(env):
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]
    slv = cases.build_tv_denoise(env, y, (32, 40), lam=0.1, solver="CondatVu")
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=env.stop.MaxIter(60))
    assert slv._plan.iter_ok is True
    data, _ = slv.stats()
    assert relnp(data["x"], g["cv_tv2d/x"]) < 1e-10 and relnp(data["z"], g["cv_tv2d/z"]) < 1e-10


def test_solver_fp32_tolerance_256cube_vs_oracle_property(env):
    """fp32 run of a 96^3 problem against the fp64 run of the same solver (<= 1e-4, the north-star tolerance)."""
    rng = np.random.default_rng(0)
    y = rng.random((96, 96, 96))
    outs = []
    for dt in (np.float64, np.float32):
        slv = cases.build_tv_denoise(env, y.astype(dt), (96, 96, 96), lam=0.08, dtype=dt)
        slv.fit(x0=y.reshape(-1).astype(dt), stop_crit=env.stop.MaxIter(40))
        assert slv._plan.iter_ok is True
        outs.append(slv.solution())
    assert relnp(outs[1], np.asarray(outs[0], dtype=np.float64)) < 1e-4


def test_large_results_leave_the_device_through_the_pipelined_copy(env):
    """HOST-origin results above 64 MiB: chunked D2H through pinned staging buffers == plain .cpu()."""
    gen = torch.Generator(device=DEV).manual_seed(5)
    for dt, n in ((torch.float32, (70 << 20) // 4 + 12345), (torch.float64, (97 << 20) // 8 + 1)):
        t = torch.randn(n, device=DEV, dtype=dt, generator=gen)
        out = env.A.restore(t, env.A.HOST)
        assert isinstance(out, np.ndarray) and out.shape == (n,) and np.array_equal(out, t.cpu().numpy())
    t2 = torch.randn(3, (80 << 20) // 12, device=DEV, generator=gen)
    assert np.array_equal(env.A.restore(t2, env.A.HOST), t2.cpu().numpy())
    # and the way in: a pageable NumPy array above 64 MiB goes through the pinned staging buffers
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        h = rng.standard_normal((3, (70 << 20) // 12 + 7)).astype(dt)
        d, origin = env.A.asdevice(h)
        assert origin == env.A.HOST and d.is_cuda and tuple(d.shape) == h.shape and np.array_equal(d.cpu().numpy(), h)
    d32, _ = env.A.asdevice(h, dtype=torch.float32)  # with a dtype conversion on the device
    assert d32.dtype == torch.float32 and np.array_equal(d32.cpu().numpy(), h.astype(np.float32))


@pytest.mark.parametrize("form", ["direct", "tma"])
def test_iter_random_small_shapes(env, form, iter_path):
    """Ragged / degenerate geometries: single planes, fewer rows than a tile, one vector per row, 1-3 chunks, batches --
    every scheme, PD3O and CondatVu, fp32 and fp64, against the two-sweep kernels."""
    K = env.K
    iter_path(1 if form == "direct" else 2)
    rng = np.random.default_rng(123)
    schemes = ["forward", "backward", "central"]
    for trial in range(36):
        dtype = torch.float64 if trial % 2 else torch.float32
        vec = 2 if dtype == torch.float64 else 4
        ndim = 3 if trial % 3 else 2
        n2 = vec * int(rng.integers(1, 40))
        shape = (int(rng.integers(1, 12)), int(rng.integers(1, 20)), n2) if ndim == 3 else (int(rng.integers(1, 40)), n2)
        batch = int(rng.integers(1, 4))
        Kop = env.operator.Gradient(arg_shape=shape, scheme=schemes[trial % 3], dtype=np.float64 if dtype == torch.float64 else np.float32,
                                    sampling=float(rng.uniform(0.5, 2.0)))
        shift = torch.randn(Kop.dim if trial % 4 else batch * Kop.dim, device=DEV, dtype=dtype)
        gspec = [(K.PROX_POS, 0.0, 0.0), (K.PROX_NONE, 0.0, 0.0), (K.PROX_L1, 0.05, 0.0)][trial % 3]
        P = params(K, 0.21, 0.19, float(rng.uniform(0.7, 1.3)), gspec, K.F_SQL2, 0.5, shift, None, K.DUAL_L21 if trial % 5 else K.DUAL_L1, 0.3)
        algo = K.ALGO_PD3O if trial % 2 == 0 else K.ALGO_CV
        a, b = both_forms(env, algo, Kop, batch, dtype, P, seed=trial)
        assert_same(env, algo, a, b, 1e-12 if dtype == torch.float64 else 3e-6)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("algo", ["pd3o", "cv"])
def test_small_2d_batch_is_one_cooperative_launch(env, dtype, algo):
    """pxb_pds_iter_n on an image whose tiles are all resident at once: ONE launch for the n iterations (k_tv_tile2d_loop, grid-wide
    barrier between iterations) -- the iterate, the per-iteration RelError sums and the device-side stop are those of n single launches;
    an image too large for that, and a folding boundary mode, keep one launch per iteration."""
    K, lib = env.K, env.lib
    tdt, ndt, kdt = (torch.float64, np.float64, K.F64) if dtype == "f64" else (torch.float32, np.float32, K.F32)
    A_ = K.ALGO_PD3O if algo == "pd3o" else K.ALGO_CV
    n = 37

    def problem(shape, mode="constant"):
        N = int(np.prod(shape))
        g = torch.Generator(device="cpu").manual_seed(3)
        y = torch.rand(N, generator=g, dtype=tdt).to(DEV)
        shift = -y
        P = params(K, 0.28, 0.28, 1.2, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.1)
        d = env.operator.Gradient(arg_shape=shape, dtype=ndt, mode=mode)._desc(1, kdt)
        zi = (0.05 * torch.randn(2 * N, generator=g, dtype=tdt)).to(DEV)  # (z0 != 0: with z0 = 0 the first x equals x0 and RelError[x] is met at once)
        st = lambda: (y.clone(), torch.empty_like(y), zi.clone(), torch.empty(2 * N, device=DEV, dtype=tdt), y.clone())
        return N, P, d, st, (y, shift)

    N, P, d, st, keep = problem((200, 264))
    # (a) no rule: n plain iterations
    u0, u1, z0, z1, x = st()
    c0 = lib.pxb_launch_count()
    K.check(lib.pxb_pds_iter_n(A_, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, n, None, None, None), "iter_n")
    assert lib.pxb_launch_count() - c0 == 1, "the batch must be one launch"
    ra_u, ra_z = (u1, z1) if n % 2 else (u0, z0)
    v0, v1, w0, w1, _ = st()
    a, b = (v0, w0), (v1, w1)
    for _ in range(n):
        K.check(lib.pxb_pds_iter(A_, C.byref(d), C.byref(P), a[0].data_ptr(), a[1].data_ptr(), b[0].data_ptr(), b[1].data_ptr(), None, None, None, None), "iter")
        a, b = b, a
    _sync()
    assert torch.equal(ra_u, a[0]) and torch.equal(ra_z, a[1])
    # (b) with the device-side rule: sums of every iteration, stop at the iteration where RelError[z] <= eps
    rule = K.StopRule()
    rule.eps_x, rule.eps_z, rule.all_x, rule.all_z = (1e-30 if algo == "pd3o" else 0.0), 0.05, 1, 1
    rule.table = 0b1110  # bit (2*px + pz): stop when either test holds
    u0, u1, z0, z1, x = st()
    sums = torch.zeros(n * 4, device=DEV, dtype=torch.float64)
    ctl = torch.zeros(4, device=DEV, dtype=torch.int32)
    c0 = lib.pxb_launch_count()
    K.check(lib.pxb_pds_iter_n(A_, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), x.data_ptr() if algo == "pd3o" else None,
                               sums.data_ptr(), n, C.byref(rule), ctl.data_ptr(), None), "iter_n")
    assert lib.pxb_launch_count() - c0 == 1
    _sync()
    done = int(ctl[1])
    assert 2 <= done < n and int(ctl[0]) == 1, (done, ctl)
    S = sums.cpu().numpy().reshape(n, 2, 1, 2)
    assert np.sqrt(S[done - 1, 1, 0, 0]) <= 0.05 * np.sqrt(S[done - 1, 1, 0, 1]) and np.all(np.sqrt(S[: done - 1, 1, 0, 0]) > 0.05 * np.sqrt(S[: done - 1, 1, 0, 1]))
    assert not S[done:].any()
    # the same iterations one launch each, sums per launch
    v0, v1, w0, w1, xx = st()
    a, b = (v0, w0), (v1, w1)
    for i in range(done):
        nx = torch.zeros(2, device=DEV, dtype=torch.float64)
        nz = torch.zeros(2, device=DEV, dtype=torch.float64)
        K.check(lib.pxb_pds_iter(A_, C.byref(d), C.byref(P), a[0].data_ptr(), a[1].data_ptr(), b[0].data_ptr(), b[1].data_ptr(), xx.data_ptr() if algo == "pd3o" else None,
                                 nx.data_ptr() if rule.eps_x > 0 else None, nz.data_ptr(), None), "iter")
        a, b = b, a
        assert np.allclose(nz.cpu().numpy(), S[i, 1, 0], rtol=1e-9 if dtype == "f64" else 1e-5)
    got_u, got_z = (u1, z1) if done % 2 else (u0, z0)
    assert torch.equal(got_u, a[0]) and torch.equal(got_z, a[1])
    if algo == "pd3o":
        assert torch.equal(x, xx)
    # (c) too many tiles to be resident at once / a folding mode: one launch per iteration
    for shape, mode in (((1600, 2048) if dtype == "f64" else (2048, 4096), "constant"), ((200, 264), "reflect")):
        N2, P2, d2, st2, keep2 = problem(shape, mode)
        u0, u1, z0, z1, x = st2()
        c0 = lib.pxb_launch_count()
        K.check(lib.pxb_pds_iter_n(A_, C.byref(d2), C.byref(P2), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, 4, None, None, None), "iter_n")
        assert lib.pxb_launch_count() - c0 == 4, (shape, mode)


def test_criterion_iteration_is_bitwise_reproducible_at_full_size(env):
    """The headline volume (1024^3 fp32, 131 072 thread blocks): K iterations of the criterion-carrying TMA instance twice from the same
    state must agree bit for bit.  Guards the one asynchrony bug of round 2 -- a value loaded from a staged box right before the
    per-plane barrier and first used in the next plane could be read after the TMA unit had refilled the box (one wrong rim cell in ~10^6
    thread-block-planes, different from run to run).  Before the fix every second pair of runs differed."""
    K, lib = env.K, env.lib
    free, _ = torch.cuda.mem_get_info()
    if free < 70 << 30:
        pytest.skip("needs ~60 GiB of free device memory")
    shape = (1024, 1024, 1024)
    N = int(np.prod(shape))
    y = torch.rand(N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    shift = -y
    P = params(K, 0.28, 0.28, 1.0, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.08)
    d = env.operator.Gradient(arg_shape=shape, dtype=np.float32)._desc(1, K.F32)
    zinit = 0.01 * torch.randn(3 * N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(7))
    ref = None
    for rep in range(4):
        u0, u1, z0, z1, x = y.clone(), torch.zeros_like(y), zinit.clone(), torch.zeros(3 * N, device=DEV), y.clone()
        nx, nz = torch.zeros(2, device=DEV, dtype=torch.float64), torch.zeros(2, device=DEV, dtype=torch.float64)
        a, b = (u0, z0), (u1, z1)
        for _ in range(6):
            K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), a[0].data_ptr(), a[1].data_ptr(), b[0].data_ptr(), b[1].data_ptr(),
                                     x.data_ptr(), nx.data_ptr(), nz.data_ptr(), None), "iter")
            a, b = b, a
        _sync()
        if ref is None:
            ref = (a[0].clone(), a[1].clone())
        else:
            assert torch.equal(ref[0], a[0]) and torch.equal(ref[1], a[1]), f"run {rep} differs from run 0"
        del u0, u1, z0, z1, x
