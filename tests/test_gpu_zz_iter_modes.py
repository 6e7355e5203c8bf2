"""
GPU parity of the single-kernel iteration (pxb_pds_iter) with FOLDING boundary modes (numpy.pad wrap / reflect /
symmetric / edge; reference pad.py:252-302): the MODES instances of the three forms (direct loads, TMA pipeline, 2-D TMA
tiles), switched on explicitly with pxb_set_iter_modes(1),

* against the two-sweep kernels (pxb_pds_primal + pxb_pds_dual) on random states, every scheme / mode / dtype;
* against the fixtures of the real reference through Solver.fit() (the plan must then report the single-kernel form).

(The file sorts last on purpose: these instances are the newest kernels of the library.)
"""
import numpy as np
import pytest

import cases
from conftest import golden
import test_gpu_iter as GI
from test_gpu_iter import assert_same, both_forms, env, params  # noqa: F401  (env is a fixture)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

MODES = ["wrap", "reflect", "symmetric", "edge"]


@pytest.fixture
def modes_on(env):
    env.K.check(env.lib.pxb_set_iter_modes(1), "pxb_set_iter_modes")
    yield lambda p: env.K.check(env.lib.pxb_set_iter_path(p), "pxb_set_iter_path")
    env.lib.pxb_set_iter_modes(-1)
    env.lib.pxb_set_iter_path(0)


@pytest.mark.parametrize("mode", MODES + [("constant", "reflect", "wrap"), ("edge", "constant", "symmetric")])
@pytest.mark.parametrize("form", ["direct", "tma"])
def test_modes_vs_two_sweeps_3d(env, mode, form, modes_on):
    K = env.K
    modes_on(1 if form == "direct" else 2)
    for dtype in (torch.float64, torch.float32):
        vec = 4 if dtype == torch.float32 else 2
        shape = (21, 19, 32 * vec * 2 + 3 * vec)
        tol = 1e-13 if dtype == torch.float64 else 2e-6
        for scheme in ("forward", "backward", "central"):
            Kop = env.operator.Gradient(arg_shape=shape, scheme=scheme, mode=mode, sampling=(1.0, 0.5, 2.0),
                                        dtype=np.float32 if dtype == torch.float32 else np.float64)
            shift = torch.randn(Kop.dim, device=GI.DEV, dtype=dtype)
            for algo, hkind, gspec in ((K.ALGO_PD3O, K.DUAL_L21, (K.PROX_POS, 0.0, 0.0)), (K.ALGO_CV, K.DUAL_L1, (K.PROX_BOX, 0.2, 0.9))):
                P = params(K, 0.21, 0.19, 0.9, gspec, K.F_SQL2, 0.7, shift, None, hkind, 0.3)
                for chunk in (0, 5):
                    a, b = both_forms(env, algo, Kop, 1, dtype, P, seed=chunk, chunk=chunk)
                    assert_same(env, algo, a, b, tol)


@pytest.mark.parametrize("mode", MODES + [("reflect", "wrap"), ("constant", "edge")])
@pytest.mark.parametrize("form", ["direct", "tile2d"])
def test_modes_vs_two_sweeps_2d_batched(env, mode, form, modes_on):
    K = env.K
    modes_on(1 if form == "direct" else 2)
    for dtype, shape, batch in ((torch.float32, (37, 300), 3), (torch.float64, (45, 1100), 2), (torch.float64, (3, 4), 2)):
        tol = 1e-13 if dtype == torch.float64 else 2e-6
        for scheme in ("forward", "backward", "central"):
            Kop = env.operator.Gradient(arg_shape=shape, scheme=scheme, mode=mode, dtype=np.float32 if dtype == torch.float32 else np.float64)
            shift = torch.randn(batch, Kop.dim, device=GI.DEV, dtype=dtype)
            garr = torch.randn(batch, Kop.dim, device=GI.DEV, dtype=dtype)
            P = params(K, 0.3, 0.25, 0.95, (K.PROX_L1, 0.05, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.2)
            for algo in (K.ALGO_PD3O, K.ALGO_CV):
                a, b = both_forms(env, algo, Kop, batch, dtype, P, seed=3)
                assert_same(env, algo, a, b, tol)
            P = params(K, 0.3, 0.25, 0.95, (K.PROX_POS, 0.0, 0.0), K.F_GRADARR, 0.0, None, garr, K.DUAL_L1, 0.2)
            a, b = both_forms(env, K.ALGO_CV, Kop, batch, dtype, P, seed=4)
            assert_same(env, K.ALGO_CV, a, b, tol)


@pytest.mark.parametrize("mode", MODES + [("wrap", "reflect", "symmetric")])
def test_modes_full_tiles_3d_and_2d(env, mode, modes_on):
    """Full tiles: the folded rim rows / columns of the edge tiles are served from the tile's own staged boxes (pxb_rim_src), and
    a larger volume (256 x 64 x 256) so that many CTAs, chunks and bands take part."""
    K = env.K
    modes_on(2)
    for dtype, shape in ((torch.float32, (16, 32, 256)), (torch.float64, (9, 16, 128)), (torch.float32, (256, 64, 256))):
        Kop = env.operator.Gradient(arg_shape=shape, mode=mode, dtype=np.float32 if dtype == torch.float32 else np.float64)
        shift = torch.randn(Kop.dim, device=GI.DEV, dtype=dtype)
        P = params(K, 0.21, 0.19, 0.9, (K.PROX_POS, 0.0, 0.0), K.F_SQL2, 0.7, shift, None, K.DUAL_L21, 0.3)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            a, b = both_forms(env, algo, Kop, 1, dtype, P, seed=5)
            assert_same(env, algo, a, b, 1e-13 if dtype == torch.float64 else 2e-6)
    mode2 = mode if isinstance(mode, str) else mode[1:]
    for dtype, shape, batch in ((torch.float32, (64, 512), 2), (torch.float64, (32, 128), 1)):
        Kop = env.operator.Gradient(arg_shape=shape, mode=mode2, dtype=np.float32 if dtype == torch.float32 else np.float64)
        shift = torch.randn(batch, Kop.dim, device=GI.DEV, dtype=dtype)
        P = params(K, 0.3, 0.25, 0.95, (K.PROX_L1, 0.05, 0.0), K.F_SQL2, 0.5, shift, None, K.DUAL_L21, 0.2)
        for algo in (K.ALGO_PD3O, K.ALGO_CV):
            a, b = both_forms(env, algo, Kop, batch, dtype, P, seed=6)
            assert_same(env, algo, a, b, 1e-13 if dtype == torch.float64 else 2e-6)


def test_modes_solver_fit_against_reference_fixtures(env, modes_on):
    """Solver.fit() with folding modes runs the single-kernel form and reproduces the real reference (<= 1e-10)."""
    import types

    px = types.SimpleNamespace(operator=env.operator, solver=env.solver, stop=env.stop)
    g = golden("solvers.npz")
    y = g["pd3o_tv2d/y"]

    def rel(a, b):
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        return np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel())

    for mode in MODES:
        slv = cases.build_tv_denoise(px, y, (32, 40), lam=0.15, mode=mode, positivity=False)
        slv.fit(x0=np.zeros(y.size), stop_crit=px.stop.MaxIter(40))
        assert slv._astate.get("error") is None, slv._astate.get("error")
        assert slv._plan.kind == "fused" and slv._plan.iter_ok is True
        data, _ = slv.stats()
        assert rel(data["x"], g[f"pd3o_tv2d/{mode}/x"]) < 1e-10 and rel(data["z"], g[f"pd3o_tv2d/{mode}/z"]) < 1e-10
    y3 = g["pd3o_tv3d/y"]
    slv = cases.build_tv_denoise(px, y3, (10, 12, 14), lam=0.08, mode=("reflect", "wrap", "constant"))
    slv.fit(x0=y3.reshape(-1).copy(), stop_crit=px.stop.MaxIter(30), tuning_strategy=3)
    assert slv._plan.iter_ok is True
    data, _ = slv.stats()
    assert rel(data["x"], g["pd3o_tv3d/mixed/x"]) < 1e-10 and rel(data["z"], g["pd3o_tv3d/mixed/z"]) < 1e-10
