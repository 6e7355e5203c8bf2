"""The GPU parity tests of the solvers (tests/test_gpu_solvers.py: every fixture the real reference produced, re-run through
Solver.fit()) replayed on the build container: same test functions, same tolerances, with the device emulated
(tests/emu_device.py: CPU tensors + the kernel bodies of tests/emu behind the C-ABI entry points).  This is how the host logic
around the kernels that have not run on a GPU yet -- the folding-mode instances of the single-kernel iteration after their last
change, Stencil._run_padded -- is checked end to end against the reference's fixtures."""
import types

import numpy as np
import pytest

import cases
import test_gpu_solvers as G
from conftest import golden
from emu_device import emulated_device


@pytest.fixture
def dev():
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    with emulated_device() as lib:
        yield types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst, lib=lib)


@pytest.mark.parametrize("strat", [1, 2, 3])
def test_pd3o_tv2d(dev, strat):
    G.test_pd3o_tv2d(dev, strat)
    assert "pds_iter:tile2d" in dev.lib.log and "pds_primal" not in dev.lib.log


def test_default_stop_iteration_count_and_cv(dev):
    G.test_pd3o_tv2d_default_stop_iteration_count(dev)
    G.test_cv_tv2d(dev)


@pytest.mark.parametrize("case", ["pd3o2d", "cv2d", "pd3o3d_maxiter_or_relerr", "pd3o_stacked_any", "maxiter_only"])
def test_iterations_queued_back_to_back_equal_one_launch_per_iteration(dev, case, monkeypatch):
    calls = []
    real = dev.lib.pxb_pds_iter_n
    dev.lib.pxb_pds_iter_n = lambda *a: (calls.append(a[9]), real(*a))[1]
    G.test_iterations_queued_back_to_back_equal_one_launch_per_iteration(dev, case, monkeypatch)
    assert len(calls) >= 1 and calls[0] == 8  # the batched loop did run (first batch: 8 iterations)


@pytest.mark.parametrize("mode", ["reflect", "wrap", "symmetric", "edge"])
def test_pd3o_tv2d_folding_modes_run_the_single_kernel_form(dev, mode):
    G.test_pd3o_tv2d_modes(dev, mode)
    assert "pds_iter:tile2d" in dev.lib.log and "pds_dual" not in dev.lib.log
    dev.lib.log.clear()
    dev.lib.pxb_set_iter_modes(0)  # ... and the two-sweep form when they are switched off
    G.test_pd3o_tv2d_modes(dev, mode)
    assert "pds_dual" in dev.lib.log and not any(s.startswith("pds_iter") for s in dev.lib.log)


def test_pd3o_tv3d_constant_and_mixed_modes(dev):
    G.test_pd3o_tv3d(dev)
    # 50 iterations 'constant' + 30 with (reflect, wrap, constant): TMA form; the fp32 run has 14 columns (not a multiple of
    # 4 samples): outside the single-kernel envelope -> 50 two-sweep iterations
    assert dev.lib.log.count("pds_iter:tma") == 80 and dev.lib.log.count("pds_dual") == 50


@pytest.mark.parametrize("tag", ["dense", "sep"])
def test_cv_deblur2d(dev, tag):
    G.test_cv_deblur2d(dev, tag)
    assert "stencil2d" in dev.lib.log


def test_cv_deblur3d_and_generic_path(dev):
    G.test_cv_deblur3d(dev)
    G.test_generic_path_equals_fused_path(dev)


@pytest.mark.parametrize("march", [False, True], ids=["per_plane_passes", "marching_kernel"])
def test_dense3d_golden(dev, march, monkeypatch):
    G.test_dense3d_stencil_golden(dev, march, monkeypatch)
    dev.lib.log.clear()
    G.test_cv_deblur3d_dense(dev, march, monkeypatch)
    assert ("stencil3d_dense" in dev.lib.log) is march and "stencil" not in dev.lib.log


@pytest.mark.parametrize("padded", [False, True])
def test_reflect_mode_blur_through_gather_and_padded_tiled_paths(dev, padded, monkeypatch):
    """PD3O / PGD with a reflect-mode blur in the data term: the gather kernels (default) and Stencil._run_padded
    (PYXU_B200_STENCIL_PADDED=1: pxb_pad2d -> tiled stencil -> pxb_pad2d_adjoint) must both reproduce the reference."""
    from pyxu_b200.operator.linop import stencil as st

    monkeypatch.setattr(st, "PADDED_TILED", padded)
    G.test_pd3o_deblur2d_semi_fused(dev)
    G.test_pgd_stacked_images_equal_per_image_solves(dev)
    assert ("pad2d" in dev.lib.log and "pad2d_adjoint" in dev.lib.log) == padded
    # CondatVu with the same blur: fused TV iteration + grad f through the two padded passes
    g = golden("solvers.npz")
    yb = g["cv_deblur2d/sep/y"]
    g9 = cases.gaussian_1d(9, 1.5)
    res = []
    for kern in (np.outer(g9, g9), [g9, g9]):
        slv, Aop = cases.build_tv_deblur(dev, yb, (28, 24), kern, (4, 4), lam=0.02, blur_mode="reflect", positivity=True)
        slv.fit(x0=np.zeros(yb.size), stop_crit=dev.stop.MaxIter(25))
        assert slv._astate.get("error") is None, slv._astate.get("error")
        assert (Aop._padded_ok is True) == padded
        res.append(slv.stats()[0]["x"])
    assert G.relerr(res[0], res[1]) < 1e-12  # dense and separable statements of the same blur


@pytest.mark.parametrize("acc", [True, False])
def test_pgd_l1_deconv(dev, acc):
    G.test_pgd_l1_deconv(dev, acc)
    assert "stencil2d_fista" in dev.lib.log


def test_pgd_default_stop_and_manual_async_modes(dev):
    G.test_pgd_default_stop_iteration_count(dev)
    G.test_manual_and_async_modes(dev)


@pytest.mark.parametrize("padded", [False, True])
def test_operator_fixtures_through_the_real_operators(dev, padded, monkeypatch):
    """Stencil / Convolve / Gradient / proximal-map fixtures of the real reference through the operator classes themselves
    (tests/test_gpu_operators.py's cases), with and without the padded tiled path for folding Stencil modes."""
    import test_gpu_operators as GO
    from pyxu_b200.operator.linop import stencil as st

    monkeypatch.setattr(st, "PADDED_TILED", padded)
    g = golden("stencil.npz")
    used = 0
    for case in cases.STENCIL_CASES:
        n = case["name"]
        for dtype, tol in ((np.float64, 1e-12), (np.float32, 5e-6)):
            op = cases.make_stencil(dev, case, dtype=dtype)
            out = op.apply(g[f"{n}/x"].astype(dtype))
            assert isinstance(out, np.ndarray) and out.dtype == dtype  # NumPy in -> NumPy out
            assert GO.relerr(out, g[f"{n}/apply"]) < tol and GO.relerr(op.adjoint(g[f"{n}/y"].astype(dtype)), g[f"{n}/adjoint"]) < tol, (n, dtype)
            used += op._padded_ok is True
    assert (used > 0) == padded
    for case in cases.GRADIENT_CASES:
        GO.test_gradient_golden(dev, case)
    GO.test_funcs_golden(dev)
    for shape, ks, modes in (((37, 53), (5, 5), ("reflect", "wrap")), ((19, 23, 17), (3, 4, 5), ("symmetric", "edge", "constant")), ((301,), (9,), ("wrap",)),
                             ((36, 64), (9, 9), ("reflect", "symmetric")), ((12, 20, 36), (7, 7, 7), ("constant", "reflect", "symmetric"))):
        GO.test_stencil_vs_oracle_random(dev, shape, ks, modes)


def test_3d_separable_stencil_with_folding_modes_padded_path(dev, monkeypatch):
    """3-D separable stencils with a folding mode on every axis through the real Stencil class with the padded tiled path on:
    streaming axis-0 pass with the boundary map (pxb_stencil_axis0_fold) + Pad -> tiled stencil / tiled stencil -> Pad^T,
    against the gather kernels and the fixture of the real reference (3d_sep_mixed)."""
    import test_gpu_operators as GO
    from pyxu_b200.operator.linop import stencil as st

    rng = np.random.default_rng(8)
    g7 = cases.gaussian_1d(7, 1.2)
    for shape, kern, cen, mode in (((9, 12, 32), [g7, cases.gaussian_1d(5, 1.0), g7], (3, 2, 3), "reflect"),
                                   ((9, 12, 32), [rng.standard_normal(4), g7, rng.standard_normal(3)], (0, 6, 2), ("wrap", "symmetric", "edge")),
                                   ((7, 10, 16), [rng.standard_normal(3), np.ones(1), g7], (2, 0, 3), ("symmetric", "constant", "reflect"))):
        x = rng.standard_normal((2, int(np.prod(shape))))
        outs = {}
        for padded in (True, False):
            monkeypatch.setattr(st, "PADDED_TILED", padded)
            op = dev.operator.Stencil(arg_shape=shape, kernel=kern, center=cen, mode=mode)
            dev.lib.log.clear()
            outs[padded] = (op.apply(x), op.adjoint(x))
            assert ("stencil_axis0_fold" in dev.lib.log and "pad2d" in dev.lib.log) == padded
        for a, b in zip(outs[True], outs[False]):
            assert GO.relerr(a, b) < 1e-13
        y = rng.standard_normal(x.shape)
        assert abs(np.vdot(outs[True][0], y) - np.vdot(x, op.adjoint(y) if False else dev.operator.Stencil(arg_shape=shape, kernel=kern, center=cen, mode=mode).adjoint(y))) < 1e-9
    monkeypatch.setattr(st, "PADDED_TILED", True)
    case = [c for c in cases.STENCIL_CASES if c["name"] == "3d_sep_mixed"][0]
    gold = golden("stencil.npz")
    # the fixture's last axis has 7 samples (not a multiple of the vector width): that operator stays on the gather kernels
    op = cases.make_stencil(dev, case)
    assert GO.relerr(op.apply(gold["3d_sep_mixed/x"]), gold["3d_sep_mixed/apply"]) < 1e-12 and op._padded_ok is False


def test_bench_usage_pattern(dev):
    """What bench.py does with the solver: Mode.MANUAL + ManualStop with m_step() driven by hand and a probe around the fused
    kernel (the kernel-only number), then fit(MaxIter | RelError) + solution() on host arrays (the end-to-end number)."""
    from pyxu_b200.abc.solver import Mode

    pxo, pxs, pxst = dev.operator, dev.solver, dev.stop
    shape = (8, 16, 64)
    N = int(np.prod(shape))
    y = np.random.default_rng(0).random(N).astype(np.float32)
    import torch

    yt = torch.from_numpy(y.copy())
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yt)
    Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
    h = 0.08 * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
    slv = pxs.PD3O(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False)
    slv.fit(x0=yt, mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
    assert slv._plan.kind == "fused"
    tags = []
    slv._probe = tags.append
    n0 = dev.lib.pxb_launch_count()
    for _ in range(7):
        slv.m_step()
    assert dev.lib.pxb_launch_count() - n0 == 7 and tags == ["iter_begin", "iter_end"] * 7 and slv._plan.iter_ok is True
    slv._probe = None
    x_manual = slv._materialize("x").clone()
    # end to end with host arrays
    f2 = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)
    slv2 = pxs.PD3O(f=f2, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False, final_writeback=False)
    slv2.fit(x0=y.copy(), stop_crit=pxst.MaxIter(7) | pxst.RelError(eps=1e-30, var="x"))
    assert slv2._astate.get("error") is None and "_fused_norms" in slv2._mstate
    x_host = slv2.solution()
    _, hist = slv2.stats()
    assert isinstance(x_host, np.ndarray) and x_host.dtype == np.float32 and len(hist) == 8 and np.isfinite(hist["RelError[x]"][1:]).all()
    assert np.allclose(x_host, x_manual.numpy(), rtol=1e-6, atol=1e-7)
    assert sum("Iteration" in ln for ln in open(slv2.logfile)) == 8


def test_opt_in_gpu_tests_are_themselves_sound(dev, monkeypatch):
    """The host-array tests of tests/test_gpu_zz_stencil_padded.py (skipped on a GPU box unless PYXU_B200_STENCIL_PADDED=1) run
    here on the emulated device, so that the first GPU run of the padded path does not stumble over the tests."""
    import test_gpu_zz_stencil_padded as GP
    from pyxu_b200.operator.linop import stencil as st

    monkeypatch.setattr(st, "PADDED_TILED", True)
    monkeypatch.setattr(GP, "DEV", "cpu")
    for case in cases.STENCIL_CASES:
        if case["mode"] != "constant" and case["arg_shape"][-1] % 2 == 0:
            GP.test_padded_golden(case)
    GP.test_cv_deblur_reflect_blur_uses_the_padded_path()
    small = [((37, 68), k, c, m) if len(sh) == 2 else ((sh[0], 18, 40) if len(sh) == 3 else sh, k, c, m) for sh, k, c, m in GP.CASES]
    monkeypatch.setattr(GP, "CASES", small)  # same operators on smaller arrays: the host replays every CTA
    for ci in range(len(small)):
        for dtype in (np.float64, np.float32):
            GP.test_padded_vs_generic(ci, dtype)


def test_gpu_tests_of_the_folding_modes_replayed(dev, monkeypatch):
    """tests/test_gpu_zz_iter_modes.py (single-kernel iteration with folding modes against the two-sweep kernels through the
    C ABI, full tiles included) with its arrays on the emulated device: the kernels' bodies AND the test code are checked before
    their next GPU run.  (The largest volume of the full-tile test is cut down here: the host replays every CTA.)"""
    import types

    import test_gpu_iter as GI
    import test_gpu_zz_iter_modes as GM
    from pyxu_b200 import _array as A
    from pyxu_b200 import _cabi as K

    monkeypatch.setattr(GI, "DEV", "cpu")
    env = types.SimpleNamespace(operator=dev.operator, solver=dev.solver, stop=dev.stop, A=A, K=K, lib=dev.lib)
    select = lambda p: K.check(dev.lib.pxb_set_iter_path(p), "pxb_set_iter_path")
    real_gradient = dev.operator.Gradient

    def small_gradient(arg_shape, **kw):  # 256 x 64 x 256 -> 24 x 64 x 256: same tiles per plane, fewer planes
        return real_gradient(arg_shape=(24,) + tuple(arg_shape[1:]) if tuple(arg_shape) == (256, 64, 256) else arg_shape, **kw)

    monkeypatch.setattr(dev.operator, "Gradient", small_gradient)
    for mode in ("reflect", ("wrap", "reflect", "symmetric")):
        GM.test_modes_full_tiles_3d_and_2d(env, mode, select)
    monkeypatch.setattr(dev.operator, "Gradient", real_gradient)
    GM.test_modes_vs_two_sweeps_3d(env, ("edge", "constant", "symmetric"), "tma", select)
    GM.test_modes_vs_two_sweeps_2d_batched(env, ("reflect", "wrap"), "tile2d", select)
    GM.test_modes_solver_fit_against_reference_fixtures(env, select)
    dev.lib.pxb_set_iter_path(0)


@pytest.fixture
def dev_rt():
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    with emulated_device(cuda_runtime=True) as lib:
        yield types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst, lib=lib)


def _tv3d(dev, shape, y, dtype, algo="PD3O", positivity=True, mode="constant", **kw):
    pxo, pxs = dev.operator, dev.solver
    N = int(np.prod(shape))
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)
    Kop = pxo.Gradient(arg_shape=shape, dtype=dtype, mode=mode)
    h = 0.08 * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
    g = pxo.PositiveOrthant(dim=N) if positivity else None
    return getattr(pxs, algo)(f=f, g=g, h=h, K=Kop, show_progress=False, final_writeback=kw.pop("final_writeback", False), **kw)


@pytest.mark.parametrize("case", ["maxiter", "maxiter_or_relerr", "relerr_fires_inside_the_epoch", "longer_than_one_epoch", "cv", "ragged_tail",
                                  "modes_fold", "modes_edge", "modes_ring", "declined"])
def test_streamed_fit_equals_the_ordinary_fit(dev_rt, case, monkeypatch):
    """fit(x0=<large host array>) on one GPU queues the first iterations as a wavefront over z-chunks behind the chunked upload
    (SlabTV.run_streamed): same iterate, same history, same stopping iteration as the one-upload, one-iteration-at-a-time loop."""
    dev = dev_rt
    pxst, PDS = dev.stop, dev.solver.PD3O.__mro__[1]
    dtype = np.float32
    shape = {"ragged_tail": (13, 8, 16), "declined": (12, 8, 14)}.get(case, (12, 8, 16))  # 14 columns: outside the single-kernel envelope in fp32
    y = (np.random.default_rng(3).random(int(np.prod(shape))) - 0.2).astype(dtype)
    algo = "CondatVu" if case == "cv" else "PD3O"
    crits = {
        "maxiter": lambda: pxst.MaxIter(7),
        "maxiter_or_relerr": lambda: pxst.MaxIter(7) | pxst.RelError(eps=1e-30, var="x"),
        "relerr_fires_inside_the_epoch": lambda: pxst.MaxIter(40) | (pxst.RelError(eps=2e-2, var="x") & pxst.RelError(eps=2e-2, var="z")),
        "longer_than_one_epoch": lambda: pxst.MaxIter(9) | pxst.RelError(eps=1e-30, var="z"),
        "cv": lambda: pxst.MaxIter(6) | pxst.RelError(eps=1e-30, var="x"),
        "ragged_tail": lambda: pxst.MaxIter(5),
    }
    mode = {"modes_fold": ("constant", "symmetric", "wrap"), "modes_edge": ("constant", "edge", "symmetric"), "modes_ring": ("constant", "reflect", "edge")}.get(case, "constant")
    if case.startswith("modes") or case == "declined":
        crits[case] = lambda: pxst.MaxIter(6) | pxst.RelError(eps=1e-30, var="x")
    monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1 << 62)
    ref = _tv3d(dev, shape, y, dtype, algo, mode=mode)
    ref.fit(x0=y.copy(), stop_crit=crits[case]())
    assert ref._astate.get("error") is None and ref._slab is None
    x_ref, (dref, href) = ref.solution(), ref.stats()
    monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1)
    monkeypatch.setattr(PDS, "_STREAM_PLANES", 2)
    monkeypatch.setattr(PDS, "_STREAM_EPOCH", 4 if case == "longer_than_one_epoch" else 32)
    calls = []
    from pyxu_b200 import slab

    real = slab.SlabTV.run_streamed
    monkeypatch.setattr(slab.SlabTV, "run_streamed", lambda self, *a, **k: (calls.append(a[2]), real(self, *a, **k))[1])
    slv = _tv3d(dev, shape, y, dtype, algo, mode=mode)
    slv.fit(x0=y.copy(), stop_crit=crits[case]())
    assert slv._astate.get("error") is None, slv._astate.get("error")
    assert slv._slab is not None and len(calls) >= 1
    x, (d, hist) = slv.solution(), slv.stats()
    assert isinstance(x, np.ndarray) and x.dtype == dtype
    assert len(hist) == len(href) and (hist["iteration"] == href["iteration"]).all()
    if case == "relerr_fires_inside_the_epoch":
        assert 2 < len(hist) - 1 < 32 and calls == [32, len(hist) - 1]  # speculated 32 iterations, redone up to the stopping one
    elif case == "longer_than_one_epoch":
        assert calls == [4] and len(hist) == 10
    elif case == "declined":
        assert slv._slab.fused is False and "pds_dual" in dev.lib.log
    for name in href.dtype.names[1:]:
        assert np.allclose(hist[name], href[name], rtol=1e-5, atol=1e-12), name
    # (z: the same kernels on the same numbers; x: written by the wave's last iteration here, rebuilt from the previous pair by the
    # two-sweep primal kernel when no criterion reads it -- the same formula, contracted differently)
    assert np.allclose(x, x_ref, rtol=2e-6, atol=1e-7)
    assert np.array_equal(d["z"], dref["z"])
    assert sum("Iteration" in ln for ln in open(slv.logfile)) == sum("Iteration" in ln for ln in open(ref.logfile))


def test_streamed_fit_is_declined_outside_its_envelope(dev_rt, monkeypatch):
    """A boundary mode that folds along axis 0, a device-resident x0, MANUAL mode, objective tracking: the ordinary path."""
    import torch

    from pyxu_b200.abc.solver import Mode

    dev = dev_rt
    pxst, PDS = dev.stop, dev.solver.PD3O.__mro__[1]
    monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1)
    monkeypatch.setattr(PDS, "_STREAM_PLANES", 2)
    shape, dtype = (12, 8, 16), np.float32
    y = np.random.default_rng(3).random(int(np.prod(shape))).astype(dtype)
    for kw, fit_kw, x0 in ((dict(mode=("reflect", "constant", "constant")), {}, y.copy()), ({}, {}, torch.from_numpy(y.copy())),
                           ({}, dict(mode=Mode.MANUAL), y.copy()), ({}, dict(track_objective=True), y.copy())):
        slv = _tv3d(dev, shape, y, dtype, **kw)
        slv.fit(x0=x0, stop_crit=pxst.MaxIter(3), **fit_kw)
        assert slv._astate.get("error") is None and slv._slab is None and slv._plan.kind == "fused"


def test_streamed_fit_brings_the_result_back_behind_the_wave(dev_rt, monkeypatch):
    """With a reserved pinned result buffer the last iteration's x is copied out chunk by chunk inside fit(); solution() hands that
    buffer out (once), stats() / a second solution() read the device."""
    import pyxu_b200
    from pyxu_b200 import _array as A_

    dev = dev_rt
    pxst, PDS = dev.stop, dev.solver.PD3O.__mro__[1]
    shape, dtype = (12, 8, 16), np.float64
    y = np.random.default_rng(5).random(int(np.prod(shape)))
    monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1 << 62)
    ref = _tv3d(dev, shape, y, dtype)
    ref.fit(x0=y.copy(), stop_crit=pxst.MaxIter(5))
    x_ref = ref.solution()
    monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1)
    monkeypatch.setattr(PDS, "_STREAM_PLANES", 4)
    import torch

    monkeypatch.setattr(A_, "reserve_host_results", lambda n, count=1: A_._RESULT_POOL.setdefault(int(n), []).extend(
        torch.empty(int(n), dtype=torch.uint8) for _ in range(count)))
    A_.reserve_host_results(y.nbytes)
    try:
        slv = _tv3d(dev, shape, y, dtype, final_writeback=True)
        slv.fit(x0=y.copy(), stop_crit=pxst.MaxIter(5))
        assert slv._astate.get("error") is None, slv._astate.get("error")
        assert slv._stream_out is not None and not A_._RESULT_POOL[y.nbytes]  # taken; the final writeback did not consume it
        x = slv.solution()
        assert slv._stream_out is None and np.allclose(x, x_ref, rtol=1e-13, atol=1e-15)
        assert np.array_equal(slv.solution(), x)  # second call: from the device
        assert np.array_equal(np.load(slv.datafile)["x"], x)
        del x
        import gc

        gc.collect()
        assert len(A_._RESULT_POOL[y.nbytes]) == 1  # the buffer is back in the pool
        assert np.isclose(float(slv.objective_func()), float(ref.objective_func()), rtol=1e-12)  # the public objective on the streamed state
        # a second fit() of the same object: the result of the first one, had it not been collected, must not be handed out again
        slv.fit(x0=y.copy(), stop_crit=pxst.MaxIter(5))
        assert slv._stream_out is not None
        slv.fit(x0=0.5 * y, stop_crit=pxst.MaxIter(5))  # (the uncollected buffer goes back to the pool, a new one is taken)
        x2 = slv.solution()
        ref2 = _tv3d(dev, shape, y, dtype)
        monkeypatch.setattr(PDS, "_STREAM_MIN_BYTES", 1 << 62)
        ref2.fit(x0=0.5 * y, stop_crit=pxst.MaxIter(5))
        assert np.allclose(x2, ref2.solution(), rtol=1e-13, atol=1e-15) and not np.allclose(x2, x_ref)
        del x2
        gc.collect()
        assert len(A_._RESULT_POOL[y.nbytes]) == 1
    finally:
        A_.release_host_results()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dense_3d_kernel_of_full_rank_runs_as_tiled_passes_per_kernel_plane(dev, dtype, monkeypatch):
    """Stencil with a dense 3-D kernel that is not an outer product ('constant' boundaries): one tiled dense 2-D pass per plane of
    the kernel, accumulated in place (Stencil._run_dense3d; what serves the kernels the marching kernel declines, selected here with
    PYXU_B200_DENSE3D_MARCH = 0) -- apply, adjoint, stacks, epilogue operand -- against the gather kernels and against the NumPy
    oracle's correlation."""
    import torch

    from pyxu_b200.operator.linop import stencil as st_mod

    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", False)

    import test_gpu_operators as GO
    from oracle import pyxu_oracle as orc

    pxo = dev.operator
    rng = np.random.default_rng(2)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for shape, ks, cen in (((9, 20, 24), (3, 3, 3), (1, 1, 1)), ((7, 12, 16), (5, 3, 4), (0, 2, 3)), ((4, 10, 8), (7, 2, 3), (6, 0, 1))):
        kern = rng.standard_normal(ks).astype(dtype)
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._dense3d_ok = False
        x = torch.from_numpy(rng.standard_normal((2, fast.dim)).astype(dtype))
        y = torch.from_numpy(rng.standard_normal(fast.dim).astype(dtype))
        for adj in (False, True):
            dev.lib.log.clear()
            a = fast.adjoint(x) if adj else fast.apply(x)
            c0 = ks[0] - 1 - cen[0] if adj else cen[0]
            live = sum(1 for a in range(ks[0]) if max(0, c0 - a) < min(shape[0], shape[0] - (a - c0)))  # kernel planes whose source plane exists
            assert fast._dense3d_ok is True and dev.lib.log.count("stencil2d") == 2 * live and "stencil" not in dev.lib.log
            b = slow.adjoint(x) if adj else slow.apply(x)
            assert slow._dense3d_ok is False
            assert GO.relerr(a.numpy(), b.numpy()) < tol, (shape, ks, adj)
        oref = orc.Stencil(shape, kern.astype(np.float64), cen, "constant")
        ref = oref.apply(x.numpy().astype(np.float64))
        assert GO.relerr(fast.apply(x).numpy(), ref) < tol and GO.relerr(fast.adjoint(x).numpy(), oref.adjoint(x.numpy().astype(np.float64))) < tol
        got = fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y)
        assert GO.relerr(got.numpy(), 0.5 * ref - y.numpy()) < tol
    # an outer product keeps the separable single pass; a folding mode keeps the gather kernels
    sep = pxo.Stencil(arg_shape=(9, 20, 24), kernel=np.ones((3, 3, 3), dtype=dtype), center=(1, 1, 1), mode="constant")
    sep.apply(torch.zeros(sep.dim, dtype=torch.float64 if dtype == np.float64 else torch.float32))
    assert sep._dense3d_ok is False and sep._tiled3d_ok is True
    fold = pxo.Stencil(arg_shape=(9, 20, 24), kernel=rng.standard_normal((3, 3, 3)).astype(dtype), center=(1, 1, 1), mode="reflect")
    fold.apply(torch.zeros(fold.dim, dtype=torch.float64 if dtype == np.float64 else torch.float32))
    assert fold._dense3d_ok in (None, False) and fold._tiled_ok is not True


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dense_3d_kernel_of_full_rank_through_the_marching_kernel(dev, dtype, monkeypatch):
    """The same operators with the marching kernel selected (PYXU_B200_DENSE3D_MARCH): ONE launch per direction inside its envelope,
    the per-plane passes outside it; results against the gather kernels and the NumPy oracle."""
    import torch

    import test_gpu_operators as GO
    from oracle import pyxu_oracle as orc
    from pyxu_b200.operator.linop import stencil as st_mod

    monkeypatch.setattr(st_mod, "DENSE3D_MARCH", True)
    pxo = dev.operator
    rng = np.random.default_rng(4)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for shape, ks, cen, march in (((9, 20, 24), (3, 3, 3), (1, 1, 1), True), ((13, 12, 16), (5, 4, 5), (0, 2, 3), True), ((12, 10, 8), (7, 7, 6), (6, 0, 1), True),
                                  ((4, 10, 8), (7, 2, 3), (6, 0, 1), False)):
        kern = rng.standard_normal(ks).astype(dtype)
        fast = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow = pxo.Stencil(arg_shape=shape, kernel=kern, center=cen, mode="constant")
        slow._dense3d_ok = False
        x = torch.from_numpy(rng.standard_normal((2, fast.dim)).astype(dtype))
        y = torch.from_numpy(rng.standard_normal(fast.dim).astype(dtype))
        for adj in (False, True):
            dev.lib.log.clear()
            a = fast.adjoint(x) if adj else fast.apply(x)
            if march:
                assert fast._march3d_ok is True and dev.lib.log == ["stencil3d_dense"], dev.lib.log
            else:
                assert fast._march3d_ok is False and "stencil3d_dense" not in dev.lib.log and "stencil2d" in dev.lib.log
            b = slow.adjoint(x) if adj else slow.apply(x)
            assert GO.relerr(a.numpy(), b.numpy()) < tol, (shape, ks, adj)
        oref = orc.Stencil(shape, kern.astype(np.float64), cen, "constant")
        ref = oref.apply(x.numpy().astype(np.float64))
        assert GO.relerr(fast.apply(x).numpy(), ref) < tol and GO.relerr(fast.adjoint(x).numpy(), oref.adjoint(x.numpy().astype(np.float64))) < tol
        got = fast._run_tiled(x, False, alpha=0.5, beta=-1.0, add=y)
        assert GO.relerr(got.numpy(), 0.5 * ref - y.numpy()) < tol
    # an outer product keeps the separable single pass
    sep = pxo.Stencil(arg_shape=(9, 20, 24), kernel=np.ones((3, 3, 3), dtype=dtype), center=(1, 1, 1), mode="constant")
    sep.apply(torch.zeros(sep.dim, dtype=torch.float64 if dtype == np.float64 else torch.float32))
    assert sep._march3d_ok is None and sep._tiled3d_ok is True
