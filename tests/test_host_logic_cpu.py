"""Host-side logic that needs no GPU: operator arithmetic (class inference, Lipschitz propagation),
descriptor construction, step-size selection, solver argument validation."""
import math
import types

import numpy as np
import pytest

import cases
import pyxu_b200.abc as pxa
import pyxu_b200.operator as pxo
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst
from conftest import golden
from oracle import pyxu_oracle as orc
from pyxu_b200 import _cabi as K


def test_arithmetic_class_inference_and_constants():
    N = 12
    sq = pxo.SquaredL2Norm(dim=N)
    f = 0.5 * sq.argshift(-np.arange(N, dtype=float))
    assert isinstance(f, pxa.QuadraticFunc) and f.diff_lipschitz == 1.0
    assert f._sql2_spec()[0] == 0.5 and np.array_equal(f._sql2_spec()[1], -np.arange(N))
    A = pxo.Stencil(arg_shape=(3, 4), kernel=np.ones((3, 3)), center=(1, 1))
    fA = f * A
    assert isinstance(fA, pxa.QuadraticFunc) and fA.shape == (1, N)
    assert math.isclose(fA.diff_lipschitz, A.lipschitz**2)
    h = 0.3 * pxo.L21Norm(arg_shape=(2, 3, 4))
    assert isinstance(h, pxa.ProxFunc) and not isinstance(h, pxa.DiffFunc)
    assert h._dual_spec() == (K.DUAL_L21, 0.3, (1, 2, 12))
    g = 2.0 * pxo.L1Norm(dim=N)
    assert g._prox_spec() == (K.PROX_L1, 2.0, 0.0) and math.isclose(g.lipschitz, 2 * math.sqrt(N))
    assert (3 * pxo.PositiveOrthant(dim=N))._prox_spec()[0] == K.PROX_POS
    assert (-1 * pxo.L1Norm(dim=N)).can_prox is False  # negative scaling loses proximability
    s = pxo.L1Norm(dim=N) + f
    assert isinstance(s, pxa.Func) and not s.can_prox
    AT = A.T
    assert AT.shape == (N, N) and AT.T is A and AT.lipschitz == A.lipschitz
    AA = A.T * A
    assert isinstance(AA, pxa.SquareOp) and math.isclose(AA.lipschitz, A.lipschitz**2)
    assert (A * 0).__class__.__name__ == "NullOp" and (A * 1) is A
    with pytest.raises(ValueError):
        f.argshift(np.zeros(N + 1))
    with pytest.raises(ValueError):
        A + pxo.Stencil(arg_shape=(5,), kernel=np.ones(2), center=(0,))


def test_stencil_constructor_validation_matches_reference():
    with pytest.raises(AssertionError):  # center outside kernel (reference: _stencil.py:123-125)
        pxo.Stencil(arg_shape=(8,), kernel=np.ones(3), center=(3,))
    with pytest.raises(AssertionError):  # reflect pad width limited to N-1 (reference: pad.py:217-229)
        pxo.Stencil(arg_shape=(4,), kernel=np.ones(6), center=(0,), mode="reflect")
    with pytest.raises(AssertionError):
        pxo.Stencil(arg_shape=(4, 4), kernel=np.ones(3), center=(0, 0))
    op = pxo.Stencil(arg_shape=(5, 6, 9), kernel=[np.r_[1, -1], np.r_[3, 2, 1], np.r_[2, -1, 3, 1]], center=(1, 0, 3))
    assert [r.tolist() for r in op.relative_indices] == [[-1, 0], [0, 1, 2], [-3, -2, -1, 0]]  # stencil.py:740-749
    assert op.center == (1, 0, 3)


@pytest.mark.parametrize("case", cases.GRADIENT_CASES, ids=lambda c: c["name"])
def test_fd_coefficients_match_reference(case):
    ns = types.SimpleNamespace(operator=pxo)
    op = cases.make_gradient(ns, case)
    ref = orc.Gradient(case["arg_shape"], mode=case["mode"], **case["diff_kwargs"])
    for t, c, rop, d in zip(op._taps, op._centers, ref.ops, op._dirs):
        assert np.allclose(t, rop._k[d].reshape(-1), rtol=0, atol=0) and c == rop._c[d][d]


def test_step_sizes_match_reference_fixture():
    g = golden("solvers.npz")
    ns = types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst)
    y = g["pd3o_tv2d/y"]
    for strat in (1, 2, 3):
        slv = cases.build_tv_denoise(ns, y, (32, 40), lam=0.1)
        slv._tuning_strategy = strat
        gamma = slv._set_gamma(strat)
        tau, sigma, delta = slv._set_step_sizes(None, None, gamma)
        rho = slv._set_momentum_term(None, delta)
        for k, v in (("tau", tau), ("sigma", sigma), ("rho", rho)):
            assert abs(v - float(g[f"pd3o_tv2d/s{strat}/{k}"])) < 1e-9, (strat, k)
        # closed form used by the oracle agrees with the LP
        t2, s2, r2 = orc.pd3o_step_sizes(slv._beta, slv._K.lipschitz, True, tuning_strategy=strat)
        assert abs(t2 - tau) < 1e-8 and abs(r2 - rho) < 1e-8
    slv = cases.build_tv_denoise(ns, y, (32, 40), lam=0.1, solver="CondatVu")
    slv._tuning_strategy = 1
    tau, sigma, delta = slv._set_step_sizes(None, None, slv._set_gamma(1))
    assert abs(tau - float(g["cv_tv2d/tau"])) < 1e-12 and abs(sigma - float(g["cv_tv2d/sigma"])) < 1e-12
    t2, s2, _ = orc.cv_step_sizes(slv._beta, slv._K.lipschitz, True, True)
    assert abs(t2 - tau) < 1e-14


def test_solver_argument_validation():
    with pytest.raises(ValueError):
        pxs.PD3O()
    with pytest.raises(ValueError):
        pxs.PGD()
    N = 6
    with pytest.raises(ValueError):  # K without h (reference: pds.py:72-77)
        pxs.CV(f=pxo.SquaredL2Norm(dim=N), K=pxo.Gradient(arg_shape=(N,)))
    with pytest.raises(ValueError):  # unbounded diff-Lipschitz (reference: pds.py:141-146)
        pxs.PD3O(f=pxo.L1Norm(dim=N).moreau_envelope(1.0) * 1.0 if False else _Unbounded(N))
    with pytest.raises(ValueError):
        pxst.MaxIter(0)
    with pytest.raises(ValueError):
        pxst.RelError(eps=-1)
    sc = pxst.MaxIter(3) & pxst.RelError(1e-3)
    assert sc._fused_vars() == {"x"} and sc._needs_host_sync()
    assert not pxst.MaxIter(3)._needs_host_sync()


class _Unbounded(pxa.DiffFunc):
    def __init__(self, n):
        super().__init__((1, n))


def test_bench_clock_sampler_windows_on_the_timed_region():
    """bench.py keeps the nvidia-smi samples taken inside the timed region (or the nearest one when the region is
    shorter than the sampling period) and reports throttle reasons only from those."""
    import importlib.util
    import os

    from conftest import ROOT

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    clk = bench.ClockSampler(0)
    row = lambda mhz, cap: ["0", str(mhz), "1965", "700", "Not Active", "Not Active", "Not Active", cap]
    clk.rows = [(10.0, row(1965, "Not Active")), (10.1, row(1800, "Active")), (10.2, row(1700, "Active")), (10.6, row(900, "Not Active"))]
    s = clk.summary(10.05, 10.25)
    assert s["samples_in_timed_region"] == 2 and s["sm_mhz"] == 1750.0 and s["reasons"] == ["sw_power_cap"] and s["sm_max_mhz"] == 1965.0
    s = clk.summary(10.30, 10.32)  # nothing inside: the nearest sample stands in
    assert s["samples_in_timed_region"] == 0 and s["samples"] == 1 and s["sm_mhz"] == 1700.0
    clk.rows = []
    assert clk.summary(0.0, 1.0)["samples"] == 0


def test_solver_loop_log_file_and_history_without_a_device():
    """The Solver loop itself (reference: abc/solver.py:571-718) on a toy iteration that needs no device: one history record and
    one log entry per iteration -- the entries are rendered in batches but the file reads as if each had been logged on its
    own --, the end-of-run message last, pending entries written before an exception is logged."""
    from pyxu_b200.abc.solver import Solver

    class Halver(Solver):
        def __init__(self, fail_at=None, **kw):
            super().__init__(log_var=("x",), final_writeback=False, **kw)
            self.fail_at = fail_at

        def m_init(self, x0):
            self._mstate["x"] = np.asarray(x0, dtype=float)

        def m_step(self):
            if self.fail_at is not None and self._astate["idx"] == self.fail_at:
                raise RuntimeError("boom")
            self._mstate["x"] = self._mstate["x"] / 2

        def solution(self):
            return self._mstate["x"]

    slv = Halver(show_progress=False)
    slv.fit(x0=np.ones(3), stop_crit=pxst.MaxIter(600))
    _, hist = slv.stats()
    assert len(hist) == 601 and hist["iteration"].tolist() == list(range(601)) and hist["N_iter"][-1] == 601
    lines = open(slv.logfile).read().splitlines()
    heads = [ln for ln in lines if ln.startswith("INFO -- [")]
    its = [int(ln.rsplit("Iteration ", 1)[1].replace("_", "")) for ln in heads if "Iteration" in ln]
    assert its == list(range(601))                       # every iteration once, in order, across several batches
    assert heads[-1].endswith("Stopping Criterion satisfied -> END") and lines[-1] == heads[-1]
    assert lines[1] == "\titeration: 0" and lines[2] == "\tN_iter: 1.0"
    assert np.allclose(slv.solution(), 0.5**600)
    # an exception inside m_step: the entries kept back are on disk before the traceback
    slv = Halver(fail_at=5, show_progress=False)
    slv.fit(x0=np.ones(3), stop_crit=pxst.MaxIter(50))
    assert isinstance(slv._astate.get("error"), RuntimeError)
    text = open(slv.logfile).read()
    assert text.index("Iteration 4") < text.index("EXCEPTION RAISED") and "boom" in text
    # show_progress=True (BLOCK mode streams to stdout): entries are logged one by one, same file
    slv = Halver(show_progress=True)
    slv.fit(x0=np.ones(3), stop_crit=pxst.MaxIter(3))
    assert sum("Iteration" in ln for ln in open(slv.logfile)) == 4


def test_pd3o_step_size_rule_equals_the_reference_linear_program():
    """PD3O default step sizes: the reference calls scipy.optimize.linprog (pds.py:831-864); here the program's
    closed-form solution is used.  Same numbers, bit for bit, on random operator norms and gammas."""
    from scipy.optimize import linprog

    rng = np.random.default_rng(0)
    N = 12
    for _ in range(60):
        L, beta = float(np.exp(rng.uniform(-3, 5))), float(np.exp(rng.uniform(-4, 4)))
        Kop = pxo.Gradient(arg_shape=(3, 4))
        Kop.lipschitz = L
        for klass, strat in ((pxs.PD3O, 1), (pxs.PD3O, 2), (pxs.PD3O, 3)):
            slv = klass(f=0.5 * beta * pxo.SquaredL2Norm(dim=N), g=None, h=pxo.L21Norm(arg_shape=(2, 3, 4), l2_axis=(0,)), K=Kop, show_progress=False)
            gamma = slv._set_gamma(strat)
            got = slv._optimize_step_sizes(gamma)
            b_ub = np.array([np.log(0.99) - 2 * np.log(L), np.log(1 / gamma)])
            ref = linprog(c=np.array([-1, -1]), A_ub=np.array([[1, 1], [1, 0]]), b_ub=b_ub, A_eq=np.array([[1, -1]]), b_eq=np.array([0]),
                          bounds=(None, None))
            assert ref.success and np.array_equal(got, np.exp(ref.x)), (L, gamma, got, np.exp(ref.x))
