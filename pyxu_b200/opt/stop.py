"""
Stopping criteria (reference: src/pyxu/opt/stop.py -- MaxIter:29, ManualStop:71, MaxDuration:93,
Memorize:181, AbsError:222, RelError:300).

RelError / AbsError evaluate their norms on the device.  When the solver's update kernels already
accumulated ||x_k - x_{k-1}||^2 and ||x_{k-1}||^2 (state["_fused_norms"][var], see
pyxu_b200.opt.solver), those are used and no extra pass over the volume is made; otherwise one
pxb_sqnorms launch does it.  Only a few scalars cross the PCIe bus per evaluation.
"""
import datetime as dt
import warnings

import numpy as np

from .. import _kernels as kr
from ..abc.solver import StoppingCriterion


def _is_real(x):
    return isinstance(x, (int, float, np.integer, np.floating))


class MaxIter(StoppingCriterion):
    def __init__(self, n):
        try:
            assert int(n) > 0
            self._n = int(n)
        except Exception:
            raise ValueError(f"n: expected positive integer, got {n}.")
        self._i = 0

    def stop(self, state):
        self._i += 1
        return self._i > self._n

    def info(self):
        return dict(N_iter=self._i)

    def clear(self):
        self._i = 0

    def _device_eval(self, px, pz):
        return False

    def _budget(self):
        return self._n - self._i + 1  # iterations that may run before the test after the last of them can be the one that fires

    def _replay(self, sums):
        calls = self._i + 1 + np.arange(len(sums))
        self._i += len(sums)
        return calls > self._n, dict(N_iter=calls)

    def _state_vars(self):
        return frozenset()

    def _needs_host_sync(self):
        return False


class ManualStop(StoppingCriterion):
    def stop(self, state):
        return False

    def info(self):
        return dict()

    def _device_eval(self, px, pz):
        return False

    def _replay(self, sums):
        return np.zeros(len(sums), dtype=bool), dict()

    def _state_vars(self):
        return frozenset()

    def _needs_host_sync(self):
        return False


class MaxDuration(StoppingCriterion):
    def __init__(self, t):
        try:
            assert t > dt.timedelta()
            self._t_max = t
        except Exception:
            raise ValueError(f"t: expected positive duration, got {t}.")
        self._t_start = dt.datetime.now()
        self._t_now = self._t_start

    def stop(self, state):
        self._t_now = dt.datetime.now()
        return (self._t_now - self._t_start) > self._t_max

    def info(self):
        return dict(duration=(self._t_now - self._t_start).total_seconds())

    def _rank_local(self):
        return True

    def clear(self):
        self._t_start = dt.datetime.now()
        self._t_now = self._t_start

    def _state_vars(self):
        return frozenset()

    def _needs_host_sync(self):
        return False


class Memorize(StoppingCriterion):
    def __init__(self, var):
        self._var = var
        self._val = np.r_[0]

    def stop(self, state):
        x = state[self._var]
        if _is_real(x):
            x = np.r_[x]
        if hasattr(x, "is_cuda"):
            x = x.detach().cpu().numpy()
        assert x.ndim == 1
        self._val = np.asarray(x)
        return False

    def info(self):
        if self._val.size == 1:
            return {f"Memorize[{self._var}]": float(self._val.max())}
        return {f"Memorize[{self._var}]_min": float(self._val.min()), f"Memorize[{self._var}]_max": float(self._val.max())}

    def clear(self):
        self._val = np.r_[0]

    def _state_vars(self):
        return frozenset([self._var])


def _device_norm(x, ord):
    """(rows, 1) host array of L`ord` norms of the rows of device array x."""
    rows = 1 if x.dim() == 1 else int(x.numel() // x.shape[-1])
    if ord == 2:
        return np.sqrt(kr.sqnorms(x, rows=rows)[:, :1].cpu().numpy())
    import torch  # non-Euclidean norms: rare, off the hot path

    return torch.linalg.vector_norm(x.reshape(rows, -1).double(), ord=ord, dim=-1, keepdim=True).cpu().numpy()


def _fused_host(fused, var):
    """Host copy of the (rows, 2) sums of `var`; all fused sums of an iteration cross the bus in ONE readback
    (the solver resets `_stamp` when it zeroes the buffer for the next iteration)."""
    if "_all" not in fused:
        return fused[var].cpu().numpy()
    if fused["_stamp"] != 0:
        fused["_host"] = fused["_all"].cpu().numpy()
        fused["_stamp"] = 0
    return fused["_host"][0 if var == "x" else 1]


class _NormCriterion(StoppingCriterion):
    def __init__(self, eps, var="x", f=None, norm=2, satisfy_all=True):
        try:
            assert eps > 0
            self._eps = eps
        except Exception:
            raise ValueError(f"eps: expected positive threshold, got {eps}.")
        self._var = var
        self._f = f
        try:
            assert norm >= 0
            self._norm = norm
        except Exception:
            raise ValueError(f"norm: expected non-negative, got {norm}.")
        self._satisfy_all = satisfy_all
        self._val = np.r_[0]

    def _label(self):
        return f"{type(self).__name__}[{self._var}]"

    def _state_vars(self):
        return frozenset([self._var])

    def info(self):
        if self._val.size == 1:
            return {self._label(): float(self._val.max())}
        return {f"{self._label()}_min": float(self._val.min()), f"{self._label()}_max": float(self._val.max())}

    def _rule(self, b):
        return bool(np.all(b) if self._satisfy_all else np.any(b))


class AbsError(_NormCriterion):
    def stop(self, state):
        x = state[self._var]
        if self._f is not None:
            x = self._f(x)
        if _is_real(x):
            self._val = np.abs(np.r_[x]).reshape(1, 1)
        elif hasattr(x, "is_cuda"):
            self._val = _device_norm(x, self._norm)
        else:
            self._val = np.linalg.norm(np.asarray(x), ord=self._norm, axis=-1, keepdims=True)
        return self._rule(self._val <= self._eps)

    def clear(self):
        self._val = np.r_[0]


class RelError(_NormCriterion):
    def __init__(self, eps, var="x", f=None, norm=2, satisfy_all=True):
        super().__init__(eps, var, f, norm, satisfy_all)
        self._x_prev = None
        self._started = False

    def _fused_vars(self):
        return frozenset([self._var]) if (self._f is None and self._norm == 2) else frozenset()

    def _device_eval(self, px, pz):
        if not self._fused_vars() or self._var not in ("x", "z"):
            raise NotImplementedError
        return px if self._var == "x" else pz

    def _device_leaves(self):
        return [(self._var, float(self._eps), bool(self._satisfy_all))]

    def _replay(self, sums):
        if not (self._started and self._fused_vars()):
            raise NotImplementedError
        sq = np.asarray(sums)[:, 0 if self._var == "x" else 1]  # (m, rows, 2)
        num, den = np.sqrt(sq[..., 0]), np.sqrt(sq[..., 1])
        ok = num <= self._eps * den
        decisions = ok.all(axis=1) if self._satisfy_all else ok.any(axis=1)
        with np.errstate(all="ignore"):
            val = num / den
        val[np.isnan(val)] = 0
        self._val = val[-1].reshape(-1, 1)
        if val.shape[1] == 1:
            return decisions, {self._label(): val[:, 0]}
        return decisions, {f"{self._label()}_min": val.min(axis=1), f"{self._label()}_max": val.max(axis=1)}

    def stop(self, state):
        fused = state.get("_fused_norms", {}).get(self._var) if isinstance(state, dict) else None
        if fused is not None and self._f is None and self._norm == 2:
            # (rows, 2) device doubles: sum (x_k - x_{k-1})^2, sum x_{k-1}^2 -- accumulated by the update kernels
            if not self._started:
                self._started = True
                self._val = np.zeros((fused.shape[0], 1))
                return False
            sq = _fused_host(state["_fused_norms"], self._var)
            num, den = np.sqrt(sq[:, :1]), np.sqrt(sq[:, 1:2])
            return self._decide(num, den)

        x = state[self._var]
        if _is_real(x):
            x = np.r_[x]
        on_dev = hasattr(x, "is_cuda")
        if self._x_prev is None:
            self._x_prev = x.clone() if on_dev else np.array(x, copy=True)
            self._started = True
            self._val = np.zeros(shape=(1,) if (x.ndim == 1) else tuple(x.shape[:-1]))
            return False
        fx, fx_prev = (x, self._x_prev) if self._f is None else (self._f(x), self._f(self._x_prev))
        if on_dev and self._norm == 2 and hasattr(fx, "is_cuda"):
            rows = 1 if fx.dim() == 1 else int(fx.numel() // fx.shape[-1])
            sq = kr.sqnorms(fx, fx_prev, rows=rows).cpu().numpy()
            num, den = np.sqrt(sq[:, :1]), np.sqrt(sq[:, 1:2])
        elif on_dev:
            num, den = _device_norm(fx - fx_prev, self._norm), _device_norm(fx_prev, self._norm)
        else:
            nrm = lambda _: np.linalg.norm(np.atleast_1d(_), ord=self._norm, axis=-1, keepdims=True)
            num, den = nrm(fx - fx_prev), nrm(fx_prev)
        if on_dev:
            self._x_prev.copy_(x)
        else:
            self._x_prev = np.array(x, copy=True)
        return self._decide(num, den)

    def _decide(self, num, den):
        decision = self._rule(num <= self._eps * den)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                val = num / den
            val[np.isnan(val)] = 0  # 0/0: no relative improvement (stop.py:373-378)
        self._val = val
        return decision

    def clear(self):
        self._val = np.r_[0]
        self._x_prev = None
        self._started = False


__all__ = ["AbsError", "ManualStop", "MaxDuration", "MaxIter", "Memorize", "RelError"]
