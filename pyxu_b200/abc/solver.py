"""
Solver / StoppingCriterion base classes
(reference: src/pyxu/abc/solver.py -- Mode:27, StoppingCriterion:37, Solver:119).

Same user-facing protocol: `fit(**kwargs, stop_crit=..., mode=..., track_objective=...)`,
`steps()`, `stats()`, `solution()`, `busy()`, `stop()`, `writeback()`, `workdir / logfile /
datafile`; sub-classes implement `m_init()` / `m_step()`.  The mathematical state `_mstate` holds
device buffers; `stats()` hands arrays back in the memory space `fit()` received them in
(NumPy in -> NumPy out).  Stopping criteria are tested before each step, as in the reference
(solver.py:588-652).
"""
import datetime as dt
import enum
import logging
import operator
import pathlib as plib
import shutil
import sys
import tempfile
import threading

import numpy as np

from .. import _array as A


@enum.unique
class Mode(enum.Enum):
    BLOCK = enum.auto()
    MANUAL = enum.auto()
    ASYNC = enum.auto()


class StoppingCriterion:
    def stop(self, state):
        raise NotImplementedError

    def info(self):
        raise NotImplementedError

    def clear(self):
        pass

    def __or__(self, other):
        return _Composition(self, other, operator.or_)

    def __and__(self, other):
        return _Composition(self, other, operator.and_)

    # variables whose ||x_k - x_{k-1}||, ||x_{k-1}|| the solver may accumulate inside its update kernels
    def _fused_vars(self):
        return frozenset()

    def _needs_host_sync(self):
        return True

    # keys of the solver's mathematical state this criterion reads; None = unknown (assume all of them)
    def _state_vars(self):
        return None

    # True when the decision can differ between the ranks of a z-slab decomposed solve (e.g. a wall-clock limit)
    def _rank_local(self):
        return False

    # -- iterations queued back to back with the rule tested on the device (pxb_pds_iter_n) ----------------------------
    # _device_eval(px, pz): the decision given the outcomes of the RelError tests on x and z, every other leaf taken as
    #     "not met"; raises NotImplementedError when the criterion cannot be expressed that way.
    # _device_leaves(): [(var, eps, satisfy_all)] of the RelError leaves.
    # _budget(): how many iterations may run before a leaf the device does not evaluate (MaxIter) could fire.
    # _replay(sums): the decisions and info() fields of len(sums) consecutive stop() calls, sums[i] = the fused sums
    #     (2, rows, 2) of call i; leaves the criterion in the state those calls would have left it in.
    def _device_eval(self, px, pz):
        raise NotImplementedError

    def _device_leaves(self):
        return []

    def _budget(self):
        return float("inf")

    def _replay(self, sums):
        raise NotImplementedError


class _Composition(StoppingCriterion):
    def __init__(self, lhs, rhs, op):
        self._lhs, self._rhs, self._op = lhs, rhs, op

    def stop(self, state):
        return self._op(self._lhs.stop(state), self._rhs.stop(state))

    def info(self):
        return {**self._lhs.info(), **self._rhs.info()}

    def clear(self):
        self._lhs.clear()
        self._rhs.clear()

    def _fused_vars(self):
        return self._lhs._fused_vars() | self._rhs._fused_vars()

    def _needs_host_sync(self):
        return self._lhs._needs_host_sync() or self._rhs._needs_host_sync()

    def _state_vars(self):
        a, b = self._lhs._state_vars(), self._rhs._state_vars()
        return None if (a is None or b is None) else (a | b)

    def _rank_local(self):
        return self._lhs._rank_local() or self._rhs._rank_local()

    def _device_eval(self, px, pz):
        return bool(self._op(self._lhs._device_eval(px, pz), self._rhs._device_eval(px, pz)))

    def _device_leaves(self):
        return self._lhs._device_leaves() + self._rhs._device_leaves()

    def _budget(self):
        return min(self._lhs._budget(), self._rhs._budget())

    def _replay(self, sums):
        (da, ia), (db, ib) = self._lhs._replay(sums), self._rhs._replay(sums)
        return self._op(da, db), {**ia, **ib}


def _workdir(folder, exist_ok):
    """The solver's scratch directory: a fresh temporary one, or `folder` (emptied; refused when it exists unless exist_ok)."""
    if folder is None:
        return plib.Path(tempfile.mkdtemp(prefix="pyxu_"))
    try:
        path = plib.Path(folder).expanduser().resolve()
    except TypeError:
        raise TypeError(f"folder: a path is expected, not {type(folder).__name__}") from None
    if path.exists() and not exist_ok:
        raise FileExistsError(f"{path} already exists.")
    shutil.rmtree(path, ignore_errors=True)
    path.mkdir(parents=True)
    return path


def _rate(name, value, base, default=None):
    """A logging / checkpoint period: a positive integer that is a multiple of `base` (stop_rate), or `default` when None."""
    if value is None:
        return default
    try:
        ok = value >= 1 and int(value) == value and int(value) % base == 0
    except TypeError:
        ok = False
    if not ok:
        what = "a positive integer" if base == 1 else f"a positive multiple of stop_rate ({base})"
        raise ValueError(f"{name}: {what} is expected, got {value!r}")
    return int(value)


class Solver:
    def __init__(self, *, folder=None, exist_ok=False, stop_rate=1, writeback_rate=None, verbosity=None,
                 show_progress=True, log_var=frozenset(), final_writeback=True):
        """Same parameters as the reference (solver.py:188-236) plus `final_writeback`: the reference always dumps
        the logged variables to <workdir>/data.npz when the solver stops; for multi-GiB volumes that disk write
        dominates everything else, so it can be switched off here (default: on, as in the reference)."""
        self._final_writeback = bool(final_writeback)
        self._mstate = dict()
        self._astate = dict(history=None, idx=0, log_rate=None, log_var=None, logger=None, stdout=None, stop_crit=None,
                            stop_rate=None, track_objective=None, wb_rate=None, workdir=None, mode=None, active=None,
                            worker=None, origin=A.DEVICE)
        ast = self._astate
        ast["workdir"] = _workdir(folder, exist_ok)
        ast["stop_rate"] = _rate("stop_rate", stop_rate, 1)
        if ast["stop_rate"] is None:
            raise ValueError("stop_rate: a positive integer is expected, got None")
        ast["wb_rate"] = _rate("writeback_rate", writeback_rate, ast["stop_rate"])
        ast["log_rate"] = _rate("verbosity", verbosity, ast["stop_rate"], default=ast["stop_rate"])
        ast["stdout"] = bool(show_progress)
        if isinstance(log_var, str):
            log_var = (log_var,)
        try:
            ast["log_var"] = frozenset(log_var)
        except TypeError:
            raise ValueError(f"log_var: a collection of variable names is expected, got {type(log_var).__name__}") from None

    # -- user API ---------------------------------------------------------------------------
    def fit(self, **kwargs):
        import time

        t0 = time.perf_counter()
        self._fit_init(
            mode=kwargs.pop("mode", Mode.BLOCK),
            stop_crit=kwargs.pop("stop_crit", None),
            track_objective=kwargs.pop("track_objective", False),
        )
        t1 = time.perf_counter()
        self.m_init(**kwargs)
        t2 = time.perf_counter()
        self._fit_run()
        # host-side wall clock of the three phases (the device may still be working when BLOCK mode returns with a
        # criterion that never synchronises); bench.py reports it as the end-to-end breakdown
        self._astate["timing"] = dict(setup_s=t1 - t0, m_init_s=t2 - t1, run_s=time.perf_counter() - t2)

    def m_init(self, **kwargs):
        raise NotImplementedError

    def m_step(self):
        raise NotImplementedError

    def steps(self, n=None):
        self._check_mode(Mode.MANUAL)
        i = 0
        while (n is None) or (i < n):
            if self._step():
                data, _ = self.stats()
                yield data
                i += 1
            else:
                self._astate["mode"] = None
                self._cleanup_logger()
                return

    def _materialize(self, name):
        """Hook: bring a lazily-maintained state variable up to date (fused solvers override)."""
        return self._mstate.get(name)

    def _logged(self, k):
        """Logged variable `k` in the memory space fit() received its arrays in (None if unknown)."""
        if k not in self._astate["log_var"]:
            return None
        v = self._materialize(k)
        if v is not None and hasattr(v, "is_cuda"):
            v = A.restore(v, self._astate["origin"])
        return v

    def stats(self):
        history = self._astate["history"]
        if history is not None:
            history = np.concatenate(history, dtype=history[0].dtype, axis=0) if len(history) > 0 else None
        # (sorted: under a z-slab decomposition every variable is a collective gather, and the iteration order of a
        # frozenset of strings differs between the ranks' interpreters)
        data = {k: self._logged(k) for k in sorted(self._astate["log_var"])}
        return data, history

    @property
    def workdir(self):
        return self._astate["workdir"]

    @property
    def logfile(self):
        return self.workdir / "solver.log"

    @property
    def datafile(self):
        return self.workdir / "data.npz"

    def busy(self):
        self._check_mode(Mode.ASYNC, Mode.BLOCK)
        return self._astate["active"].is_set()

    def solution(self):
        raise NotImplementedError

    def stop(self):
        self._check_mode(Mode.ASYNC, Mode.BLOCK)
        self._astate["active"].clear()
        self._astate["worker"].join()
        self._astate.update(mode=None, active=None, worker=None)
        self._cleanup_logger()

    def writeback(self):
        data, history = self.stats()
        kwargs = {}
        for k, v in dict(history=history, **data).items():
            if v is None:
                continue
            v = getattr(v, "local", v)  # a ShardedArray: every rank dumps its own planes
            kwargs[k] = v.cpu().numpy() if hasattr(v, "is_cuda") else np.asarray(v)
        np.savez(self.datafile, **kwargs)

    def default_stop_crit(self):
        raise NotImplementedError("No default stopping criterion defined.")

    def objective_func(self):
        raise NotImplementedError("No objective function defined.")

    # -- internals --------------------------------------------------------------------------
    def _fit_init(self, mode, stop_crit, track_objective):
        def _init_logger():
            logger = logging.getLogger(str(self.workdir))
            logger.handlers.clear()
            logger.setLevel("DEBUG")
            fmt = logging.Formatter(fmt="{levelname} -- {message}", style="{")
            handlers = [logging.FileHandler(self.logfile, mode="w")]
            if (mode is Mode.BLOCK) and self._astate["stdout"]:
                handlers.append(logging.StreamHandler(sys.stdout))
            for h in handlers:
                h.setLevel("DEBUG")
                h.setFormatter(fmt)
                logger.addHandler(h)
            return logger

        self._mstate.clear()
        if stop_crit is None:
            stop_crit = self.default_stop_crit()
        stop_crit.clear()
        if track_objective:
            from ..opt.stop import Memorize

            stop_crit |= Memorize(var="objective_func")
        self._astate.update(history=[], idx=0, logger=_init_logger(), stop_crit=stop_crit, track_objective=track_objective,
                            mode=mode, active=None, worker=None, pending_log=[], history_dtype=None,
                            live_log=(mode is Mode.BLOCK) and self._astate["stdout"])  # a stream handler shows every record as it comes

    def _fit_run(self):
        mode = self._astate["mode"]
        if mode is Mode.MANUAL:
            return
        if mode is Mode.BLOCK:
            # The reference runs BLOCK mode in a worker thread it immediately joins (solver.py:541-560); the
            # loop is run inline here so that CUDA's thread-local current device / stream stay the caller's.
            while self._step():
                pass
            self._astate.update(mode=None, active=None, worker=None)
            self._cleanup_logger()
            return
        import torch

        self._astate["device"] = torch.cuda.current_device() if torch.cuda.is_available() else None
        self._astate.update(active=threading.Event(), worker=Solver._Worker(self))
        self._astate["active"].set()
        self._astate["worker"].start()

    def _check_mode(self, *modes):
        m = self._astate["mode"]
        if m in modes:
            return
        if m is None:
            msg = "Illegal method call: invoke Solver.fit() first."
        else:
            msg = " ".join(["Illegal method call: can only be used if Solver.fit() invoked with",
                            "mode=Any[" + ", ".join(map(lambda _: str(_.name), modes)) + "]"])
        raise ValueError(msg)

    def _step(self):
        """One turn of the loop (reference: solver.py:588-667): test the criterion on the current state, log, checkpoint,
        then iterate.  Returns False once the criterion is met or an exception was raised (kept in _astate["error"])."""
        try:
            if not self._pre_step():
                return False
            self._astate["idx"] += 1
            self.m_step()
            return True
        except Exception as e:
            self._on_error(e)
            return False

    def _pre_step(self):
        """What precedes an iteration: criterion test (on the stop_rate beat), history record, log line, checkpoint.
        False when the criterion is met (the closing log line and the final writeback are then done)."""
        ast = self._astate
        idx = ast["idx"]
        if idx % ast["stop_rate"] == 0:
            if ast["track_objective"]:
                self._mstate["objective_func"] = self.objective_func().reshape(-1)
            met = ast["stop_crit"].stop(self._mstate)
            self._record_history()
            if met:
                self._log_iteration()
                self._log_message(f"[{dt.datetime.now()}] Stopping Criterion satisfied -> END")
                if self._final_writeback:
                    self.writeback()
                return False
        if idx % ast["log_rate"] == 0:
            self._log_iteration()
        if ast["wb_rate"] is not None and idx % ast["wb_rate"] == 0:
            self.writeback()
        return True

    def _record_history(self):
        ast = self._astate
        data = ast["stop_crit"].info()
        keys = tuple(data)
        cache = ast.get("history_dtype")
        if cache is None or cache[0] != keys:  # the fields never change during a fit(): build the record type once
            cache = ast["history_dtype"] = (keys, np.dtype([("iteration", np.int64)] + [(k, np.float64) for k in keys]))
        rec = np.zeros(1, dtype=cache[1])
        rec["iteration"] = ast["idx"]
        for k, v in data.items():
            rec[k] = v
        ast["history"].append(rec)

    def _log_iteration(self):
        ast = self._astate
        when, rec = dt.datetime.now(), ast["history"][-1][0]
        if ast.get("live_log"):
            ast["logger"].info(self._render_iteration(when, ast["idx"], rec))
            return
        # The per-iteration record goes to the log FILE only: keep (time, record) and render the same text in batches
        # (Python's logging costs ~100 us per record -- more than an iteration of a small problem takes on the GPU).
        # Nothing is lost: the batch is written before any other message, at the latest every second / 256 records, on
        # stop, on error and on cleanup.
        pend = ast.setdefault("pending_log", [])
        pend.append((when, ast["idx"], rec))
        if len(pend) >= 256 or (when - pend[0][0]).total_seconds() > 1.0:
            self._flush_log()

    def _record_block(self, first_idx, info):
        """History records (and log lines) of consecutive criterion tests replayed in one go: `info` maps the criterion's
        fields to arrays, one entry per test; the test of iteration first_idx + i is entry i."""
        ast = self._astate
        keys = tuple(info)
        m = len(next(iter(info.values()))) if info else 0
        if m == 0:
            return
        cache = ast.get("history_dtype")
        if cache is None or cache[0] != keys:
            cache = ast["history_dtype"] = (keys, np.dtype([("iteration", np.int64)] + [(k, np.float64) for k in keys]))
        recs = np.zeros(m, dtype=cache[1])
        recs["iteration"] = first_idx + np.arange(m)
        for k, v in info.items():
            recs[k] = v
        ast["history"].append(recs)
        sel = recs if ast["log_rate"] == 1 else recs[recs["iteration"] % ast["log_rate"] == 0]
        if len(sel):
            if ast.get("live_log"):
                for r in sel:
                    ast["logger"].info(self._render_iteration(dt.datetime.now(), int(r["iteration"]), r))
            else:
                ast.setdefault("pending_log", []).append((dt.datetime.now(), None, sel))
                if len(ast["pending_log"]) >= 256:
                    self._flush_log()

    def _log_message(self, msg):
        self._flush_log()
        self._astate["logger"].info(msg)

    def _on_error(self, e):
        ast = self._astate
        lines = [f"[{dt.datetime.now()}] Something went wrong -> EXCEPTION RAISED"]
        print(lines[0], f"More information: {self.logfile}.", sep="\n", file=sys.stderr)
        if ast["wb_rate"] is not None:
            lines.append(f"Last valid checkpoint done at iteration={ast['idx'] - ast['idx'] % ast['wb_rate']}.")
        self._flush_log()
        ast["logger"].exception("\n".join(lines), exc_info=e)
        ast["error"] = e

    @staticmethod
    def _render_iteration(when, idx, h):
        msg = [f"[{when}] Iteration {idx:>_d}"]
        for field, value in zip(h.dtype.names, h):
            msg.append(f"\t{field}: {value}")
        return "\n".join(msg)

    def _flush_log(self):
        """Writes the iteration records kept back by _step() to the log file, in the format the logger gives them."""
        pend = self._astate.get("pending_log")
        if not pend:
            return
        logger = self._astate.get("logger")
        parts = []
        for when, idx, rec in pend:
            if idx is None:  # a block of records (see _record_block): one line group per row
                names = rec.dtype.names[1:]
                head = f"INFO -- [{when}] Iteration "
                for row in rec.tolist():
                    parts.append(head + f"{row[0]:>_d}\n\titeration: {row[0]}" + "".join(f"\n\t{n}: {v}" for n, v in zip(names, row[1:])) + "\n")
            else:
                parts.append(f"INFO -- {self._render_iteration(when, idx, rec)}\n")
        text = "".join(parts)
        pend.clear()
        for h in (logger.handlers if logger is not None else ()):
            if isinstance(h, logging.FileHandler):
                h.acquire()
                try:
                    if h.stream is None:
                        h.stream = h._open()
                    h.stream.write(text)
                    h.flush()
                finally:
                    h.release()

    def _cleanup_logger(self):
        self._flush_log()
        logger = logging.getLogger(str(self.workdir))
        for handler in logger.handlers:
            handler.close()

    class _Worker(threading.Thread):
        def __init__(self, solver):
            super().__init__()
            self.slvr = solver

        def run(self):
            import torch

            dev = self.slvr._astate.get("device")
            if dev is not None:
                torch.cuda.set_device(dev)  # worker threads start on device 0
            while self.slvr.busy() and self.slvr._step():
                pass
            self.slvr._astate["active"].clear()
