"""
Primal-dual splitting solvers
(reference: src/pyxu/opt/solver/pds.py -- _PrimalDualSplitting:26, CondatVu:210, PD3O:523,
ChambollePock:867, LorisVerhoeven:970, DavisYin:1115, ForwardBackward:1698, ProximalPoint:1788).

Problem:  min_x f(x) + g(x) + h(Kx).  Constructor / fit() arguments, automatic step-size selection,
default stopping criterion and logged variables follow the reference.  What differs is how an
iteration executes: `m_step()` asks a planner (``_Plan``) which of three execution paths applies:

* ``fused``  -- h o K is a (scaled) L21 / L1 norm of a Gradient stack, g has a pointwise prox and f is
                 null / a shifted squared-l2 data term (or, for CondatVu, anything differentiable):
                 one iteration = ONE kernel (pxb_pds_iter: primal + dual half-steps + RelError norms in a
                 single sweep, ping-pong buffers) when the Gradient is a 2-/3-direction first-order scheme
                 with 'constant' boundaries; otherwise pxb_pds_primal + pxb_pds_dual (two sweeps).  PD3O's
                 x is then only materialised when something reads it (stopping criterion, log, solution).
* ``semi``   -- same h o K, but PD3O with a non-local f (e.g. a blur in the data term): K^T z and the
                 prox are separate passes, the dual half-step stays fused.
* ``generic``-- any other composition: the reference's formulas evaluated through the operators'
                 own methods, glued by fused elementwise kernels (pxb_lincomb / pxb_prox_lincomb /
                 pxb_dual_update).
"""
import ctypes as C
import math
import warnings

import numpy as np

from ... import _array as A
from ... import _cabi as K
from ... import _kernels as kr
from ...abc import operator as pxo
from ...abc.solver import Solver
from ...operator.linop.base import IdentityOp, NullFunc, NullOp
from .. import stop as pxs

__all__ = ["CondatVu", "CV", "PD3O", "ChambollePock", "CP", "LorisVerhoeven", "LV", "DavisYin", "DY",
           "ForwardBackward", "FB", "ProximalPoint", "PP"]


def _is_null(op):
    return getattr(op, "_name", "") == "NullFunc"


class _Plan:
    """Decides how one iteration is executed and owns the descriptors / work buffers."""

    def __init__(self, solver, algo, x0, part=None, defer_shift=False):
        """x0: the primal iterate on the device (a slab's planes under a z-slab decomposition; `part` = (volume shape,
        rank, world) then makes the planner keep only this rank's planes of per-voxel arrays).  defer_shift: a per-voxel
        shift of the data term that lives in host memory stays there (`fshift_host`; the streamed fit uploads it chunk by
        chunk together with x0) -- x0 may then be a `meta` tensor that only carries shape and dtype."""
        f, g, h, Kop = solver._f, solver._g, solver._h, solver._K
        self.algo = algo
        self.fshift_host = None
        self._defer_shift = bool(defer_shift)
        self._x0_like = x0
        self.kind = "generic"
        self.batch = max(1, x0.numel() // x0.shape[-1])
        self.gspec = g._prox_spec()
        self.hspec = None if _is_null(h) else h._dual_spec()
        self.h_null = _is_null(h)
        self.fkind, self.falpha, self.fshift = None, 0.0, None
        self.fstencil = None  # f = alpha*||A x + c||^2 with A a Stencil: (A, alpha, c on the device or None)
        if _is_null(f):
            self.fkind = K.F_NONE
        else:
            s = f._sql2_spec()
            if s is not None:
                self.fkind, self.falpha = K.F_SQL2, float(s[0])
                c = s[1]
                if self._defer_shift and isinstance(c, np.ndarray) and c.size == x0.numel():
                    self.fshift_host = A.host_flat(c, x0.dtype)
                else:
                    self.fshift = self._shift_on_device(c, x0, part)
                if self.fshift is not None and x0.numel() % self.fshift.numel() != 0:
                    self.fkind = None
            elif part is not None:
                from ...operator.linop.stencil import Stencil

                lhs, rhs = getattr(f, "_lhs", None), getattr(f, "_rhs", None)
                s = lhs._sql2_spec() if hasattr(lhs, "_sql2_spec") else None
                if s is not None and isinstance(rhs, Stencil):
                    self.fstencil = (rhs, float(s[0]), self._shift_on_device(s[1], x0, part))

        from ...operator.linop.diff import _DiffStack

        tv = isinstance(Kop, _DiffStack) and Kop._fusable and self.gspec is not None
        if tv and not self.h_null:
            hs = self.hspec
            N, nd = Kop.dim, len(Kop._dirs)
            ok_l21 = hs is not None and hs[0] == K.DUAL_L21 and tuple(hs[2]) == (1, nd, N)
            ok_l1 = hs is not None and hs[0] == K.DUAL_L1 and h.dim == nd * N
            tv = ok_l21 or ok_l1
        elif self.h_null:
            tv = False
        if tv:
            if self.fkind is not None:
                self.kind = "fused"
            elif algo == K.ALGO_CV:
                self.kind = "fused"
                self.fkind = K.F_GRADARR
            else:
                self.kind = "semi"
        self.K = Kop
        if self.kind != "generic":
            self.gdesc = Kop._desc(self.batch, A.dcode(x0))
        self._w = None
        self.iter_ok = None if self.kind == "fused" else False  # single-kernel iteration: None = not tried yet
        self.alt = None  # (primal, dual) spare buffers of the ping-pong

    @staticmethod
    def _shift_on_device(c, x0, part):
        """The data-term shift as a flat device tensor of x0's dtype (None: no shift; one sample: a scalar shift)."""
        if c is None:
            return None
        if isinstance(c, float):
            import torch

            return torch.full((1,), c, dtype=x0.dtype, device=A.current_device() if x0.device.type == "meta" else x0.device)
        if part is not None:
            from ... import slab

            n = int(c.numel()) if hasattr(c, "numel") else int(np.size(c))
            if isinstance(c, slab.ShardedArray) or n == int(np.prod(part[0])):
                c, _ = slab.local_part(c, *part)  # only this rank's planes cross the bus
        dev, _ = A.asdevice(c, dtype=x0.dtype)
        return dev.reshape(-1)

    @property
    def w(self):  # work array of the two-sweep form, allocated on first use
        if self._w is None:
            self._w = A.empty_like(self._x0_like)
        return self._w

    def params(self, mst, garr=None):
        # the block only changes when a step size does (or, for CondatVu with a non-local f, the grad f array moves): built once
        key = (mst["tau"], mst["sigma"], mst["rho"])
        hit = getattr(self, "_params_cache", None)
        if hit is not None and hit[0] == key:
            p = hit[1]
            if garr is not None:
                p.f.garr = garr.data_ptr()
            return p
        p = self._build_params(mst, garr)
        self._params_cache = (key, p)
        return p

    def _build_params(self, mst, garr=None):
        p = K.PdsParams()
        p.tau, p.sigma, p.rho = float(mst["tau"]), float(mst["sigma"]), float(mst["rho"])
        p.g = K.ProxSpec(int(self.gspec[0]), 0, float(self.gspec[1]), float(self.gspec[2]))
        fk = self.fkind if self.fkind is not None else K.F_NONE
        f = K.FTerm()
        f.kind, f.alpha = fk, self.falpha
        if fk == K.F_SQL2 and self.fshift is not None:
            f.shift, f.shift_period = self.fshift.data_ptr(), self.fshift.numel()
        if fk == K.F_GRADARR:
            f.garr = garr.data_ptr()
        p.f = f
        p.hkind, p.lam = int(self.hspec[0]), float(self.hspec[1])
        return p


class _PrimalDualSplitting(Solver):
    _ALGO = None
    _probe = None  # optional callable(tag) invoked around the fused kernels (bench.py records CUDA events with it)

    def __init__(self, f=None, g=None, h=None, K=None, beta=None, **kwargs):
        kwargs.update(log_var=kwargs.get("log_var", ("x", "z")))
        super().__init__(**kwargs)
        if (f is None) and (g is None) and (h is None):
            raise ValueError("Cannot minimize always-0 functional. At least one of Parameter[f, g, h] must be specified.")
        primal_dim = f.dim if f is not None else (g.dim if g is not None else h.dim)
        if h is not None:
            dual_dim = h.dim
        elif K is not None:
            dual_dim = K.shape[0]
        else:
            dual_dim = primal_dim
        self._f = NullFunc(dim=primal_dim) if (f is None) else f
        self._g = NullFunc(dim=primal_dim) if (g is None) else g
        self._h = NullFunc(dim=dual_dim) if (h is None) else h
        self._beta = self._set_beta(beta)
        if h is not None:
            self._K = IdentityOp(dim=h.dim) if (K is None) else K
        else:
            if K is None:
                K_dim = f.dim if f is not None else g.dim
                self._K = NullOp(shape=(K_dim, K_dim))
            else:
                raise ValueError("Optional argument ``h`` mut be specified if ``K`` is not None.")
        self._plan = None
        self._slab = None

    # ---------------------------------------------------------------------------------------
    def m_init(self, x0, z0=None, tau=None, sigma=None, rho=None, tuning_strategy=1, distributed=None):
        """Same arguments as the reference (pds.py:158-207, 723-745) plus `distributed`: under torch.distributed with more
        than one rank a 3-D problem with the fused TV structure is decomposed into z-slabs, one per rank (pyxu_b200.slab;
        None = automatic, True = required, False = every rank solves its own problem)."""
        mst = self._mstate
        self._slab = None
        self._tuning_strategy = int(tuning_strategy)
        gamma = self._set_gamma(tuning_strategy)
        mst["tau"], mst["sigma"], delta = self._set_step_sizes(tau, sigma, gamma)
        mst["rho"] = self._set_momentum_term(rho, delta)
        if self._m_init_slab(x0, z0, distributed):
            return
        x0d, origin = A.asdevice(x0)
        self._astate["origin"] = origin
        mst["x"] = x0d.clone() if origin == A.DEVICE else x0d  # never write into the caller's buffer (a host array was copied already)
        if z0 is None:
            mst["z"] = self._K(mst["x"])
            if mst["z"].data_ptr() == mst["x"].data_ptr():
                mst["z"] = mst["z"].clone()
        else:
            z0d, zo = A.asdevice(z0, dtype=x0d.dtype)
            mst["z"] = z0d.clone() if zo == A.DEVICE else z0d
        self._plan = _Plan(self, self._ALGO, mst["x"])
        self._setup_fused_norms()

    # -- z-slab decomposition over the ranks of torch.distributed (pyxu_b200.slab) -------------------------------
    def _m_init_slab(self, x0, z0, distributed):
        """Hands the iteration to a slab engine when the problem decomposes; returns False for a single-domain solve."""
        from ... import slab
        from ...operator.linop.diff import _DiffStack

        sharded_in = isinstance(x0, slab.ShardedArray)
        ctx = slab.context(True if (sharded_in and distributed is None) else distributed)
        self._stream_src = None
        so = getattr(self, "_stream_out", None)  # a result an earlier streamed fit() of this object brought back and nobody collected
        if so is not None and not isinstance(so[1], np.ndarray):
            A._give_back(so[1].numel(), so[1])
        self._stream_out = None
        if ctx is None:
            return self._m_init_streamed(x0, z0)
        required = distributed is True or sharded_in
        Kop = self._K
        n_in = int(x0.numel()) if hasattr(x0, "numel") else int(np.size(x0))
        ok = isinstance(Kop, _DiffStack) and len(Kop.arg_shape) == 3 and Kop._dirs == (0, 1, 2) and n_in == Kop.dim
        if ok:
            vol = Kop.arg_shape
            ok = vol[0] >= 3 * ctx[1]
        if not ok:
            if required:
                raise NotImplementedError("z-slab decomposition needs K = Gradient of a 3-D volume (all three directions), one volume per solve, "
                                          "and at least 3 planes per rank")
            return False
        mst, ast = self._mstate, self._astate
        rank, world = ctx
        part = (vol, rank, world)
        x_loc, _ = slab.local_part(x0, *part)
        x0d, origin = A.asdevice(x_loc)
        x0d = x0d.reshape(-1)
        plan = _Plan(self, self._ALGO, x0d, part=part)
        stencil_term = plan.kind == "fused" and plan.fkind == K.F_GRADARR and plan.fstencil is not None
        if plan.kind != "fused" or (plan.fkind == K.F_GRADARR and not stencil_term):
            if required:
                raise NotImplementedError("this problem has no z-slab form: it needs h o K = (L21 | L1) o Gradient, a pointwise prox for g and a data term "
                                          "that is null, alpha*||x + c||^2, or (CondatVu) alpha*||A x + c||^2 with A a separable Stencil")
            return False
        ast["origin"] = origin
        ast["sharded"] = sharded_in
        z0d = None
        if z0 is not None:
            z_loc, _ = slab.local_part(z0, *part, comps=3)
            z0d, _ = A.asdevice(z_loc, dtype=x0d.dtype)
        self._plan = plan
        mst["x"] = mst["z"] = None  # live in the engine's slab buffers; materialised by _logged()
        crit = ast["stop_crit"]
        reads = crit._state_vars() if crit is not None else None
        self._setup_fused_norms(device=x0d.device)
        served = frozenset(k for k in ("x", "z") if k in mst.get("_fused_norms", {}))
        if ast["track_objective"] or reads is None or not (reads <= served) or crit._rank_local():
            raise NotImplementedError("z-slab decomposition: the stopping criterion must be built from MaxIter / ManualStop / RelError (2-norm, no f) "
                                      "on x or z, evaluated every iteration (stop_rate = 1); objective tracking is not available")
        params = plan._build_params(mst, garr=x0d)  # (garr is re-pointed by the engine)
        if stencil_term:
            Aop, alpha, c = plan.fstencil
            self._slab = slab.SlabDeblurCV(Kop, Aop, alpha, params, x0d, z0d, c, rank, world)
        else:
            self._slab = slab.SlabTV(self._ALGO, Kop, params, x0d, z0d, plan.fshift, rank, world)
        plan.iter_ok = True
        return True

    # -- single GPU, large HOST arrays: iterations as a wavefront behind the upload (slab.SlabTV.run_streamed) ----------
    _STREAM_MIN_BYTES = 256 << 20  # smaller volumes are uploaded in one go (PYXU_B200_STREAM_MIN_BYTES overrides; 0 = never stream)
    _STREAM_EPOCH = 32             # iterations run speculatively behind the upload before the criterion's sums are looked at
    _STREAM_PLANES = 16

    def _m_init_streamed(self, x0, z0):
        """fit(x0=<large host array>) in BLOCK mode on a fused 3-D TV problem: nothing is uploaded here; _fit_run() streams x0
        (and a per-voxel data-term shift that also lives in host memory) in z-chunks with the first iterations queued behind
        the chunks (the host->device transfer is what bounds a 1024^3 solve end to end).  False: the ordinary path."""
        import os

        import torch

        from ... import slab
        from ...abc.solver import Mode
        from ...operator.linop.diff import _DiffStack

        mst, ast, Kop = self._mstate, self._astate, self._K
        floor = int(os.environ.get("PYXU_B200_STREAM_MIN_BYTES", self._STREAM_MIN_BYTES))
        if floor <= 0 or z0 is not None or ast["mode"] is not Mode.BLOCK or not isinstance(x0, np.ndarray):
            return False
        if x0.dtype not in (np.float32, np.float64) or x0.nbytes < floor:
            return False
        if not (isinstance(Kop, _DiffStack) and len(Kop.arg_shape) == 3 and Kop._dirs == (0, 1, 2) and x0.size == Kop.dim):
            return False
        vol = Kop.arg_shape
        if tuple(Kop._mode)[0] != "constant":  # a z-chunk's cut faces are open sides: exact as long as no fold reaches along axis 0
            return False
        if vol[0] < 2 * self._STREAM_PLANES or ast["track_objective"] or ast["wb_rate"] is not None or ast["stop_rate"] != 1:
            return False
        tdt = A.torch_dtype(x0.dtype)
        meta = torch.empty(x0.size, dtype=tdt, device="meta")
        plan = _Plan(self, self._ALGO, meta, part=(vol, 0, 1), defer_shift=True)
        if plan.kind != "fused" or plan.fkind == K.F_GRADARR:
            return False
        crit = ast["stop_crit"]
        reads = crit._state_vars()
        self._plan = plan
        mst["x"] = mst["z"] = None
        self._setup_fused_norms(device=A.current_device())
        served = frozenset(k for k in ("x", "z") if k in mst.get("_fused_norms", {}))
        if reads is None or not (reads <= served) or crit._rank_local():
            self._plan = None
            mst.pop("_fused_norms", None)
            return False
        ast["origin"], ast["sharded"] = A.HOST, False
        params = plan._build_params(mst)
        self._slab = slab.SlabTV(self._ALGO, Kop, params, None, None, plan.fshift, 0, 1, dtype=tdt, shift_streamed=plan.fshift_host is not None)
        self._stream_src = (A.host_flat(x0, tdt), plan.fshift_host)
        self._stream_out = None
        plan.iter_ok = True
        return True

    def _run_streamed(self):
        """The solver loop (abc/solver.py:_step) with the first epoch of iterations queued behind the upload.  The criterion is
        evaluated afterwards from the sums every iteration left behind, exactly as the loop would have: history, log and the
        iteration at which it stops are those of the one-iteration-at-a-time loop.  Should the criterion turn out to have been met
        INSIDE the epoch (the wave had already moved on), the solve is redone from the host arrays up to that iteration."""
        import copy

        import torch

        ast, mst, eng = self._astate, self._mstate, self._slab
        crit = ast["stop_crit"]
        x0_host, shift_host = self._stream_src
        self._stream_src = None
        eng.probe = self._probe
        planes = self._STREAM_PLANES
        if not self._pre_step():  # met before the first iteration: upload only, so that solution() / stats() have their arrays
            eng.run_streamed(x0_host, shift_host, 0, planes=planes)
            return
        budget = crit._budget()  # (inf without a MaxIter leaf)
        E = int(max(1, min(budget, self._STREAM_EPOCH)))
        final = E == budget  # a MaxIter leaf fires at the test after iteration E: the result can travel back behind the wave
        use_x, use_z = self._nx is not None, self._nz is not None
        nrm = torch.zeros((E, 2, 1, 2), dtype=torch.float64, device=eng.dev) if (use_x or use_z) else None
        out = None
        if final and "x" in ast["log_var"]:
            buf = A.take_reserved(eng.local_voxels * eng._pb[0].element_size())
            out = buf.view(eng.dtype) if buf is not None else None
        done = eng.run_streamed(x0_host, shift_host, E, nrm, use_x, use_z, out, planes=planes)
        if done == 0:  # outside the single-kernel envelope: everything is on the device, iterate the ordinary way
            if out is not None:
                A._give_back(buf.numel(), buf)
            ast["idx"] += 1  # (the test that precedes the first iteration was made above)
            self.m_step()
            self._loop_steps()
            return
        idx0 = ast["idx"]
        fused = mst.get("_fused_norms")
        if nrm is None:  # MaxIter / ManualStop only: nothing to look at
            sums = np.empty((E, 0))
            stop_at = None
        else:
            sums = nrm.cpu().numpy()  # (synchronises: the wave has left the last chunk)
            dec, _ = copy.deepcopy(crit)._replay(sums[: E - 1]) if E > 1 else (np.zeros(0, bool), None)
            hit = np.flatnonzero(dec)
            stop_at = int(hit[0]) if hit.size else None  # the test after iteration stop_at + 1 fires
        if stop_at is not None:
            if out is not None:
                torch.cuda.synchronize()
                A._give_back(buf.numel(), buf)
                out = None
            E = stop_at + 1
            eng.reset_streamed()
            eng.run_streamed(x0_host, shift_host, E, None, False, False, None, planes=planes)  # no sums: the recorded ones stand
        if E > 1:
            _, info = crit._replay(sums[: E - 1])
            self._record_block(idx0 + 1, info)
        ast["idx"] = idx0 + E
        if fused is not None:
            fused["_host"], fused["_stamp"] = sums[E - 1], 0  # the test at the top of the loop reads the last iteration's sums
        if out is not None:
            self._stream_out = (ast["idx"], buf)
        self._loop_steps()

    def _loop_steps(self):
        while self._step():
            pass

    def _step_slab(self):
        self._zero_norms()
        eng = self._slab
        want_x = self._nx is not None or self._x_every
        eng.probe = self._probe
        eng.step(self._nx, self._nz, want_x)
        if self._nrm is not None and eng.world > 1:
            import torch.distributed as dist

            dist.all_reduce(self._nrm, group=eng.group)  # the single fused scalar all-reduce of the stopping criterion

    def _logged(self, k):
        if getattr(self, "_slab", None) is None:
            return super()._logged(k)
        if k not in self._astate["log_var"] or k not in ("x", "z"):
            return None
        from ... import slab

        eng = self._slab
        so = getattr(self, "_stream_out", None)
        if k == "x" and so is not None and not getattr(self, "_in_writeback", False):  # the streamed fit has already brought x back behind the wave
            self._stream_out = None
            if so[0] == self._astate["idx"]:
                A.synchronize()
                return A.as_result(so[1], eng.dtype, (eng.local_voxels,))
            A._give_back(so[1].numel(), so[1])
        loc = eng.x_local().unsqueeze(0) if k == "x" else eng.z_local()
        if self._astate.get("sharded"):
            out = A.restore(loc.contiguous().reshape(-1), self._astate["origin"])
            return slab.ShardedArray(out, eng.shape, comps=loc.shape[0], rank=eng.rank, world=eng.world)
        full = slab.gather_planes(loc, eng.shape, eng.world, group=eng.group) if eng.world > 1 else loc.contiguous()
        return A.restore(full.reshape(-1), self._astate["origin"])

    def writeback(self):
        self._in_writeback = True  # (the checkpoint copies x out of the device; the streamed result stays for solution())
        try:
            super().writeback()
        finally:
            self._in_writeback = False

    def _setup_fused_norms(self, device=None):
        """If the stopping criterion is RelError on x / z evaluated every iteration, let the update kernels
        accumulate its norms (no extra pass, no x_prev copy).  Also decides whether PD3O's x must be written by
        every iteration or only when somebody asks for it (`_materialize`)."""
        import torch

        mst, ast = self._mstate, self._astate
        device = mst["x"].device if device is None else device
        crit = ast["stop_crit"]
        want = crit._fused_vars() if crit is not None else frozenset()
        self._nx = self._nz = self._nrm = None
        self._x_stale = False
        reads = crit._state_vars() if crit is not None else None
        fusable = self._plan.kind != "generic" and ast["stop_rate"] == 1 and want <= {"x", "z"}
        if fusable:
            if self._plan.kind == "semi":
                want = want - {"x"}
            if self._plan.h_null:
                want = want - {"z"}
            rows = self._plan.batch
            if want:
                # one (2, rows, 2) buffer: a single memset per iteration and a single 32-byte readback
                self._nrm = torch.zeros((2, rows, 2), dtype=torch.float64, device=device)
                fused = {"_all": self._nrm, "_host": None, "_stamp": -1}
                if "x" in want:
                    self._nx = fused["x"] = self._nrm[0]
                if "z" in want:
                    self._nz = fused["z"] = self._nrm[1]
                mst["_fused_norms"] = fused
        else:
            want = frozenset()
        # x is read behind the solver's back by: an unknown criterion, a criterion on "x" that is not served by the
        # fused norms, or objective tracking
        self._x_every = bool(ast["track_objective"]) or reads is None or ("x" in reads and "x" not in want)

    def _zero_norms(self):
        if self._nrm is not None:
            self._nrm.zero_()
            self._mstate["_fused_norms"]["_stamp"] = -1

    def _iter_fused(self, algo, garr=None):
        """One iteration as a single kernel (pxb_pds_iter).  Returns False when the problem is outside the kernel's
        envelope (decided once, on the first call: nothing was launched) -- the caller then runs the two-sweep form."""
        mst, pl = self._mstate, self._plan
        key = "u" if algo == K.ALGO_PD3O else "x"
        if pl.alt is None:
            pl.alt = (A.empty_like(mst[key]), A.empty_like(mst["z"]))
        want_x = algo == K.ALGO_PD3O and (self._nx is not None or self._x_every)
        self._zero_norms()
        p = pl.params(mst, garr=garr)
        if self._probe:
            self._probe("iter_begin")
        rc = K.lib().pxb_pds_iter(algo, C.byref(pl.gdesc), C.byref(p), A.ptr(mst[key]), A.ptr(mst["z"]), A.ptr(pl.alt[0]), A.ptr(pl.alt[1]),
                                  A.ptr(mst["x"]) if want_x else None, A.ptr(self._nx), A.ptr(self._nz), A.stream())
        if rc == -3 and pl.iter_ok is None:  # PXB_ENOSUP
            pl.iter_ok, pl.alt = False, None
            return False
        K.check(rc, "pxb_pds_iter")
        if self._probe:
            self._probe("iter_end")
        pl.iter_ok = True
        new_primal, new_dual = pl.alt
        pl.alt = (mst[key], mst["z"])  # the previous iterate: next call's output buffers (and what a lazy x is rebuilt from)
        mst[key], mst["z"] = new_primal, new_dual
        if algo == K.ALGO_PD3O and not want_x:
            self._x_stale = True
        return True

    # -- iterations queued back to back, the stopping rule tested on the device (pxb_pds_iter_n) ---------------------------
    _BATCH_MAX = 256       # iterations per native call, at most
    _BATCH_SECONDS = 0.25  # ... and no more than about this much device time: the log / history lag the device by one batch

    def _batch_rule(self):
        """The device-side form of the stopping criterion (K.StopRule) when this solve can run with the host out of the loop:
        BLOCK mode, single-kernel iteration with a pointwise data term, RelError sums fused, a criterion made of MaxIter /
        ManualStop / RelError leaves, nothing that needs the host between iterations (objective tracking, checkpoints)."""
        ast, pl = self._astate, self._plan
        from ...abc.solver import Mode

        if ast["mode"] is not Mode.BLOCK or self._slab is not None or pl is None or pl.kind != "fused" or pl.fkind == K.F_GRADARR:
            return None
        if ast["track_objective"] or ast["wb_rate"] is not None or ast["stop_rate"] != 1 or self._x_every:
            return None
        crit = ast["stop_crit"]
        try:
            table = sum(int(bool(crit._device_eval(bool(px), bool(pz)))) << (2 * px + pz) for px in (0, 1) for pz in (0, 1))
        except NotImplementedError:
            return None
        leaves = {}
        for var, eps, every in crit._device_leaves():
            if leaves.setdefault(var, (eps, every)) != (eps, every):
                return None  # two different tests on the same variable
        if not leaves and self._nrm is None:
            return False if table == 0 else None  # nothing for the device to test (MaxIter / ManualStop only): plain batches
        if not leaves or not set(leaves) <= {"x", "z"} or ("x" in leaves) != (self._nx is not None) or ("z" in leaves) != (self._nz is not None):
            return None
        r = K.StopRule()
        r.eps_x, r.all_x = leaves.get("x", (0.0, True))
        r.eps_z, r.all_z = leaves.get("z", (0.0, True))
        r.table = table
        return r

    def _fit_run(self):
        if getattr(self, "_stream_src", None) is not None:
            try:
                self._run_streamed()
            except Exception as e:
                self._on_error(e)
            self._astate.update(mode=None, active=None, worker=None)
            self._cleanup_logger()
            return
        rule = self._batch_rule() if self._plan is not None else None
        if rule is None:  # (False: batches without a device-side rule)
            return super()._fit_run()
        try:
            self._run_batched(rule)
        except Exception as e:
            self._on_error(e)
        self._astate.update(mode=None, active=None, worker=None)
        self._cleanup_logger()

    def _run_batched(self, rule):
        import time

        import torch

        ast, mst, pl = self._astate, self._mstate, self._plan
        crit = ast["stop_crit"]
        # the first iteration goes through m_step(): it decides whether pxb_pds_iter serves this problem at all
        if not self._pre_step():
            return
        ast["idx"] += 1
        self.m_step()
        if pl.iter_ok is not True:
            while self._step():
                pass
            return
        algo = self._ALGO
        key = "u" if algo == K.ALGO_PD3O else "x"
        rows = pl.batch
        cap, per_iter = 8, None
        p = pl.params(mst)
        if rule is False:  # the device has nothing to test: plain batches, the host counts
            while True:
                if not self._pre_step():
                    return
                n = int(min(cap, self._BATCH_MAX, crit._budget()))
                if n < 2:
                    ast["idx"] += 1
                    self.m_step()
                    continue
                if pl.alt is None:
                    pl.alt = (A.empty_like(mst[key]), A.empty_like(mst["z"]))
                t0 = time.perf_counter()
                rc = K.lib().pxb_pds_iter_n(algo, C.byref(pl.gdesc), C.byref(p), A.ptr(mst[key]), A.ptr(mst["z"]), A.ptr(pl.alt[0]), A.ptr(pl.alt[1]),
                                            None, None, n, None, None, A.stream())
                K.check(rc, "pxb_pds_iter_n")
                if n % 2:
                    cur = pl.alt
                    pl.alt = (mst[key], mst["z"])
                    mst[key], mst["z"] = cur
                if algo == K.ALGO_PD3O:
                    self._x_stale = True
                idx0 = ast["idx"]
                ast["idx"] = idx0 + n
                decisions, info = crit._replay(np.empty((n - 1, 0)))
                if bool(np.any(decisions)):
                    raise RuntimeError("a batch of iterations went past the one at which the criterion stops")
                self._record_block(idx0 + 1, info)
                if cap < self._BATCH_MAX:  # (launches are asynchronous: time a batch only while the batch size is still growing)
                    A.synchronize()
                    dt = (time.perf_counter() - t0) / n
                    per_iter = dt if per_iter is None else min(per_iter, dt)
                    cap = max(2, min(self._BATCH_MAX, int(self._BATCH_SECONDS / max(per_iter, 1e-7))))  # (one timed batch is enough: no doubling)
        buf = torch.empty(self._BATCH_MAX * 4 * rows + 2, dtype=torch.float64, device=mst["z"].device)  # sums of every iteration | ctl
        ctl_view = buf[-2:].view(torch.int32)
        fused = mst["_fused_norms"]
        while True:
            if not self._pre_step():  # the test that follows the last iteration carried out (host side, as in the reference's loop)
                return
            n = int(min(cap, self._BATCH_MAX, crit._budget()))
            if n < 2:
                ast["idx"] += 1
                self.m_step()
                continue
            if pl.alt is None:
                pl.alt = (A.empty_like(mst[key]), A.empty_like(mst["z"]))
            want_x = algo == K.ALGO_PD3O and self._nx is not None
            buf.zero_()
            t0 = time.perf_counter()
            rc = K.lib().pxb_pds_iter_n(algo, C.byref(pl.gdesc), C.byref(p), A.ptr(mst[key]), A.ptr(mst["z"]), A.ptr(pl.alt[0]), A.ptr(pl.alt[1]),
                                        A.ptr(mst["x"]) if want_x else None, A.ptr(buf), n, C.byref(rule), A.ptr(ctl_view), A.stream())
            K.check(rc, "pxb_pds_iter_n")
            host = buf.cpu()  # ONE readback per batch: the sums of its iterations and the control block
            done = int(host[-2:].view(torch.int32)[1])
            dt = time.perf_counter() - t0
            assert 1 <= done <= n, (done, n)
            sums = host[: done * 4 * rows].numpy().reshape(done, 2, rows, 2)
            if done % 2:  # the iterate is in the other pair
                cur = pl.alt
                pl.alt = (mst[key], mst["z"])
                mst[key], mst["z"] = cur
            if algo == K.ALGO_PD3O and not want_x:
                self._x_stale = True
            idx0 = ast["idx"]
            ast["idx"] = idx0 + done
            if done > 1:  # the tests between the iterations of the batch: none of them stopped it
                decisions, info = crit._replay(sums[: done - 1])
                if bool(np.any(decisions)):
                    raise RuntimeError("pxb_pds_iter_n went past an iteration at which the host-side criterion stops")
                self._record_block(idx0 + 1, info)
            # the sums of the last iteration wait for the test at the top of the loop
            fused["_host"], fused["_stamp"] = sums[done - 1], 0
            per_iter = dt / done if per_iter is None else min(per_iter, dt / done)
            cap = max(2, min(self._BATCH_MAX, int(self._BATCH_SECONDS / max(per_iter, 1e-7))))

    def _materialize(self, name):
        """x of PD3O is not written by the single-kernel iteration unless something needs it every step; rebuild it
        on demand from the previous iterate: x_k = prox_{tau g}(u_{k-1} - tau K^T z_{k-1})  (pds.py:747-750)."""
        mst = self._mstate
        eng = getattr(self, "_slab", None)
        if eng is not None and name in ("x", "z"):  # the iterate lives in the slab engine's buffers
            if eng.world > 1:
                raise NotImplementedError("under a z-slab decomposition the iterate is distributed: use stats() / solution() (they gather it)")
            return eng.x_local().reshape(-1) if name == "x" else eng.z_local().reshape(-1)
        if name == "x" and getattr(self, "_x_stale", False):
            u_prev, z_prev = self._plan.alt
            ktz = self._K.jacobian(u_prev).adjoint(z_prev)
            mst["x"] = self._prox_g(1.0, u_prev, -mst["tau"], ktz, out=mst["x"])
            self._x_stale = False
        return mst.get(name)

    def default_stop_crit(self):
        stop_crit_x = pxs.RelError(eps=1e-4, var="x", f=None, norm=2, satisfy_all=True)
        stop_crit_z = pxs.RelError(eps=1e-4, var="z", f=None, norm=2, satisfy_all=True)
        return stop_crit_x & stop_crit_z if self._h._name != "NullFunc" else stop_crit_x

    def solution(self, which="primal"):
        logged = self._astate["log_var"]
        if which == "primal":
            assert "x" in logged, "Primal variable x was not logged (declare it in log_var to log it)."
        elif which == "dual":
            assert "z" in logged, "Dual variable z was not logged (declare it in log_var to log it)."
        else:
            raise ValueError(f"Parameter which must be one of ['primal', 'dual'] got: {which}.")
        return self._logged("x" if which == "primal" else "z")  # only the requested variable leaves the device

    def objective_func(self):
        x = self._materialize("x")
        out = self._f(x) + self._g(x)
        if not _is_null(self._h):
            out = out + self._h(self._K(x))
        return out

    def _set_beta(self, beta):
        if beta is None:
            dl = self._f.diff_lipschitz
            if math.isfinite(dl):
                return float(dl)
            raise ValueError("beta: automatic inference not supported for operators with unbounded Lipschitz gradients.")
        return float(beta)

    def _set_gamma(self, tuning_strategy):
        return float(self._beta) if tuning_strategy != 2 else float(self._beta / 1.9)

    def _set_step_sizes(self, tau, sigma, gamma):
        raise NotImplementedError

    def _set_momentum_term(self, rho, delta):
        if rho is None:
            rho = 1.0 if self._tuning_strategy != 3 else delta - 0.1
        else:
            assert rho <= delta, f"Parameter rho must be smaller than delta: {rho} > {delta}."
        return float(rho)

    def _K_lipschitz(self):
        if math.isfinite(self._K.lipschitz):
            return self._K.lipschitz
        raise ValueError("Please compute the Lipschitz constant of the linear operator K by calling its method 'estimate_lipschitz()'")

    def _check_K_linear(self):
        if not isinstance(self._K, pxo.LinOp):
            raise ValueError("Automatic selection of parameters is only supported in the case in which K is a linear operator. "
                             f"Got operator of type {self._K.__class__}.")

    # -- shared kernels ---------------------------------------------------------------------
    def _dual_fused(self, w):
        mst, pl = self._mstate, self._plan
        p = pl.params(mst, garr=w)  # garr unused by the dual kernel
        rc = K.lib().pxb_pds_dual(C.byref(pl.gdesc), C.byref(p), A.ptr(w), A.ptr(mst["z"]), A.ptr(self._nz), A.stream())
        K.check(rc, "pxb_pds_dual")

    def _dual_generic(self, w):
        """z <- (1-rho) z + rho prox_{sigma h*}(z + sigma K w) through operator methods."""
        mst, pl = self._mstate, self._plan
        t = self._K(w)
        hs = pl.hspec
        if hs is not None and hs[0] == K.DUAL_L21:
            outer, group, inner = hs[2]
            kr.dual_update(K.DUAL_L21, mst["z"], t, pl.batch * outer, group, inner, hs[1], mst["sigma"], mst["rho"])
        elif hs is not None and hs[0] == K.DUAL_L1:
            kr.dual_update(K.DUAL_L1, mst["z"], t, pl.batch, 1, self._h.dim, hs[1], mst["sigma"], mst["rho"])
        else:
            p = kr.lincomb(1.0, mst["z"], mst["sigma"], t, out=t)
            z_temp = self._h.fenchel_prox(p, sigma=mst["sigma"])
            kr.lincomb(1.0 - mst["rho"], mst["z"], mst["rho"], z_temp, out=mst["z"])

    def _prox_g(self, a, x, b=0.0, y=None, c=0.0, z=None, out=None):
        """prox_{tau g}(a x + b y + c z): one pass when g is pointwise, else lincomb + g.prox."""
        mst, pl = self._mstate, self._plan
        if pl.gspec is not None:
            return kr.prox_lincomb(pl.gspec, mst["tau"], a, x, b, y, c, z, out=out)
        return self._g.prox(kr.lincomb(a, x, b, y, c, z), tau=mst["tau"])


_PDS = _PrimalDualSplitting


class CondatVu(_PrimalDualSplitting):
    r"""Condat-Vu primal-dual splitting (reference: pds.py:210-517, iteration pds.py:429-442)."""

    _ALGO = K.ALGO_CV

    def m_step(self):
        if self._slab is not None:
            return self._step_slab()
        mst, pl = self._mstate, self._plan
        if pl.kind == "fused":
            garr = self._f.grad(mst["x"]) if pl.fkind == K.F_GRADARR else None
            if pl.iter_ok is not False and self._iter_fused(K.ALGO_CV, garr=garr):
                return
            self._zero_norms()
            p = pl.params(mst, garr=garr)
            rc = K.lib().pxb_pds_primal(K.ALGO_CV, C.byref(pl.gdesc), C.byref(p), A.ptr(mst["x"]), A.ptr(mst["z"]), None, None,
                                        A.ptr(pl.w), A.ptr(self._nx), A.stream())
            K.check(rc, "pxb_pds_primal")
            self._dual_fused(pl.w)
            return
        # generic: x_temp = prox_g(x - tau grad f(x) - tau K^T z)
        x = mst["x"]
        gf = None if _is_null(self._f) else self._f.grad(x)
        ktz = None if pl.h_null else self._K.jacobian(x).adjoint(mst["z"])
        x_temp = self._prox_g(1.0, x, -mst["tau"], gf, -mst["tau"], ktz)
        if not pl.h_null:
            u = kr.lincomb(2.0, x_temp, -1.0, x)
            self._dual_generic(u)
        mst["x"] = kr.lincomb(mst["rho"], x_temp, 1.0 - mst["rho"], x, out=x_temp)

    def _set_step_sizes(self, tau, sigma, gamma):
        self._check_K_linear()
        tau = None if tau == 0 else tau
        sigma = None if sigma == 0 else sigma
        h_null = _is_null(self._h)
        if (tau is not None) and (sigma is None):
            assert tau > 0, f"Parameter tau must be positive, got {tau}."
            if h_null:
                assert tau <= 1 / gamma, f"Parameter tau must be smaller than 1/gamma: {tau} > {1 / gamma}."
                sigma = 0
            else:
                sigma = ((1 / tau) - gamma) * (1 / self._K_lipschitz() ** 2)
        elif (tau is None) and (sigma is not None):
            assert sigma > 0
            tau = 1 / gamma if h_null else 1 / (gamma + (sigma * self._K_lipschitz() ** 2))
        elif (tau is None) and (sigma is None):
            if self._beta > 0:
                if h_null:
                    tau, sigma = 1 / gamma, 0
                else:
                    L = self._K_lipschitz()
                    tau = sigma = (1 / L**2) * ((-gamma / 2) + math.sqrt((gamma**2 / 4) + L**2))
            else:
                if h_null:
                    tau, sigma = 1, 0
                else:
                    tau = sigma = 1 / self._K_lipschitz()
        delta = 2 if (self._beta == 0 or (isinstance(self._f, pxo.QuadraticFunc) and gamma <= self._beta)) else 2 - self._beta / (2 * gamma)
        return float(tau), float(sigma), float(delta)


CV = CondatVu


class PD3O(_PrimalDualSplitting):
    r"""Primal-dual three-operator splitting (reference: pds.py:523-864, iteration pds.py:747-761)."""

    _ALGO = K.ALGO_PD3O

    def m_init(self, x0, z0=None, tau=None, sigma=None, rho=None, tuning_strategy=1, distributed=None):
        super().m_init(x0=x0, z0=z0, tau=tau, sigma=sigma, rho=rho, tuning_strategy=tuning_strategy, distributed=distributed)
        mst = self._mstate
        if self._slab is not None:  # (g = h = 0 never takes the slab path: u0 = x0 there)
            return
        # if x0 == u0 the first step would not move x when g = h = 0 (reference: pds.py:741-745)
        if _is_null(self._g) and _is_null(self._h):
            mst["u"] = kr.lincomb(1.01, mst["x"])
        else:
            mst["u"] = mst["x"].clone()

    def m_step(self):
        if self._slab is not None:
            return self._step_slab()
        mst, pl = self._mstate, self._plan
        tau, rho = mst["tau"], mst["rho"]
        if pl.kind == "fused":
            if pl.iter_ok is not False and self._iter_fused(K.ALGO_PD3O):
                return
            self._zero_norms()
            p = pl.params(mst)
            if self._probe:
                self._probe("primal_begin")
            rc = K.lib().pxb_pds_primal(K.ALGO_PD3O, C.byref(pl.gdesc), C.byref(p), A.ptr(mst["u"]), A.ptr(mst["z"]), None,
                                        A.ptr(mst["x"]), A.ptr(pl.w), A.ptr(self._nx), A.stream())
            K.check(rc, "pxb_pds_primal")
            if self._probe:
                self._probe("primal_end")
            self._dual_fused(pl.w)
            if self._probe:
                self._probe("dual_end")
            return
        u = mst["u"]
        ktz = None if pl.h_null else self._K.jacobian(u).adjoint(mst["z"])
        x = self._prox_g(1.0, u, -tau, ktz, out=mst["x"] if pl.gspec is not None else None)
        mst["x"] = x
        gf = None if _is_null(self._f) else self._f.grad(x)
        if not pl.h_null:
            # w = x + u_temp - u = 2x - tau grad f(x) - u
            w = kr.lincomb(2.0, x, -tau, gf, -1.0, u, out=getattr(pl, "w", None))
            if pl.kind == "semi":
                self._zero_norms()
                self._dual_fused(w)
            else:
                self._dual_generic(w)
        # u <- (1-rho) u + rho (x - tau grad f(x))
        kr.lincomb(1.0 - rho, u, rho, x, -rho * tau, gf, out=u)

    def _set_step_sizes(self, tau, sigma, gamma):
        self._check_K_linear()
        tau = None if tau == 0 else tau
        sigma = None if sigma == 0 else sigma
        h_null = _is_null(self._h)
        if (tau is not None) and (sigma is None):
            assert 0 < tau <= 1 / gamma, "tau must be positive and smaller than 1/gamma."
            sigma = 0 if h_null else 1 / (tau * self._K_lipschitz() ** 2)
        elif (tau is None) and (sigma is not None):
            assert sigma > 0, f"sigma must be positive, got {sigma}."
            tau = 1 / gamma if h_null else min(1 / (sigma * self._K_lipschitz() ** 2), 1 / gamma)
        elif (tau is None) and (sigma is None):
            if self._beta > 0:
                if h_null:
                    tau, sigma = 1 / gamma, 0
                else:
                    self._K_lipschitz()
                    tau, sigma = self._optimize_step_sizes(gamma)
            else:
                if h_null:
                    tau, sigma = 1, 0
                else:
                    tau = sigma = 1 / self._K_lipschitz()
        delta = 2 if self._beta == 0 else 2 - self._beta * tau / 2
        return float(tau), float(sigma), float(delta)

    def _optimize_step_sizes(self, gamma):
        """The reference solves a linear program with scipy's HiGHS (pds.py:831-864):
            maximise log tau + log sigma   s.t.  log tau + log sigma <= log 0.99 - 2 log ||K||,  log tau <= log(1/gamma),  tau = sigma.
        With the equality it has one unknown and the unique solution log tau = min(b0 / 2, b1), which HiGHS' presolve
        returns exactly (tests/test_host_logic_cpu.py checks bit equality against scipy on random instances): no solver call,
        and scipy.optimize is not imported on the first fit()."""
        b0 = np.log(0.99) - 2 * np.log(self._K.lipschitz)
        b1 = np.log(1 / gamma)
        t = np.exp(min(b0 / 2, b1))
        return np.array([t, t])


def ChambollePock(g=None, h=None, K=None, base=CondatVu, **kwargs):
    """Chambolle-Pock / PDHG: f = 0 in CondatVu or PD3O (reference: pds.py:867-964)."""
    kwargs.update(log_var=kwargs.get("log_var", ("x", "z")))
    obj = base(f=None, g=g, h=h, K=K, beta=0, **kwargs)
    obj.__repr__ = lambda _: "ChambollePock"
    return obj


CP = ChambollePock


class LorisVerhoeven(PD3O):
    """min f(x) + h(Kx): PD3O with g = 0 (reference: pds.py:970-1109)."""

    def __init__(self, f=None, h=None, K=None, beta=None, **kwargs):
        kwargs.update(log_var=kwargs.get("log_var", ("x", "z")))
        super().__init__(f=f, g=None, h=h, K=K, beta=beta, **kwargs)


LV = LorisVerhoeven


class DavisYin(PD3O):
    """min f(x) + g(x) + h(x): PD3O with K = Id (reference: pds.py:1115-1237)."""

    def __init__(self, f, g=None, h=None, beta=None, **kwargs):
        kwargs.update(log_var=kwargs.get("log_var", ("x", "z")))
        super().__init__(f=f, g=g, h=h, K=None, beta=beta, **kwargs)


DY = DavisYin


def ForwardBackward(f=None, g=None, beta=None, **kwargs):
    """Forward-backward splitting: CondatVu with h = 0 (reference: pds.py:1698-1782)."""
    kwargs.update(log_var=kwargs.get("log_var", ("x",)))
    obj = CondatVu(f=f, g=g, h=None, K=None, beta=beta, **kwargs)
    obj.__repr__ = lambda _: "ForwardBackward"
    return obj


FB = ForwardBackward


def ProximalPoint(g=None, base=CondatVu, **kwargs):
    """Proximal-point method: f = h = 0 (reference: pds.py:1788-1862)."""
    kwargs.update(log_var=kwargs.get("log_var", ("x",)))
    obj = base(f=None, g=g, h=None, K=None, beta=None if False else 0, **kwargs)
    obj.__repr__ = lambda _: "ProximalPoint"
    return obj


PP = ProximalPoint
