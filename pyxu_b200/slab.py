"""
z-slab domain decomposition of large 3-D volumes across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The volume is cut along axis 0
into `world` contiguous slabs; every rank stores its slab with `halo` ghost planes on each side:

    component layout  (halo + n0_local + halo, n1, n2)      <- pointers handed to kernels address owned plane 0

Single-kernel form (default; 'constant' boundaries): one pxb_pds_iter launch per iteration and slab, ping-pong
buffers.  The kernel recomputes w on the ghost plane above the slab, so what crosses each interface per iteration is
the NEW iterate's boundary planes: u, z_0, z_1, z_2 of the first owned plane go down, z_0 of the last owned plane
goes up (5 planes, 20 MiB at 1024^2 fp32).  The two boundary chunks of a slab are launched first; their planes travel
on a side stream (NCCL send/recv) while the interior chunks are computed.

Two-sweep form (other boundary modes): per PD3O-TV iteration two planes cross each slab interface:
    * before the primal half-step: the dual component along z, z_0, last owned plane  -> upper neighbour
      (K^T z at a slab's first plane needs z_0 of the plane below),
    * before the dual half-step:   w, first owned plane                              -> lower neighbour
      (forward difference of w at a slab's last plane needs w of the plane above),
through NCCL send/recv (`torch.distributed.batch_isend_irecv`).  The stopping-criterion norms of all
ranks are combined by ONE all-reduce of a 4-double vector per evaluation.  Independent images of a
batch need none of this: they are simply dealt out to the ranks (see `split_batch`).

With `overlap=True` the exchange runs on a side stream while the interior planes are being computed: each
half-step is issued as [interior planes] on the main stream and [boundary plane] after the halo arrived.
"""
import ctypes as C
import math
import os

import torch
import torch.distributed as dist

from . import _array as A
from . import _cabi as K


_HP_GROUP = None


def _high_priority_group():
    """A process group over all ranks whose NCCL kernels run on a high-priority stream (created once)."""
    global _HP_GROUP
    if _HP_GROUP is None and dist.is_initialized() and dist.get_backend() == "nccl" and dist.get_world_size() > 1:
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            _HP_GROUP = dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)
        except Exception:
            _HP_GROUP = False
    return _HP_GROUP or None


def partition(n0, world):
    """Balanced contiguous split of n0 planes into `world` slabs: list of (start, stop)."""
    base, rem = divmod(int(n0), int(world))
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def split_batch(n_items, world):
    """Independent images of a batch: contiguous shares, no communication."""
    return partition(n_items, world)


class HaloExchanger:
    """Moves boundary planes between neighbouring slabs.  Works on any torch.distributed backend
    (NCCL on GPUs; gloo on CPU tensors in the unit tests of the plumbing)."""

    def __init__(self, rank=None, world=None, group=None, periodic=False):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.periodic = bool(periodic) and self.world > 1
        self.lo = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.periodic else None)
        self.hi = self.rank + 1 if self.rank < self.world - 1 else (0 if self.periodic else None)

    def exchange(self, buf, halo, n_owned, up=True, down=True):
        """buf: (halo + n_owned + halo, n1, n2) tensor.
        up:   my last `halo` owned planes  -> upper neighbour's lower ghost planes;
        down: my first `halo` owned planes -> lower neighbour's upper ghost planes.
        Returns the list of outstanding requests (call .wait() on each)."""
        ops = []
        h = halo
        if up:
            if self.hi is not None:
                ops.append(dist.P2POp(dist.isend, buf[n_owned : n_owned + h], self.hi, self.group))
            if self.lo is not None:
                ops.append(dist.P2POp(dist.irecv, buf[0:h], self.lo, self.group))
        if down:
            if self.lo is not None:
                ops.append(dist.P2POp(dist.isend, buf[h : 2 * h], self.lo, self.group))
            if self.hi is not None:
                ops.append(dist.P2POp(dist.irecv, buf[h + n_owned : 2 * h + n_owned], self.hi, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def exchange_many(self, items, halo, n_owned):
        """One batched exchange of several buffers: items = [(buf, up, down), ...] (same order on every rank)."""
        ops = []
        h = halo
        for buf, up, down in items:
            if up:
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.isend, buf[n_owned : n_owned + h], self.hi, self.group))
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[0:h], self.lo, self.group))
            if down:
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.isend, buf[h : 2 * h], self.lo, self.group))
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[h + n_owned : 2 * h + n_owned], self.hi, self.group))
        return dist.batch_isend_irecv(ops) if ops else []


class SlabPD3OTV:
    """PD3O on  min 1/2||x - y||^2 + i_+(x) + lam*||grad x||_{2,1}  for a volume decomposed in z-slabs.

    Same iteration as pyxu_b200.opt.solver.PD3O on the fused path (pxb_pds_primal + pxb_pds_dual), issued
    per rank on its slab with `open_lo/open_hi` set where a neighbour exists.  Step sizes follow
    PD3O._set_step_sizes for beta = 1 and the Gradient's Lipschitz bound.
    """

    HALO = 1

    def __init__(self, shape, y_full=None, y_local=None, lam=0.08, positivity=True, dtype=torch.float32, mode="constant",
                 rho=1.0, tau=None, sigma=None, group=None, overlap=True, fused=True, edge=8):
        from .operator.linop.diff import Gradient

        A.require_cuda()
        assert len(shape) == 3
        self.shape = tuple(int(s) for s in shape)
        self.group = group if group is not None else _high_priority_group()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = A.current_device()
        self.dtype = dtype
        modes = (mode,) * 3 if isinstance(mode, str) else tuple(mode)
        self.hx = HaloExchanger(self.rank, self.world, self.group, periodic=(modes[0] == "wrap"))
        self.start, self.stop = partition(self.shape[0], self.world)[self.rank]
        self.n0 = self.stop - self.start
        n1, n2 = self.shape[1:]
        self.plane = n1 * n2
        self.local_voxels = self.n0 * self.plane
        h = self.HALO
        assert self.n0 >= 2 * h + 1, "slabs thinner than 3 planes are not supported"
        self.K = Gradient(arg_shape=self.shape, mode=modes, dtype=A.np_dtype(dtype))
        alloc = self.n0 + 2 * h

        def field(ncomp=1):
            return torch.zeros((ncomp, alloc, n1, n2), dtype=dtype, device=self.dev)

        # single-kernel iteration; ping-pong (u, z) pairs, no w array.  'constant' boundaries by default; with folding modes
        # (in-plane folds on every rank, folds along z on the closed sides of the end ranks, 'wrap' along z as a ring of open
        # sides) the kernels are checked rank by rank on the CPU emulation (tests/test_slab_cpu.py) but have not run on
        # several GPUs yet: PYXU_B200_SLAB_FUSED_MODES=1 switches them on, otherwise those problems take the two-sweep form.
        self.fused = bool(fused) and (all(m == "constant" for m in modes) or os.environ.get("PYXU_B200_SLAB_FUSED_MODES", "0") == "1")
        self.edge = max(1, min(int(edge), self.n0 // 2))  # planes of the boundary launches that precede the exchange
        if y_local is None:
            y_local = y_full.reshape(self.shape)[self.start : self.stop]
        own = slice(h, h + self.n0)
        self.x = field()
        self._x_stale = False
        self._iter_cache, self._fn_iter = {}, K.lib().pxb_pds_iter
        if self.fused:
            self._ub, self._zb, self.cur = [field(), field()], [field(3), field(3)], 0
            self.w = None
            self.shift_h = field()  # shift (= -y) WITH ghost planes: the kernel evaluates grad f on the ghost plane too
            self.shift_h[0, own].copy_(y_local)
            self.shift_h.neg_()
            self.shift = self.shift_h[0, own]
        else:
            self._ub, self._zb, self.cur = [field()], [field(3)], 0
            self.w = field()
            self.shift = (-y_local).to(dtype).contiguous()  # f = 1/2 ||x + shift||^2
        self.u[0, own].copy_(y_local)
        self.x[0, own].copy_(y_local)
        # step sizes: PD3O defaults (reference: pds.py:807-829, 849-864) for beta = 1
        L = self.K.lipschitz
        t = math.exp(min(0.5 * (math.log(0.99) - 2 * math.log(L)), 0.0))
        self.tau = t if tau is None else tau
        self.sigma = t if sigma is None else sigma
        self.rho = rho
        self.lam, self.positivity = lam, positivity
        # Sub-range launches treat the cut as an open side, which is only exact when no boundary fold reaches across
        # it: 'constant' (nothing folds) or 'wrap' (ring exchange: every side is open).  Other modes along z run
        # whole-slab launches after the exchange.
        self.overlap = bool(overlap) and self.world > 1 and modes[0] in ("constant", "wrap")
        # The exchange must not queue behind the interior kernel's ~10^5 pending CTAs: measured on 8 GPUs, with default
        # priorities the NCCL send/recv kernel only ran once the interior grid had drained (0.25 ms of an 1.04 ms iteration
        # exposed).  Side stream and NCCL stream are therefore high-priority: their CTAs are placed as soon as a CTA
        # of the interior kernel retires (every ~16 us).
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1) if (self.overlap or self.fused) else None
        self.event_log = []   # (tag, cuda event) pairs, filled when `record_events`
        self.record_events = True
        # z0 = K x0 needs x0's upper ghost plane
        self._wait(self.hx.exchange(self.x[0], h, self.n0, up=False, down=True))
        d = self._desc(0, self.n0)
        K.check(K.lib().pxb_gradient_apply(C.byref(d), self._p(self.x, 0, 0), self._p(self.z, 0, 0), A.stream()), "gradient_apply")
        if self.fused and self.world > 1:  # ghost planes of the data term (once) and of the initial iterate
            self._wait(self.hx.exchange(self.shift_h[0], h, self.n0, up=True, down=True))
            self._wait(self._exchange_state(self.cur))
        torch.cuda.synchronize()

    # current iterate (the ping-pong index flips every fused iteration)
    @property
    def u(self):
        return self._ub[self.cur]

    @property
    def z(self):
        return self._zb[self.cur]

    # -- helpers ----------------------------------------------------------------------------------
    @staticmethod
    def _wait(reqs):
        for r in reqs:
            r.wait()

    def _p(self, t, comp, plane):
        """Device pointer to owned plane `plane` of component `comp`."""
        h = self.HALO
        return C.c_void_p(t.data_ptr() + t.element_size() * ((comp * t.shape[1] + h + plane) * self.plane))

    def _desc(self, p0, p1):
        """Gradient descriptor for owned planes [p0, p1) of this slab."""
        h = self.HALO
        open_lo = 1 if (p0 > 0 or self.hx.lo is not None) else 0
        open_hi = 1 if (p1 < self.n0 or self.hx.hi is not None) else 0
        slab = K.Slab(open_lo, open_hi, h, self.n0 + 2 * h)
        return self.K._desc(1, A.dcode(self.u), slab=slab, shape0=p1 - p0)

    def _params(self):
        p = K.PdsParams()
        p.tau, p.sigma, p.rho = self.tau, self.sigma, self.rho
        p.g = K.ProxSpec(K.PROX_POS if self.positivity else K.PROX_NONE, 0, 0.0, 0.0)
        f = K.FTerm()
        f.kind, f.alpha = K.F_SQL2, 0.5
        p.f = f
        p.hkind, p.lam = K.DUAL_L21, self.lam
        return p

    def _primal(self, p0, p1, norms=None):
        d, p = self._desc(p0, p1), self._params()
        # the shift (= -y) is stored without ghost planes: address its plane p0 directly
        p.f.shift = self.shift.data_ptr() + self.shift.element_size() * p0 * self.plane
        p.f.shift_period = (self.n0 - p0) * self.plane
        rc = K.lib().pxb_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(p), self._p(self.u, 0, p0), self._p(self.z, 0, p0), None,
                                    self._p(self.x, 0, p0), self._p(self.w, 0, p0), A.ptr(norms), A.stream())
        K.check(rc, "pxb_pds_primal")

    def _dual(self, p0, p1, norms=None):
        d, p = self._desc(p0, p1), self._params()
        rc = K.lib().pxb_pds_dual(C.byref(d), C.byref(p), self._p(self.w, 0, p0), self._p(self.z, 0, p0), A.ptr(norms), A.stream())
        K.check(rc, "pxb_pds_dual")

    def _tick(self, tag):
        if self.record_events:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.event_log.append((tag, ev))

    # -- single-kernel form ----------------------------------------------------------------------------
    def _exchange_state(self, idx):
        """Boundary planes of iterate `idx` -> the neighbours' ghost planes: u, z_0..2 first plane down, z_0 last plane up."""
        u, z = self._ub[idx], self._zb[idx]
        items = [(u[0], False, True), (z[0], True, True), (z[1], False, True), (z[2], False, True)]
        return self.hx.exchange_many(items, self.HALO, self.n0)

    def run(self, n_steps):
        """n_steps iterations without stopping-criterion norms."""
        for _ in range(n_steps):
            self.step(False)

    def _iter(self, p0, p1, src, dst, x_out, nx, nz):
        """pxb_pds_iter on owned planes [p0, p1): reads iterate `src` (ghost planes valid), writes iterate `dst`."""
        key = (p0, p1, src)
        c = self._iter_cache.get(key)
        if c is None:  # descriptors and pointers never change: build the ctypes arguments once (host time matters at 1 ms/iteration)
            d, p = self._desc(p0, p1), self._params()
            sh = self.shift_h
            p.f.shift = sh.data_ptr() + sh.element_size() * (self.HALO + p0) * self.plane
            p.f.shift_period = sh.shape[1] * self.plane  # >= the span of the launch: addressed like u (ghost planes included)
            ptr = lambda t, comp: self._p(t, comp, p0)
            c = (d, p, C.byref(d), C.byref(p), ptr(self._ub[src], 0), ptr(self._zb[src], 0), ptr(self._ub[dst], 0), ptr(self._zb[dst], 0), ptr(self.x, 0))
            self._iter_cache[key] = c
        rc = self._fn_iter(K.ALGO_PD3O, c[2], c[3], c[4], c[5], c[6], c[7], c[8] if x_out else None, A.ptr(nx), A.ptr(nz), A.stream())
        if rc:
            K.check(rc, "pxb_pds_iter")

    def _step_fused(self, want_norms):
        n0, e = self.n0, self.edge
        src, dst = self.cur, 1 - self.cur
        nx = nz = None
        if want_norms:
            if self._x_stale:
                self.materialize_x()
            nrm = torch.zeros((2, 1, 2), dtype=torch.float64, device=self.dev)
            nx, nz = nrm[0], nrm[1]
        main = torch.cuda.current_stream()
        self._tick("iter_begin")
        if self.world > 1 and self.overlap and n0 >= 4 * e:
            lo, hi = (e if self.hx.lo is not None else 0), (n0 - e if self.hx.hi is not None else n0)
            if lo:
                self._iter(0, lo, src, dst, want_norms, nx, nz)
            if hi < n0:
                self._iter(hi, n0, src, dst, want_norms, nx, nz)
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                self._wait(self._exchange_state(dst))  # the new boundary planes travel while the interior is computed
            self._iter(lo, hi, src, dst, want_norms, nx, nz)
        else:
            self._iter(0, n0, src, dst, want_norms, nx, nz)
            if self.world > 1:
                self.comm.wait_stream(main)
                with torch.cuda.stream(self.comm):
                    self._wait(self._exchange_state(dst))
        if self.world > 1:
            main.wait_stream(self.comm)  # the next iteration reads the ghost planes this exchange fills
        self._tick("iter_end")
        self.cur = dst
        self._x_stale = not want_norms
        if want_norms:
            v = nrm.reshape(-1)
            dist.all_reduce(v, group=self.group)
            return v.cpu().numpy()
        return None

    def materialize_x(self):
        """x_k = prox_g(u_{k-1} - tau K^T z_{k-1}) from the previous iterate (kept in the other ping-pong pair)."""
        if not (self.fused and self._x_stale):
            return
        prev = 1 - self.cur
        tmp_u, tmp_w = self._ub[prev].clone(), torch.empty_like(self._ub[prev])
        d, p = self._desc(0, self.n0), self._params()
        sh = self.shift_h
        p.f.shift = sh.data_ptr() + sh.element_size() * self.HALO * self.plane
        p.f.shift_period = sh.shape[1] * self.plane
        rc = K.lib().pxb_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(p), self._p(tmp_u, 0, 0), self._p(self._zb[prev], 0, 0), None,
                                    self._p(self.x, 0, 0), self._p(tmp_w, 0, 0), None, A.stream())
        K.check(rc, "pxb_pds_primal")
        self._x_stale = False

    # -- one PD3O iteration ------------------------------------------------------------------------
    def step(self, want_norms=False):
        if self.fused:
            return self._step_fused(want_norms)
        h, n0 = self.HALO, self.n0
        nx = nz = None
        if want_norms:  # kernels accumulate into (rows, 2) buffers
            nx = torch.zeros((1, 2), dtype=torch.float64, device=self.dev)
            nz = torch.zeros((1, 2), dtype=torch.float64, device=self.dev)
        main = torch.cuda.current_stream()
        if not self.overlap:
            if self.world > 1:
                self._wait(self.hx.exchange(self.z[0], h, n0, up=True, down=False))
            self._tick("primal_begin")
            self._primal(0, n0, nx)
            self._tick("primal_end")
            if self.world > 1:
                self._wait(self.hx.exchange(self.w[0], h, n0, up=False, down=True))
            self._dual(0, n0, nz)
            self._tick("dual_end")
        else:
            comm = self.comm
            # (1) z_0 ghost plane travels while planes [1, n0) take their primal half-step
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                self._wait(self.hx.exchange(self.z[0], h, n0, up=True, down=False))
            self._tick("primal_begin")
            self._primal(1, n0, nx)
            main.wait_stream(comm)
            self._primal(0, 1, nx)
            self._tick("primal_end")
            # (2) w's first plane travels while planes [0, n0-1) take their dual half-step
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                self._wait(self.hx.exchange(self.w[0], h, n0, up=False, down=True))
            self._dual(0, n0 - 1, nz)
            main.wait_stream(comm)
            self._dual(n0 - 1, n0, nz)
            self._tick("dual_end")
        if want_norms:
            v = torch.cat([nx.reshape(-1), nz.reshape(-1)])
            dist.all_reduce(v, group=self.group)  # the single fused scalar all-reduce of the stopping criterion
            return v.cpu().numpy()
        return None

    def rel_errors(self, v):
        """(RelError[x], RelError[z]) from the all-reduced norm vector returned by step(want_norms=True)."""
        import numpy as np

        with np.errstate(all="ignore"):
            rx, rz = np.sqrt(v[0]) / np.sqrt(v[1]), np.sqrt(v[2]) / np.sqrt(v[3])
        return float(np.nan_to_num(rx)), float(np.nan_to_num(rz))

    def gather_x(self):
        """Full primal iterate on every rank (tests / small volumes only)."""
        self.materialize_x()
        h = self.HALO
        parts = partition(self.shape[0], self.world)
        nmax = max(b - a for a, b in parts)
        mine = torch.zeros((nmax, *self.shape[1:]), dtype=self.dtype, device=self.dev)
        mine[: self.n0].copy_(self.x[0, h : h + self.n0])
        bufs = [torch.empty_like(mine) for _ in parts]
        dist.all_gather(bufs, mine, group=self.group)
        return torch.cat([b_[: e - a] for b_, (a, e) in zip(bufs, parts)], dim=0)


class SlabCondatVuDeblur:
    """CondatVu on  min 1/2||A x - y||^2 [+ i_+(x)] + lam*||grad x||_{2,1}  for a volume decomposed in z-slabs
    (configs[4] of BASELINE.json: 3-D TV deblurring with a separable Stencil PSF).

    A is a `constant`-mode separable Stencil (its factor along axis 0 reaches H planes across a cut), the Gradient is
    the forward-difference stack.  One iteration on every rank (reference iteration: pds.py:429-442):
        tmp = A_0 x          (axis-0 streaming pass; reads H ghost planes of x)
        r   = A_12 tmp - y   (tiled in-plane pass, epilogue carries -y)                  -> r ghosts exchanged (H planes)
        tmp = A_0^T r ;  grad f = A_12^T tmp                                              -> first plane of grad f goes down
        (x, z) <- pxb_pds_iter(CV, grad f array)   (single kernel, ping-pong)             -> x ghosts (H planes), z planes
    Exchanges are NCCL send/recv batches on a high-priority stream.
    """

    def __init__(self, shape, psf, center, y_full=None, y_local=None, lam=0.05, positivity=True, dtype=torch.float32, rho=1.0,
                 tau=None, sigma=None, group=None, overlap=True):
        import numpy as np

        from .operator.linop.diff import Gradient
        from .operator.linop.stencil import Stencil

        A.require_cuda()
        assert len(shape) == 3 and len(psf) == 3, "3-D volume and a separable PSF (one 1-D factor per axis)"
        self.shape = tuple(int(s) for s in shape)
        self.group = group if group is not None else _high_priority_group()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev, self.dtype = A.current_device(), dtype
        npdt = A.np_dtype(dtype)
        self.Aop = Stencil(arg_shape=self.shape, kernel=[np.asarray(k, dtype=npdt) for k in psf], center=tuple(center), mode="constant")
        self.K = Gradient(arg_shape=self.shape, dtype=npdt)
        self.hx = HaloExchanger(self.rank, self.world, self.group)
        self.start, self.stop = partition(self.shape[0], self.world)[self.rank]
        self.n0 = n0 = self.stop - self.start
        n1, n2 = self.shape[1:]
        self.plane = n1 * n2
        self.local_voxels = n0 * self.plane
        k0, c0 = len(psf[0]), int(center[0])
        self.H = H = max(1, c0, k0 - 1 - c0)
        assert n0 >= H, "slabs thinner than the PSF's reach along z are not supported"
        alloc = n0 + 2 * H

        def field(ncomp=1):
            return torch.zeros((ncomp, alloc, n1, n2), dtype=dtype, device=self.dev)

        self._xb, self._zb, self.cur = [field(), field()], [field(3), field(3)], 0
        self.r, self.garr, self.tmp = field(), field(), field()
        if y_local is None:
            y_local = y_full.reshape(self.shape)[self.start : self.stop]
        self.neg_y = (-y_local).to(dtype).contiguous()
        own = slice(H, H + n0)
        self._xb[0][0, own].copy_(y_local)
        # step sizes: CondatVu defaults (reference: pds.py:444-517) with beta = ||A||^2, gamma = beta
        beta = float(self.Aop.lipschitz) ** 2
        L = float(self.K.lipschitz)
        t = (1.0 / L**2) * ((-beta / 2) + math.sqrt((beta**2 / 4) + L**2))
        self.tau = t if tau is None else tau
        self.sigma = t if sigma is None else sigma
        self.rho, self.lam, self.positivity = rho, lam, positivity
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)
        self.Aop._slab_dcode = A.dcode(self.r)
        self._plans = {adj: self.Aop._tiled_plan(adj) for adj in (False, True)}
        assert all(p is not None and p[0] is not None for p in self._plans.values()), "expected a separable PSF with a factor along axis 0"
        self._alloc = alloc
        self._slab = K.Slab(1 if self.hx.lo is not None else 0, 1 if self.hx.hi is not None else 0, H, alloc)
        self._cache = {}   # descriptors per (kind, adjoint / parity, p0, p1): built once
        self.single_pass = self.Aop._desc3d(A.dcode(self.r), False, 1, slab=self._slab, shape0=n0) is not None
        self.edge = max(H, min(8, n0 // 4))  # planes of the boundary launches that precede each exchange
        self.overlap = bool(overlap) and self.world > 1 and n0 >= 4 * self.edge
        self._gdesc = self.K._desc(1, A.dcode(self.r), slab=self._slab, shape0=n0)
        # z0 = K x0 needs x0's ghost planes
        self._wait(self._exchange([(self._xb[0][0], H, H)]))
        K.check(K.lib().pxb_gradient_apply(C.byref(self._gdesc), self._p(self._xb[0], 0), self._p(self._zb[0], 0), A.stream()), "gradient_apply")
        self._wait(self._exchange([(self._zb[0][0], 1, 1), (self._zb[0][1], 0, 1), (self._zb[0][2], 0, 1)]))
        torch.cuda.synchronize()

    @property
    def x(self):
        return self._xb[self.cur]

    @property
    def z(self):
        return self._zb[self.cur]

    @staticmethod
    def _wait(reqs):
        for r in reqs:
            r.wait()

    def _p(self, t, comp, plane=0):
        return C.c_void_p(t.data_ptr() + t.element_size() * ((comp * t.shape[1] + self.H + plane) * self.plane))

    def _exchange(self, items):
        """items: [(buf of shape (alloc, n1, n2), planes going up, planes going down)] -- my last `up` owned planes fill
        the upper neighbour's ghost planes next to its first owned plane, my first `down` owned planes the lower
        neighbour's ghost planes next to its last owned plane."""
        H, n0, ops = self.H, self.n0, []
        for buf, up, down in items:
            if up:
                if self.hx.hi is not None:
                    ops.append(dist.P2POp(dist.isend, buf[H + n0 - up : H + n0], self.hx.hi, self.group))
                if self.hx.lo is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[H - up : H], self.hx.lo, self.group))
            if down:
                if self.hx.lo is not None:
                    ops.append(dist.P2POp(dist.isend, buf[H : H + down], self.hx.lo, self.group))
                if self.hx.hi is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[H + n0 : H + n0 + down], self.hx.hi, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def _sub_slab(self, p0, p1):
        """Slab flags of a launch restricted to owned planes [p0, p1): an interior cut is an open side."""
        lo = 1 if (p0 > 0 or self.hx.lo is not None) else 0
        hi = 1 if (p1 < self.n0 or self.hx.hi is not None) else 0
        return K.Slab(lo, hi, self.H, self._alloc)

    def _stencil(self, adjoint, src, dst, p0=0, p1=None):
        """dst[p0:p1] = (A src - y | A^T src)[p0:p1], reading src's neighbouring / ghost planes: one marching pass
        (pxb_stencil3d_apply), or axis-0 streaming pass + tiled in-plane pass outside its envelope."""
        p1 = self.n0 if p1 is None else p1
        key = ("st", adjoint, p0, p1)
        c = self._cache.get(key)
        if c is None:
            slab = self._sub_slab(p0, p1)
            add = None if adjoint else self.neg_y[p0:p1]
            d3 = self.Aop._desc3d(A.dcode(self.r), adjoint, 1, 1.0, 0.0 if adjoint else 1.0, add, slab=slab, shape0=p1 - p0) if self.single_pass else None
            d2 = None
            if d3 is None:
                own = slice(self.H + p0, self.H + p1)
                d2, _ = self.Aop._tiled_desc(self.r[0, own], adjoint, alpha=1.0, beta=0.0 if adjoint else 1.0, add=add)
                d2.nimg = p1 - p0
            c = self._cache[key] = (d3, d2, slab, self._plans[adjoint][0])
        d3, d2, slab, axis0 = c
        if d3 is not None:
            rc = K.lib().pxb_stencil3d_apply(C.byref(d3), self._p(src, 0, p0), self._p(dst, 0, p0), A.stream())
            if rc == -3:  # outside the marching kernel's envelope (e.g. an even number of taps along z): two passes from now on
                self.single_pass = False
                self._cache = {k: v for k, v in self._cache.items() if k[0] != "st"}
                return self._stencil(adjoint, src, dst, p0, p1)
            K.check(rc, "pxb_stencil3d_apply")
            return
        self.Aop._axis0_pass(axis0, self._p(src, 0, p0), self._p(self.tmp, 0, p0), 1, slab=slab, shape0=p1 - p0)
        K.check(K.lib().pxb_stencil2d_apply(C.byref(d2), self._p(self.tmp, 0, p0), self._p(dst, 0, p0), A.stream()), "pxb_stencil2d_apply")

    def _iter(self, src, dst, nx, nz, p0=0, p1=None):
        p1 = self.n0 if p1 is None else p1
        key = ("it", src, p0, p1)
        c = self._cache.get(key)
        if c is None:
            gd = self.K._desc(1, A.dcode(self.r), slab=self._sub_slab(p0, p1), shape0=p1 - p0)
            p = K.PdsParams()
            p.tau, p.sigma, p.rho = self.tau, self.sigma, self.rho
            p.g = K.ProxSpec(K.PROX_POS if self.positivity else K.PROX_NONE, 0, 0.0, 0.0)
            f = K.FTerm()
            f.kind, f.garr = K.F_GRADARR, self._p(self.garr, 0, p0).value
            p.f = f
            p.hkind, p.lam = K.DUAL_L21, self.lam
            xs, zs, xd, zd = self._xb[src], self._zb[src], self._xb[dst], self._zb[dst]
            c = self._cache[key] = (gd, p, self._p(xs, 0, p0), self._p(zs, 0, p0), self._p(xd, 0, p0), self._p(zd, 0, p0))
        gd, p, a0, a1, a2, a3 = c
        rc = K.lib().pxb_pds_iter(K.ALGO_CV, C.byref(gd), C.byref(p), a0, a1, a2, a3, None, A.ptr(nx), A.ptr(nz), A.stream())
        K.check(rc, "pxb_pds_iter")

    def _staged(self, fn, exchange_items):
        """One stage of the iteration: `fn(p0, p1)` on the boundary chunks, their planes sent on the side stream while
        `fn` runs on the interior; the next stage starts once the ghost planes have arrived."""
        n0, e = self.n0, self.edge
        main = torch.cuda.current_stream()
        if self.overlap:
            lo, hi = (e if self.hx.lo is not None else 0), (n0 - e if self.hx.hi is not None else n0)
            if lo:
                fn(0, lo)
            if hi < n0:
                fn(hi, n0)
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                self._wait(self._exchange(exchange_items))
            fn(lo, hi)
            main.wait_stream(self.comm)
        else:
            fn(0, n0)
            self._wait(self._exchange(exchange_items))

    def step(self, want_norms=False):
        H = self.H
        src, dst = self.cur, 1 - self.cur
        xs, xd, zd = self._xb[src], self._xb[dst], self._zb[dst]
        nrm = nx = nz = None
        if want_norms:
            nrm = torch.zeros((2, 1, 2), dtype=torch.float64, device=self.dev)
            nx, nz = nrm[0], nrm[1]
        self._staged(lambda a, b: self._stencil(False, xs, self.r, a, b), [(self.r[0], H, H)])          # r = A x - y
        self._staged(lambda a, b: self._stencil(True, self.r, self.garr, a, b), [(self.garr[0], 0, 1)])  # grad f = A^T r
        self._staged(lambda a, b: self._iter(src, dst, nx, nz, a, b),
                     [(xd[0], H, H), (zd[0], 1, 1), (zd[1], 0, 1), (zd[2], 0, 1)])                        # (x, z) <- CV iteration
        self.cur = dst
        if want_norms:
            v = nrm.reshape(-1)
            dist.all_reduce(v, group=self.group)
            return v.cpu().numpy()
        return None

    def gather_x(self):
        H = self.H
        parts = partition(self.shape[0], self.world)
        nmax = max(b - a for a, b in parts)
        mine = torch.zeros((nmax, *self.shape[1:]), dtype=self.dtype, device=self.dev)
        mine[: self.n0].copy_(self.x[0, H : H + self.n0])
        bufs = [torch.empty_like(mine) for _ in parts]
        dist.all_gather(bufs, mine, group=self.group)
        return torch.cat([b_[: e - a] for b_, (a, e) in zip(bufs, parts)], dim=0)
