"""
z-slab domain decomposition of large 3-D volumes across the GPUs of one node, behind `Solver.fit()`.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The volume is cut along axis 0 into `world`
contiguous slabs; every rank stores its slab with `halo` ghost planes on each side:

    component layout  (halo + n0_local + halo, n1, n2)      <- pointers handed to kernels address owned plane 0

The user-facing side is the reference's own (pds.py:523, 723, 747): the SAME `PD3O(f, g, h, K)` / `CondatVu(f, g, h, K)`
object is built on every rank over the GLOBAL volume shape, and `fit(x0=...)` is called on every rank.  When
torch.distributed is initialised with more than one rank and the problem has the fused TV structure (`_Plan.kind ==
"fused"` on a 3-D volume), `m_init` hands the iteration to one of the engines below; step sizes, momentum, stopping
criteria, history and `stats()` / `solution()` stay the solver's (nothing is re-derived here).  Arrays go in either as
full-volume arrays (every rank passes the same array and keeps its own planes; `solution()` returns the gathered
volume on every rank) or as `ShardedArray`s (each rank passes / receives only its planes).

Engines:

* `SlabTV` -- PD3O / CondatVu with a pointwise data term.  Single-kernel form: one pxb_pds_iter launch per iteration
  and slab, ping-pong buffers.  The kernel recomputes w on the ghost plane above the slab, so what crosses each
  interface per iteration is the NEW iterate's boundary planes: u, z_0, z_1, z_2 of the first owned plane go down, z_0
  of the last owned plane goes up (5 planes, 20 MiB at 1024^2 fp32).  The two boundary chunks of a slab are launched
  first; their planes travel on a side stream (NCCL send/recv) while the interior chunks are computed.
  Two-sweep form (outside the single-kernel envelope): per iteration two planes cross each slab interface:
    * before the primal half-step: the dual component along z, z_0, last owned plane  -> upper neighbour,
    * before the dual half-step:   w, first owned plane                              -> lower neighbour.
* `SlabDeblurCV` -- CondatVu with f = alpha*||A x + c||^2, A a separable 'constant'-mode 3-D Stencil (configs[4]).

The stopping-criterion sums of all ranks are combined by ONE all-reduce of a 4-double vector per evaluation (issued by
the solver on the buffer the kernels accumulate into).  Independent images of a batch need none of this: they are
simply dealt out to the ranks (see `split_batch`).
"""
import ctypes as C

import numpy as np
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


def context(distributed=None):
    """(rank, world) of the slab decomposition, or None for a single-domain solve.
    distributed: None = decompose when torch.distributed runs more than one rank; True = always (a single rank, with or
    without torch.distributed, included: the slab engines then run without neighbours); False = never."""
    if distributed is False:
        return None
    on = dist.is_available() and dist.is_initialized()
    if distributed is True:
        return (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    if on and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return None


class ShardedArray:
    """This rank's planes of a global array laid out (comps, n0, n1, n2) -> flattened (comps * N,).

    `local` holds planes [start, stop) (see `partition`) of every component, C-ordered (comps, n0_local, n1, n2), as a
    NumPy array (host) or a device tensor.  Operators and solvers see the GLOBAL flattened shape, so the object can stand
    where the reference takes the full array: `argshift(ShardedArray)`, `fit(x0=ShardedArray)`; `solution()` then
    returns a ShardedArray too (no gather)."""

    def __init__(self, local, vol_shape, comps=1, rank=None, world=None):
        self.vol_shape = tuple(int(s) for s in vol_shape)
        assert len(self.vol_shape) == 3, "ShardedArray: 3-D volumes cut along axis 0"
        if rank is None:
            rank, world = context(True)
        self.rank, self.world, self.comps = int(rank), int(world), int(comps)
        self.start, self.stop = partition(self.vol_shape[0], self.world)[self.rank]
        self.plane = self.vol_shape[1] * self.vol_shape[2]
        n_local = self.comps * (self.stop - self.start) * self.plane
        got = int(local.numel()) if hasattr(local, "numel") else int(np.size(local))
        if got != n_local:
            raise ValueError(f"ShardedArray: rank {self.rank} of {self.world} holds planes [{self.start}, {self.stop}) = {n_local} samples, got {got}")
        self.local = local
        self.size = self.comps * int(np.prod(self.vol_shape))
        self.shape = (self.size,)
        self.ndim = 1

    @property
    def dtype(self):
        return self.local.dtype

    def numel(self):
        return self.size

    def __repr__(self):
        return f"ShardedArray(planes [{self.start}, {self.stop}) of {self.vol_shape}, comps={self.comps}, rank {self.rank}/{self.world})"


def local_part(arr, vol_shape, rank, world, comps=1):
    """This rank's planes of `arr` (full flattened array of comps * N samples, or a ShardedArray), still in the
    caller's memory space: a (comps, n0_local, n1, n2) view / copy.  Returns (local, was_sharded)."""
    if isinstance(arr, ShardedArray):
        if arr.vol_shape != tuple(vol_shape) or arr.comps != comps or (arr.rank, arr.world) != (rank, world):
            raise ValueError(f"{arr!r} does not match the decomposition of {tuple(vol_shape)} (comps={comps}) over rank {rank}/{world}")
        loc = arr.local
        a, b = arr.start, arr.stop
        return loc.reshape(comps, b - a, *vol_shape[1:]), True
    a, b = partition(vol_shape[0], world)[rank]
    if not (isinstance(arr, (np.ndarray, torch.Tensor))):
        arr = torch.from_dlpack(arr) if hasattr(arr, "__dlpack__") else np.asarray(arr)
    full = arr.reshape(comps, *vol_shape)
    return full[:, a:b], False


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
        return self.exchange_many([(buf, up, down)], halo, n_owned)

    def exchange_many(self, items, halo, n_owned):
        """One batched exchange of several buffers: items = [(buf, up, down), ...] (same order on every rank)."""
        return self.exchange_planes([(buf, halo if up else 0, halo if down else 0) for buf, up, down in items], halo, n_owned)

    def exchange_planes(self, items, halo, n_owned):
        """items: [(buf of shape (alloc, n1, n2), planes going up, planes going down)] -- my last `up` owned planes fill the
        upper neighbour's ghost planes next to its first owned plane, my first `down` owned planes the lower neighbour's
        ghost planes next to its last owned plane."""
        H, n0, ops = halo, n_owned, []
        for buf, up, down in items:
            if up:
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.isend, buf[H + n0 - up : H + n0], self.hi, self.group))
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[H - up : H], self.lo, self.group))
            if down:
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.isend, buf[H : H + down], self.lo, self.group))
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.irecv, buf[H + n0 : H + n0 + down], self.hi, self.group))
        return dist.batch_isend_irecv(ops) if ops else []


# -- peer memory: the neighbours' slab buffers mapped into this process ---------------------------------------------------
_PEER_POOL = {}  # (kind, shape, dtype, rank, world) -> dict(bufs=..., lo=..., hi=...): kept for the life of the process (see PeerBuffers)


def release_pool():
    """Drop the pooled slab buffers (collective in spirit: call it on every rank, with no solve in flight)."""
    _PEER_POOL.clear()


def _p2p_enabled():
    import os

    return (os.environ.get("PYXU_B200_SLAB_P2P", "1") != "0" and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
            and dist.get_world_size() > 1 and torch.cuda.is_available())


class PeerBuffers:
    """The state buffers of a SlabTV engine together with the lower / upper neighbour's, the latter mapped into this process
    through CUDA IPC (torch's own tensor sharing: cudaIpcGetMemHandle on the exporting rank, cudaIpcOpenMemHandle with lazy
    peer access here), so that the iteration kernel can store boundary planes straight into the neighbours' ghost planes over
    NVLink (pxb_pds_iter_p2p).  Buffers and mappings are POOLED per (shape, dtype, rank, world): mapping costs milliseconds, and a
    neighbour may still be writing its last boundary planes into this rank's ghost planes when this rank's solve has already
    returned -- memory that is never handed back to the allocator cannot be corrupted by that."""

    def __init__(self, make, rank, world, lo, hi):
        from torch.multiprocessing.reductions import reduce_tensor

        self.bufs = make()  # name -> tensor
        mine = {k: reduce_tensor(t) for k, t in self.bufs.items()}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        opened = {}

        me = torch.cuda.current_device()

        def open_rank(r):
            """Neighbour r's buffers as tensors whose data_ptr() is valid in kernels launched on THIS device: the IPC handle is
            opened with this device current (argument 6 of torch's rebuild function is the device the mapping is made for;
            cudaIpcOpenMemHandle then sets up peer access to the exporting GPU by itself), after peer access has been switched
            on explicitly as well."""
            if r is None:
                return None
            if r not in opened:
                out = {}
                for k, (fn, args) in everyone[r].items():
                    args = list(args)
                    if int(args[6]) != me:
                        K.check(K.lib().pxb_enable_peer_access(int(args[6])), "pxb_enable_peer_access")
                        args[6] = me
                    out[k] = fn(*args)
                opened[r] = out
            return opened[r]

        self.lo, self.hi = open_rank(lo), open_rank(hi)


def peer_buffers(key, make, rank, world, lo, hi):
    """Pooled PeerBuffers, or None when peer memory cannot be set up on EVERY rank (the decision is agreed on collectively)."""
    if key in _PEER_POOL:
        return _PEER_POOL[key]
    ok, pb = 1, None
    try:
        pb = PeerBuffers(make, rank, world, lo, hi)
    except Exception as e:  # an allocator that cannot export IPC handles, no peer access between the two GPUs, ...
        import warnings

        warnings.warn(f"pyxu_b200.slab: peer-memory halo exchange unavailable ({type(e).__name__}: {e}); using NCCL send/recv")
        ok = 0
    flag = torch.tensor([ok], device=A.current_device(), dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None
    _PEER_POOL[key] = pb
    return pb


def _copy_params(p):
    q = K.PdsParams.from_buffer_copy(p)
    return q


class _Engine:
    """What the two engines share: the slab's geometry, the exchanger, plane-addressed pointers."""

    def _geometry(self, Kop, rank, world, halo, group, periodic=False):
        self.K = Kop
        self.shape = tuple(int(s) for s in Kop.arg_shape)
        assert len(self.shape) == 3, "slab decomposition: 3-D volumes"
        self.rank, self.world = rank, world
        self.group = group if group is not None else _high_priority_group()
        self.dev = A.current_device()
        self.hx = HaloExchanger(rank, world, self.group, periodic=periodic)
        self.start, self.stop = partition(self.shape[0], world)[rank]
        self.n0 = self.stop - self.start
        self.plane = self.shape[1] * self.shape[2]
        self.local_voxels = self.n0 * self.plane
        self.H = halo
        self._alloc = self.n0 + 2 * halo
        self.probe = None  # optional callable(tag) invoked around the kernels (bench.py records CUDA events with it)

    def _field(self, ncomp=1):
        return torch.zeros((ncomp, self._alloc, *self.shape[1:]), dtype=self.dtype, device=self.dev)

    @staticmethod
    def _wait(reqs):
        for r in reqs:
            r.wait()

    def _p(self, t, comp, plane=0):
        """Device pointer to owned plane `plane` of component `comp`."""
        return C.c_void_p(t.data_ptr() + t.element_size() * ((comp * t.shape[1] + self.H + plane) * self.plane))

    def _sub_slab(self, p0, p1):
        """Slab flags of a launch restricted to owned planes [p0, p1): an interior cut is an open side."""
        lo = 1 if (p0 > 0 or self.hx.lo is not None) else 0
        hi = 1 if (p1 < self.n0 or self.hx.hi is not None) else 0
        return K.Slab(lo, hi, self.H, self._alloc)

    def _own(self, t, comp=None):
        """Owned planes of a field as a tensor view: (n0, n1, n2) of one component, or (ncomp, n0, n1, n2)."""
        v = t[:, self.H : self.H + self.n0]
        return v if comp is None else v[comp]

    def _tick(self, tag):
        if self.probe is not None:
            self.probe(tag)


class SlabTV(_Engine):
    """PD3O / CondatVu on  min f(x) + g(x) + h(grad x)  with f pointwise (null or alpha*||x + c||^2), for a volume
    decomposed in z-slabs.  Same kernels and the same parameter block as the single-domain fused path of
    pyxu_b200.opt.solver (pxb_pds_iter, or pxb_pds_primal + pxb_pds_dual outside its envelope), issued per rank on
    its slab with `open_lo / open_hi` set where a neighbour exists.

    algo    K.ALGO_PD3O | K.ALGO_CV
    Kop     first-order Gradient over the GLOBAL volume
    params  K.PdsParams filled by the solver's planner (step sizes, g, f kind / alpha, h); the shift pointers are set here
    x0      (n0, n1, n2): this rank's planes of the initial iterate (device)
    z0      None (-> K x0) | (3, n0, n1, n2)
    shift   None | this rank's planes of the data-term shift c (n0 * plane samples, device) | a 1-sample tensor
    """

    def __init__(self, algo, Kop, params, x0, z0=None, shift=None, rank=0, world=1, group=None, overlap=True, fused=True, edge=8, p2p=None,
                 dtype=None, shift_streamed=False):
        """x0 = None (with `dtype`; world 1 only): the initial iterate -- and, with `shift_streamed`, the per-voxel shift of the data
        term -- is still in host memory and arrives through run_streamed()."""
        A.require_cuda()
        self.algo, self.dtype = algo, (x0.dtype if x0 is not None else dtype)
        assert x0 is not None or (world == 1 and z0 is None), "deferred initial iterate: single z-slab, z0 = K x0"
        modes = tuple(Kop._mode)
        self._geometry(Kop, rank, world, 1, group, periodic=(modes[0] == "wrap"))
        h, n0 = self.H, self.n0
        assert n0 >= 2 * h + 1, "slabs thinner than 3 planes are not supported"
        self.p = _copy_params(params)
        self._shift_arr = bool(shift_streamed) or (shift is not None and shift.numel() > 1)
        if self._shift_arr and not shift_streamed:
            assert shift.numel() == self.local_voxels, "data-term shift: one sample per voxel of the slab"
        self.fused = bool(fused)  # single-kernel iteration with ping-pong (primal, z) pairs; decided for good on the first step
        self.edge = max(1, min(int(edge), n0 // 2))  # planes of the boundary launches that precede the exchange
        own = slice(h, h + n0)
        self.x = self._field() if algo == K.ALGO_PD3O else None  # PD3O's x (CondatVu's primal variable IS x)
        self._x_stale = False
        self._iter_cache, self._fn_iter = {}, K.lib().pxb_pds_iter
        # Halo exchange fused into the kernel through peer memory (pxb_pds_iter_p2p) when every rank can map its neighbours'
        # buffers; NCCL send/recv on a side stream otherwise (PYXU_B200_SLAB_P2P=0 forces the latter for A/B runs).
        self.p2p, self._epoch, self._peers = None, 0, {}
        if self.fused and world > 1 and p2p is not False and _p2p_enabled():
            make = lambda: dict(p0=self._field(), p1=self._field(), z0=self._field(3), z1=self._field(3),
                                flags=torch.zeros(4, dtype=torch.int32, device=self.dev))
            key = ("tv", self.shape, self.dtype, rank, world, bool(self.hx.periodic))
            reused = key in _PEER_POOL
            self.p2p = peer_buffers(key, make, rank, world, self.hx.lo, self.hx.hi)
            if self.p2p is not None:
                # nobody may still be writing into pooled buffers (a neighbour's last iteration of an earlier solve), and
                # nobody may start before every rank has reset its counters
                torch.cuda.synchronize()
                dist.barrier()
                b = self.p2p.bufs
                if reused:
                    for t in b.values():
                        t.zero_()
                self._pb, self._zb, self._flags = [b["p0"], b["p1"]], [b["z0"], b["z1"]], b["flags"]
        if self.p2p is None:
            self._pb = [self._field(), self._field()]
            self._zb = [self._field(3), self._field(3)]
        self.cur = 0
        self.w = None
        self.shift_h = None
        if self._shift_arr:  # shift WITH ghost planes: the single-kernel form evaluates grad f on the ghost plane too
            self.shift_h = self._field()
            if not shift_streamed:
                self.shift_h[0, own].copy_(shift.reshape(n0, *self.shape[1:]))
        self._shift1 = shift if (shift is not None and not self._shift_arr) else None
        self._streams = None
        if x0 is None:  # run_streamed() fills the fields chunk by chunk
            self.overlap, self.comm = False, None
            return
        self._pb[0][0, own].copy_(x0.reshape(n0, *self.shape[1:]))
        if self.x is not None:
            self.x[0, own].copy_(self._pb[0][0, own])
        # Sub-range launches treat the cut as an open side, which is only exact when no boundary fold reaches across
        # it: 'constant' (nothing folds) or 'wrap' (ring exchange: every side is open).  Other modes along z run
        # whole-slab launches after the exchange.
        self.overlap = bool(overlap) and world > 1 and modes[0] in ("constant", "wrap")
        # The exchange must not queue behind the interior kernel's ~10^5 pending CTAs: measured on 8 GPUs, with default
        # priorities the NCCL send/recv kernel only ran once the interior grid had drained (0.25 ms of an 1.04 ms iteration
        # exposed).  Side stream and NCCL stream are therefore high-priority: their CTAs are placed as soon as a CTA
        # of the interior kernel retires (every ~16 us).
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1) if world > 1 else None
        if z0 is None:  # z0 = K x0 needs x0's upper ghost plane
            self._wait(self.hx.exchange(self._pb[0][0], h, n0, up=False, down=True))
            d = self._desc(0, n0)
            K.check(K.lib().pxb_gradient_apply(C.byref(d), self._p(self._pb[0], 0), self._p(self._zb[0], 0), A.stream()), "gradient_apply")
        else:
            self._zb[0][:, own].copy_(z0.reshape(3, n0, *self.shape[1:]))
        if world > 1:  # ghost planes of the data term (once) and of the initial iterate
            if self.shift_h is not None:
                self._wait(self.hx.exchange(self.shift_h[0], h, n0, up=True, down=True))
            self._wait(self._exchange_state(self.cur))
        if self.p2p is not None:  # every rank's counters are zero and its ghost planes filled before anybody iterates
            torch.cuda.synchronize()
            dist.barrier()

    # current iterate (the ping-pong index flips every single-kernel iteration)
    @property
    def primal(self):
        return self._pb[self.cur]

    @property
    def z(self):
        return self._zb[self.cur]

    def _desc(self, p0, p1):
        """Gradient descriptor for owned planes [p0, p1) of this slab."""
        return self.K._desc(1, A.dcode(self._pb[0]), slab=self._sub_slab(p0, p1), shape0=p1 - p0)

    def _params(self, p0, ghost):
        """Parameter block of a launch starting at owned plane p0: the shift is addressed like the primal variable
        (`ghost`: through the copy that carries ghost planes -- the single-kernel form reads it on the ghost plane too)."""
        p = _copy_params(self.p)
        if self._shift_arr:
            sh = self.shift_h
            p.f.shift = sh.data_ptr() + sh.element_size() * (self.H + p0) * self.plane
            p.f.shift_period = sh.shape[1] * self.plane if ghost else (self.n0 - p0) * self.plane
        elif self._shift1 is not None:
            p.f.shift, p.f.shift_period = self._shift1.data_ptr(), 1
        return p

    # -- two-sweep form ----------------------------------------------------------------------------------
    def _primal(self, p0, p1, norms=None):
        d, p = self._desc(p0, p1), self._params(p0, False)
        if self.w is None:
            self.w = self._field()
        xo = self._p(self.x, 0, p0) if self.x is not None else None
        rc = K.lib().pxb_pds_primal(self.algo, C.byref(d), C.byref(p), self._p(self.primal, 0, p0), self._p(self.z, 0, p0), None,
                                    xo, self._p(self.w, 0, p0), A.ptr(norms), A.stream())
        K.check(rc, "pxb_pds_primal")

    def _dual(self, p0, p1, norms=None):
        d, p = self._desc(p0, p1), self._params(p0, False)
        rc = K.lib().pxb_pds_dual(C.byref(d), C.byref(p), self._p(self.w, 0, p0), self._p(self.z, 0, p0), A.ptr(norms), A.stream())
        K.check(rc, "pxb_pds_dual")

    # -- single-kernel form ----------------------------------------------------------------------------
    def _exchange_state(self, idx):
        """Boundary planes of iterate `idx` -> the neighbours' ghost planes: primal, z_0..2 first plane down, z_0 last plane up."""
        u, z = self._pb[idx], self._zb[idx]
        items = [(u[0], False, True), (z[0], True, True), (z[1], False, True), (z[2], False, True)]
        return self.hx.exchange_many(items, self.H, self.n0)

    def _iter(self, p0, p1, src, dst, x_out, nx, nz):
        """pxb_pds_iter on owned planes [p0, p1): reads iterate `src` (ghost planes valid), writes iterate `dst`.
        Returns False when the single-kernel form does not apply (nothing was launched)."""
        key = (p0, p1, src)
        c = self._iter_cache.get(key)
        if c is None:  # descriptors and pointers never change: build the ctypes arguments once (host time matters at 1 ms/iteration)
            d, p = self._desc(p0, p1), self._params(p0, True)
            ptr = lambda t, comp: self._p(t, comp, p0)
            c = (d, p, C.byref(d), C.byref(p), ptr(self._pb[src], 0), ptr(self._zb[src], 0), ptr(self._pb[dst], 0), ptr(self._zb[dst], 0),
                 ptr(self.x, 0) if self.x is not None else None)
            self._iter_cache[key] = c
        rc = self._fn_iter(self.algo, c[2], c[3], c[4], c[5], c[6], c[7], c[8] if x_out else None, A.ptr(nx), A.ptr(nz), A.stream())
        if rc == -3 and self.fused is True and not self._iter_cache.get("ran"):
            return False
        if rc:
            K.check(rc, "pxb_pds_iter")
        self._iter_cache["ran"] = True
        return True

    # -- wavefront over z-chunks behind the upload (single z-slab) -----------------------------------------
    def run_streamed(self, x0_host, shift_host, n_iter, nrm=None, use_x=False, use_z=False, out=None, planes=16):
        """Uploads the initial iterate (and the data-term shift) chunk by chunk along axis 0 and runs the first `n_iter` iterations
        as a WAVEFRONT behind the upload: iteration i of chunk c only reads iterate i of chunks c-1, c, c+1 (one plane of each
        neighbour), so it is queued as soon as chunk c+i+2 has arrived -- the iterations hide under the PCIe transfer instead of
        waiting for its end.  With `out` (pinned host tensor, one sample per voxel) the result of the last iteration is copied
        back chunk by chunk behind the wave, i.e. while later chunks are still being uploaded and iterated.

        Task (i, c), i = -1 .. n_iter-1 (i = -1: z0 = K x0 on the chunk), is issued in the order of i + c (ascending i inside a
        group), which respects every dependency -- (i, c) needs (i-1, c-1), (i-1, c), (i-1, c+1) -- on ONE in-order compute stream;
        the ping-pong pair a task overwrites holds iterate i-1, whose last readers are exactly those three tasks.

        x0_host, shift_host : flat host arrays of the slab's dtype (shift_host None: the shift is not per-voxel or is on the device)
        nrm                 : (n_iter, 2, 1, 2) zeroed device doubles -- RelError sums of every iteration ([:, 0] x when use_x,
                              [:, 1] z when use_z) -- or None
        Returns the number of iterations carried out: n_iter, or 0 when the single-kernel form declined the problem (everything
        is uploaded and z0 initialised then; the caller iterates with step())."""
        from concurrent.futures import ThreadPoolExecutor

        assert self.world == 1 and self.cur == 0
        h, n0, plane = self.H, self.n0, self.plane
        lib = K.lib()
        main = torch.cuda.current_stream()
        if self._streams is None:
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        up, dn = self._streams
        up.wait_stream(main)  # (the fields were zero-filled on the main stream)
        # (quarter-size chunks for the last n_iter + 2 chunks -- the triangle of tasks still pending when the upload ends -- were tried:
        # no gain, 0.230 against 0.226 s at 1024^3, K = 20)
        bounds = [(p, min(n0, p + planes)) for p in range(0, n0, planes)]
        if len(bounds) > 1 and bounds[-1][1] - bounds[-1][0] < 2:  # no one-plane tail
            bounds[-2:] = [(bounds[-2][0], n0)]
        marks = self._stream_marks = [] if self.probe is not None else None  # (tag, CUDA event) for tools/probe_stream.py

        def mark(tag, stream=None):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream) if stream is not None else ev.record()
                marks.append((tag, ev))

        mark("begin")
        nchunk = len(bounds)
        pinned = A.is_pinned(x0_host) and (shift_host is None or A.is_pinned(shift_host))
        pool = None if pinned else ThreadPoolExecutor(A._copy_threads())
        write_last = True  # PD3O: the last iteration of the wave writes x (4 B/voxel, once) -- no rebuild from the previous pair later
        declined = [n_iter <= 0]
        gdesc = {}

        def upload(c):
            p0, p1 = bounds[c]
            with torch.cuda.stream(up):
                A.h2d_into(self._pb[0][0, h + p0 : h + p1], x0_host[p0 * plane : p1 * plane], pinned, pool)
                if shift_host is not None:
                    A.h2d_into(self.shift_h[0, h + p0 : h + p1], shift_host[p0 * plane : p1 * plane], pinned, pool)
                ev = torch.cuda.Event()
                ev.record()
                if c == nchunk - 1:
                    mark("upload_end")
            main.wait_event(ev)

        def task(i, c):
            p0, p1 = bounds[c]
            if i < 0:
                d = gdesc.get(c)
                if d is None:
                    d = gdesc[c] = self._desc(p0, p1)
                K.check(lib.pxb_gradient_apply(C.byref(d), self._p(self._pb[0], 0, p0), self._p(self._zb[0], 0, p0), A.stream()), "gradient_apply")
                if self.x is not None and use_x:  # RelError[x] of the first iteration compares with x0
                    self.x[0, h + p0 : h + p1].copy_(self._pb[0][0, h + p0 : h + p1])
                return
            if declined[0]:
                return
            last = i == n_iter - 1
            nx = nrm[i, 0] if (nrm is not None and use_x) else None
            nz = nrm[i, 1] if (nrm is not None and use_z) else None
            if not self._iter(p0, p1, i % 2, 1 - i % 2, use_x or (last and write_last), nx, nz):
                declined[0] = True  # outside the single-kernel envelope (decided by the first launch): upload only from here on
                return
            if last and out is not None:
                ev = torch.cuda.Event()
                ev.record()
                dn.wait_event(ev)
                src = (self.x if self.x is not None else self._pb[n_iter % 2])[0, h + p0 : h + p1]
                with torch.cuda.stream(dn):
                    out[p0 * plane : p1 * plane].copy_(src.reshape(-1), non_blocking=True)

        def group(s):
            for i in range(-1, n_iter):
                c = s - i
                if 0 <= c < nchunk:
                    task(i, c)

        try:
            for j in range(nchunk):
                upload(j)
                if j < nchunk - 1:
                    group(j - 2)  # (needs chunks up to j: empty for j = 0)
                else:
                    for s in range(j - 2, n_iter + nchunk - 1):
                        group(s)
        finally:
            if pool is not None:
                pool.shutdown()
        mark("compute_end")
        if out is not None:
            mark("download_end", dn)
            main.wait_stream(dn)
        done = 0 if declined[0] else n_iter
        self.cur = done % 2
        if self.x is not None:
            self._x_stale = done > 0 and not (use_x or write_last)
        return done

    def reset_streamed(self):
        """Back to the state before run_streamed() (the host arrays are uploaded again by the next call)."""
        torch.cuda.synchronize()
        self.cur, self._x_stale = 0, False
        self._iter_cache.pop("ran", None)

    def _peer_block(self, dst):
        """pxb_peer for an iteration that writes iterate `dst`: where this slab's new boundary planes go in the neighbours' copies
        of that iterate, and the counters on both sides."""
        pr = self._peers.get(dst)
        if pr is None:
            es, H, plane = self._pb[0].element_size(), self.H, self.plane
            parts = partition(self.shape[0], self.world)
            pr = K.Peer()
            mine = self._flags.data_ptr()
            if self.hx.lo is not None:
                nb = self.p2p.lo
                n_lo = parts[self.hx.lo][1] - parts[self.hx.lo][0]
                up, zz = nb["p%d" % dst], nb["z%d" % dst]
                pr.dn_u = up.data_ptr() + es * (H + n_lo) * plane
                pr.dn_z = zz.data_ptr() + es * (H + n_lo) * plane
                pr.dn_zvol = zz.shape[1] * plane
                pr.dn_flag = nb["flags"].data_ptr() + 4  # I am its upper neighbour
                pr.lo_wait = mine
            if self.hx.hi is not None:
                nb = self.p2p.hi
                pr.up_z0 = nb["z%d" % dst].data_ptr() + es * (H - 1) * plane
                pr.up_flag = nb["flags"].data_ptr()      # I am its lower neighbour
                pr.hi_wait = mine + 4
            self._peers[dst] = pr
        pr.epoch = self._epoch
        return pr

    def _step_p2p(self, nx, nz, want_x):
        """One launch: the kernel stores the new boundary planes into the neighbours' ghost planes itself (pxb_pds_iter_p2p)."""
        src, dst = self.cur, 1 - self.cur
        key = ("p2p", src)
        c = self._iter_cache.get(key)
        if c is None:
            d, p = self._desc(0, self.n0), self._params(0, True)
            ptr = lambda t, comp: self._p(t, comp, 0)
            c = (d, p, C.byref(d), C.byref(p), ptr(self._pb[src], 0), ptr(self._zb[src], 0), ptr(self._pb[dst], 0), ptr(self._zb[dst], 0),
                 ptr(self.x, 0) if self.x is not None else None)
            self._iter_cache[key] = c
        pr = self._peer_block(dst)
        self._tick("iter_begin")
        rc = K.lib().pxb_pds_iter_p2p(self.algo, c[2], c[3], c[4], c[5], c[6], c[7], c[8] if want_x else None, A.ptr(nx), A.ptr(nz), C.byref(pr), A.stream())
        if rc == -3 and self._epoch == 0:
            return False
        if rc:
            K.check(rc, "pxb_pds_iter_p2p")
        self._tick("iter_end")
        self._epoch += 1
        return True

    def _step_fused(self, nx, nz, want_x):
        n0, e = self.n0, self.edge
        src, dst = self.cur, 1 - self.cur
        want_x = bool(want_x) and self.x is not None
        if want_x and nx is not None and self._x_stale:
            self.materialize_x()  # RelError[x] compares with the previous x
        if self.p2p is not None:
            if self._step_p2p(nx, nz, want_x):
                self.cur = dst
                if self.x is not None:
                    self._x_stale = not want_x
                return True
            self.p2p = None  # the kernel declined (outside the TMA form's envelope): NCCL exchange from here on, same buffers
        main = torch.cuda.current_stream()
        self._tick("iter_begin")
        if self.world > 1 and self.overlap and n0 >= 4 * e:
            lo, hi = (e if self.hx.lo is not None else 0), (n0 - e if self.hx.hi is not None else n0)
            if lo and not self._iter(0, lo, src, dst, want_x, nx, nz):
                return False
            if hi < n0 and not self._iter(hi, n0, src, dst, want_x, nx, nz):
                return False
            self._tick("edges_end")
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                self._tick("exchange_begin")
                self._wait(self._exchange_state(dst))  # the new boundary planes travel while the interior is computed
                self._tick("exchange_end")
            if not self._iter(lo, hi, src, dst, want_x, nx, nz):
                return False
            self._tick("interior_end")
        else:
            if not self._iter(0, n0, src, dst, want_x, nx, nz):
                return False
            if self.world > 1:
                self.comm.wait_stream(main)
                with torch.cuda.stream(self.comm):
                    self._wait(self._exchange_state(dst))
        if self.world > 1:
            main.wait_stream(self.comm)  # the next iteration reads the ghost planes this exchange fills
        self._tick("iter_end")
        self.cur = dst
        if self.x is not None:
            self._x_stale = not want_x
        return True

    def materialize_x(self):
        """PD3O: x_k = prox_g(u_{k-1} - tau K^T z_{k-1}) from the previous iterate (kept in the other ping-pong pair)."""
        if not (self.x is not None and self._x_stale):
            return
        prev = 1 - self.cur
        tmp_u, tmp_w = self._pb[prev].clone(), torch.empty_like(self._pb[prev])
        d, p = self._desc(0, self.n0), self._params(0, False)
        rc = K.lib().pxb_pds_primal(self.algo, C.byref(d), C.byref(p), self._p(tmp_u, 0), self._p(self._zb[prev], 0), None,
                                    self._p(self.x, 0), self._p(tmp_w, 0), None, A.stream())
        K.check(rc, "pxb_pds_primal")
        self._x_stale = False

    # -- one iteration ---------------------------------------------------------------------------------
    def step(self, nx=None, nz=None, want_x=True):
        """nx, nz: (1, 2) device double buffers the kernels accumulate this rank's RelError sums into (or None)."""
        if self.fused:
            if self._step_fused(nx, nz, want_x):
                return
            self.fused = False  # outside the envelope of pxb_pds_iter (decided on the first step): two sweeps from now on
        h, n0 = self.H, self.n0
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

    # -- results (this rank's planes) ------------------------------------------------------------------
    def x_local(self):
        if self.x is None:
            return self._own(self.primal, 0)
        self.materialize_x()
        return self._own(self.x, 0)

    def z_local(self):
        return self._own(self.z)


class SlabDeblurCV(_Engine):
    """CondatVu on  min alpha*||A x + c||^2 + g(x) + h(grad x)  for a volume decomposed in z-slabs
    (configs[4] of BASELINE.json: 3-D TV deblurring with a separable Stencil PSF).

    A is a `constant`-mode separable Stencil (its factor along axis 0 reaches H planes across a cut), the Gradient is
    the forward-difference stack.  One iteration on every rank (reference iteration: pds.py:429-442):
        r   = 2 alpha (A x + c)     (one marching pass, or axis-0 streaming pass + tiled in-plane pass; reads H ghost planes of x)
                                                                                          -> r ghosts exchanged (H planes)
        grad f = A^T r                                                                    -> first plane of grad f goes down
        (x, z) <- pxb_pds_iter(CV, grad f array)   (single kernel, ping-pong)             -> x ghosts (H planes), z planes
    Exchanges are NCCL send/recv batches on a high-priority stream.
    """

    def __init__(self, Kop, Aop, alpha, params, x0, z0=None, shift=None, rank=0, world=1, group=None, overlap=True):
        A.require_cuda()
        self.dtype = x0.dtype
        self.Aop = Aop
        kshape = [int(k.size) for k in Aop._kernels] if Aop._separable else list(Aop._kernels[0].shape)
        cen = [int(c[i]) for i, c in enumerate(Aop._centers)] if Aop._separable else [int(v) for v in Aop._centers[0]]
        H = max(1, cen[0], kshape[0] - 1 - cen[0])
        self._geometry(Kop, rank, world, H, group)
        n0 = self.n0
        assert n0 >= H, "slabs thinner than the PSF's reach along z are not supported"
        self.p = _copy_params(params)
        self._xb, self._zb, self.cur = [self._field(), self._field()], [self._field(3), self._field(3)], 0
        self.r, self.garr, self.tmp = self._field(), self._field(), None
        self.two_alpha = 2.0 * float(alpha)
        self.add = None
        if shift is not None:
            self.add = shift.reshape(-1).contiguous()
            assert self.add.numel() in (1, self.local_voxels)
        own = slice(H, H + n0)
        self._xb[0][0, own].copy_(x0.reshape(n0, *self.shape[1:]))
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1) if world > 1 else None
        Aop._slab_dcode = A.dcode(self.r)
        self._slab = K.Slab(1 if self.hx.lo is not None else 0, 1 if self.hx.hi is not None else 0, H, self._alloc)
        self._cache = {}   # descriptors per (kind, adjoint / parity, p0, p1): built once
        self._plans = {adj: Aop._tiled_plan(adj) for adj in (False, True)}
        # a dense PSF of full rank: the marching kernel reads the ghost planes itself (pxb_stencil3d_dense_apply)
        from .operator.linop import stencil as _st

        self.dense = (_st.DENSE3D_MARCH and not all(p is not None for p in self._plans.values())
                      and Aop._desc3d_dense(A.dcode(self.r), False, 1, slab=self._slab, shape0=n0) is not None)
        if not self.dense and not all(p is not None and p[0] is not None for p in self._plans.values()):
            raise NotImplementedError("slab decomposition of a Stencil data term: expected a 'constant'-mode PSF that is separable with a factor along "
                                      "axis 0, or dense with at most 7 taps per axis (PYXU_B200_DENSE3D_MARCH)")
        self.single_pass = not self.dense and Aop._desc3d(A.dcode(self.r), False, 1, slab=self._slab, shape0=n0) is not None
        self.edge = max(H, min(8, n0 // 4))  # planes of the boundary launches that precede each exchange
        self.overlap = bool(overlap) and world > 1 and n0 >= 4 * self.edge
        self._gdesc = Kop._desc(1, A.dcode(self.r), slab=self._slab, shape0=n0)
        self._wait(self.hx.exchange_planes([(self._xb[0][0], H, H)], H, n0))
        if z0 is None:  # z0 = K x0 needs x0's ghost planes
            K.check(K.lib().pxb_gradient_apply(C.byref(self._gdesc), self._p(self._xb[0], 0), self._p(self._zb[0], 0), A.stream()), "gradient_apply")
        else:
            self._zb[0][:, own].copy_(z0.reshape(3, n0, *self.shape[1:]))
        self._wait(self.hx.exchange_planes([(self._zb[0][0], 1, 1), (self._zb[0][1], 0, 1), (self._zb[0][2], 0, 1)], H, n0))

    @property
    def primal(self):
        return self._xb[self.cur]

    @property
    def z(self):
        return self._zb[self.cur]

    def _stencil(self, adjoint, src, dst, p0=0, p1=None):
        """dst[p0:p1] = (2 alpha (A src + c) | A^T src)[p0:p1], reading src's neighbouring / ghost planes: one marching pass
        (pxb_stencil3d_apply), or axis-0 streaming pass + tiled in-plane pass outside its envelope."""
        p1 = self.n0 if p1 is None else p1
        key = ("st", adjoint, p0, p1)
        c = self._cache.get(key)
        if c is None:
            slab = self._sub_slab(p0, p1)
            add = None
            if not adjoint and self.add is not None:
                add = self.add if self.add.numel() == 1 else self.add[p0 * self.plane : p1 * self.plane]
            al, be = (1.0, 0.0) if adjoint else (self.two_alpha, self.two_alpha if add is not None else 0.0)
            if self.dense:
                c = self._cache[key] = (self.Aop._desc3d_dense(A.dcode(self.r), adjoint, 1, al, be, add, slab=slab, shape0=p1 - p0), None, slab, None)
                return self._stencil(adjoint, src, dst, p0, p1)
            d3 = self.Aop._desc3d(A.dcode(self.r), adjoint, 1, al, be, add, slab=slab, shape0=p1 - p0) if self.single_pass else None
            d2 = None
            if d3 is None:
                own = slice(self.H + p0, self.H + p1)
                d2, _ = self.Aop._tiled_desc(self.r[0, own], adjoint, alpha=al, beta=be, add=add)
                d2.nimg = p1 - p0
                if self.tmp is None:
                    self.tmp = self._field()
            c = self._cache[key] = (d3, d2, slab, self._plans[adjoint][0])
        d3, d2, slab, axis0 = c
        if self.dense:
            K.check(K.lib().pxb_stencil3d_dense_apply(C.byref(d3), self._p(src, 0, p0), self._p(dst, 0, p0), A.stream()), "pxb_stencil3d_dense_apply")
            return
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
            p = _copy_params(self.p)
            p.f.kind, p.f.garr = K.F_GRADARR, self._p(self.garr, 0, p0).value
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
                self._wait(self.hx.exchange_planes(exchange_items, self.H, n0))
            fn(lo, hi)
            main.wait_stream(self.comm)
        else:
            fn(0, n0)
            if self.world > 1:
                self._wait(self.hx.exchange_planes(exchange_items, self.H, n0))

    def step(self, nx=None, nz=None, want_x=True):
        H = self.H
        src, dst = self.cur, 1 - self.cur
        xs, xd, zd = self._xb[src], self._xb[dst], self._zb[dst]
        self._tick("iter_begin")
        self._staged(lambda a, b: self._stencil(False, xs, self.r, a, b), [(self.r[0], H, H)])          # r = 2 alpha (A x + c)
        self._staged(lambda a, b: self._stencil(True, self.r, self.garr, a, b), [(self.garr[0], 0, 1)])  # grad f = A^T r
        self._staged(lambda a, b: self._iter(src, dst, nx, nz, a, b),
                     [(xd[0], H, H), (zd[0], 1, 1), (zd[1], 0, 1), (zd[2], 0, 1)])                        # (x, z) <- CV iteration
        self._tick("iter_end")
        self.cur = dst

    def x_local(self):
        return self._own(self.primal, 0)

    def z_local(self):
        return self._own(self.z)


def gather_planes(local, vol_shape, world, group=None):
    """Full (comps, N0, n1, n2) array on every rank from every rank's (comps, n0_local, n1, n2) planes."""
    parts = partition(vol_shape[0], world)
    comps = local.shape[0]
    nmax = max(b - a for a, b in parts)
    mine = torch.zeros((comps, nmax, *vol_shape[1:]), dtype=local.dtype, device=local.device)
    mine[:, : local.shape[1]].copy_(local)
    bufs = [torch.empty_like(mine) for _ in parts]
    dist.all_gather(bufs, mine, group=group)
    return torch.cat([b_[:, : e - a] for b_, (a, e) in zip(bufs, parts)], dim=1)
