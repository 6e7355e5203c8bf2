"""
CPU coverage of the multi-GPU path:
 (1) kernel-side slab semantics (open sides, ghost planes, sub-range launches) -- the kernel bodies compiled
     for the host (tests/emu) run a z-slab-decomposed PD3O-TV iteration with the halo exchange done in NumPy,
     and must reproduce the single-domain iterates bit for bit;
 (2) host-side plumbing -- partition() and HaloExchanger over torch.distributed (gloo, world_size 2 and 3).
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_util as E
import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K
from pyxu_b200.slab import HaloExchanger, partition


def test_partition():
    assert partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert partition(1024, 8)[-1] == (896, 1024)
    for n, w in ((17, 4), (5, 5), (100, 7)):
        p = partition(n, w)
        assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))
        assert max(e - s for s, e in p) - min(e - s for s, e in p) <= 1


def _single_domain(shape, y, mode, n_iter, tau, sigma, rho, lam):
    Kop = pxo.Gradient(arg_shape=shape, mode=mode)
    shift = np.ascontiguousarray(-y.reshape(-1))
    P = E.pds_params(tau, sigma, rho, gspec=(K.PROX_POS, 0, 0), fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=lam)
    d = Kop._desc(1, K.F64)
    x = y.reshape(-1).copy()
    z = E.gradient_run(Kop, x, False)
    u, w = x.copy(), np.empty_like(x)
    for _ in range(n_iter):
        E.lib().emu_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(u), E.p(z), None, E.p(x), E.p(w), None)
        E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(z), None)
    return x.reshape(shape), z.reshape((3,) + shape)


class _Rank:
    """One simulated rank: same buffer layout / descriptors as pyxu_b200.slab.SlabPD3OTV."""

    H = 1

    def __init__(self, shape, y, a, b, rank, world, mode, periodic):
        self.n0, self.plane = b - a, shape[1] * shape[2]
        self.a, self.b, self.rank, self.world = a, b, rank, world
        alloc = self.n0 + 2 * self.H
        self.K = pxo.Gradient(arg_shape=shape, mode=mode)
        self.has_lo = rank > 0 or periodic
        self.has_hi = rank < world - 1 or periodic
        f = lambda c=1: np.zeros((c, alloc) + shape[1:])
        self.u, self.x, self.w, self.z = f(), f(), f(), f(3)
        self.shift = np.ascontiguousarray(-y[a:b])
        self.u[0, 1:-1] = y[a:b]
        self.x[0, 1:-1] = y[a:b]

    def ptr(self, t, comp, plane):
        return C.c_void_p(t.ctypes.data + 8 * ((comp * t.shape[1] + self.H + plane) * self.plane))

    def desc(self, p0, p1):
        lo = 1 if (p0 > 0 or self.has_lo) else 0
        hi = 1 if (p1 < self.n0 or self.has_hi) else 0
        return self.K._desc(1, K.F64, slab=K.Slab(lo, hi, self.H, self.n0 + 2 * self.H), shape0=p1 - p0)

    def params(self, tau, sigma, rho, lam, p0):
        P = E.pds_params(tau, sigma, rho, gspec=(K.PROX_POS, 0, 0), fkind=K.F_SQL2, alpha=0.5, hkind=K.DUAL_L21, lam=lam)
        P.f.shift = self.shift.ctypes.data + 8 * p0 * self.plane
        P.f.shift_period = (self.n0 - p0) * self.plane
        return P

    fast = 0  # 0: generic bodies; 1: vectorised fast bodies (VEC=1; the shape's row length is odd)

    def primal(self, p0, p1, prm):
        d, P = self.desc(p0, p1), self.params(*prm, p0)
        a = (self.ptr(self.u, 0, p0), self.ptr(self.z, 0, p0))
        b = (self.ptr(self.x, 0, p0), self.ptr(self.w, 0, p0))
        if self.fast:
            assert E.lib().emu_tv_fast(self.fast, 0, K.ALGO_PD3O, C.byref(d), C.byref(P), *a, *b, None) == 0
        else:
            E.lib().emu_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(P), *a, None, *b, None)

    def dual(self, p0, p1, prm):
        d, P = self.desc(p0, p1), self.params(*prm, p0)
        if self.fast:
            assert E.lib().emu_tv_fast(self.fast, 1, 0, C.byref(d), C.byref(P), None, self.ptr(self.z, 0, p0), None, self.ptr(self.w, 0, p0), None) == 0
        else:
            E.lib().emu_pds_dual(C.byref(d), C.byref(P), self.ptr(self.w, 0, p0), self.ptr(self.z, 0, p0), None)


def _exchange(ranks, name, comp, up, down, periodic):
    W = len(ranks)
    for r, rk in enumerate(ranks):
        arr = getattr(rk, name)[comp]
        if up:
            dst = r + 1 if r + 1 < W else (0 if periodic else None)
            if dst is not None:
                getattr(ranks[dst], name)[comp][0] = arr[rk.n0]  # last owned plane -> lower ghost
        if down:
            dst = r - 1 if r > 0 else (W - 1 if periodic else None)
            if dst is not None:
                getattr(ranks[dst], name)[comp][-1] = arr[1]  # first owned plane -> upper ghost


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("mode,subrange", [("constant", True), ("constant", False), ("wrap", True),
                                           (("reflect", "wrap", "symmetric"), False), (("edge", "constant", "reflect"), False)])
def test_slab_decomposed_iteration_equals_single_domain(world, mode, subrange, fast):
    shape, n_iter, lam = (11, 6, 7), 12, 0.08
    tau = sigma = 0.28
    rho = 1.3
    y = np.random.default_rng(0).random(shape)
    x_ref, z_ref = _single_domain(shape, y, mode, n_iter, tau, sigma, rho, lam)

    periodic = (mode if isinstance(mode, str) else mode[0]) == "wrap"
    parts = partition(shape[0], world)
    ranks = [_Rank(shape, y, a, b, r, world, mode, periodic) for r, (a, b) in enumerate(parts)]
    for rk in ranks:
        rk.fast = fast
    prm = (tau, sigma, rho, lam)
    _exchange(ranks, "x", 0, False, True, periodic)
    for rk in ranks:  # z0 = K x0
        d = rk.desc(0, rk.n0)
        E.lib().emu_gradient(C.byref(d), 0, rk.ptr(rk.x, 0, 0), rk.ptr(rk.z, 0, 0))
    for _ in range(n_iter):
        _exchange(ranks, "z", 0, True, False, periodic)
        for rk in ranks:
            if subrange:  # interior first, boundary plane after "the halo arrived" (same order as SlabPD3OTV.step)
                rk.primal(1, rk.n0, prm)
                rk.primal(0, 1, prm)
            else:
                rk.primal(0, rk.n0, prm)
        _exchange(ranks, "w", 0, False, True, periodic)
        for rk in ranks:
            if subrange:
                rk.dual(0, rk.n0 - 1, prm)
                rk.dual(rk.n0 - 1, rk.n0, prm)
            else:
                rk.dual(0, rk.n0, prm)
    x = np.concatenate([rk.x[0, 1:-1] for rk in ranks], axis=0)
    z = np.concatenate([rk.z[:, 1:-1] for rk in ranks], axis=1)
    if fast:  # fast bodies sum the taps in the same order: equal up to FMA contraction
        assert np.allclose(x, x_ref, rtol=1e-13, atol=1e-15) and np.allclose(z, z_ref, rtol=1e-13, atol=1e-15)
    else:
        assert np.array_equal(x, x_ref) and np.array_equal(z, z_ref)


# ---- torch.distributed plumbing over gloo --------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n0s = [e - s for s, e in partition(11, world)]
        n0, h = n0s[rank], 1
        buf = torch.full((n0 + 2 * h, 3, 4), -1.0, dtype=torch.float64)
        buf[h : h + n0] = torch.arange(n0, dtype=torch.float64).reshape(-1, 1, 1) + 100 * rank
        hx = HaloExchanger(periodic=periodic)
        for r in hx.exchange(buf, h, n0, up=True, down=True):
            r.wait()
        lo = rank - 1 if rank > 0 else (world - 1 if periodic else None)
        hi = rank + 1 if rank < world - 1 else (0 if periodic else None)
        ok = True
        ok &= bool((buf[0] == (-1.0 if lo is None else 100 * lo + n0s[lo] - 1)).all())
        ok &= bool((buf[-1] == (-1.0 if hi is None else 100 * hi)).all())
        ok &= bool((buf[h] == 100 * rank).all())  # owned planes untouched
        # one-directional exchange leaves the other ghost alone
        buf[0], buf[-1] = -2.0, -2.0
        for r in hx.exchange(buf, h, n0, up=False, down=True):
            r.wait()
        ok &= bool((buf[0] == -2.0).all()) and bool((buf[-1] == (-2.0 if hi is None else 100 * hi)).all())
        # the stopping criterion's single fused all-reduce
        v = torch.tensor([1.0 + rank, 2.0, 3.0, 4.0 * rank], dtype=torch.float64)
        dist.all_reduce(v)
        ok &= bool(torch.allclose(v, torch.tensor([world * (world + 1) / 2, 2.0 * world, 3.0 * world, 2.0 * world * (world - 1)], dtype=torch.float64)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic", [(2, False), (3, False), (3, True)])
def test_halo_exchanger_gloo(world, periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


# ---- single-kernel iteration on slabs (SlabPD3OTV fused path) -------------------------------------------
class _RankIter:
    """Simulated rank of the single-kernel slab path: ping-pong (u, z) pairs, shift stored WITH ghost planes."""

    H = 1

    def __init__(self, shape, y, a, b, rank, world, scheme, mode="constant", periodic=False):
        self.n0, self.plane = b - a, shape[1] * shape[2]
        self.has_lo, self.has_hi = rank > 0 or periodic, rank < world - 1 or periodic
        alloc = self.n0 + 2 * self.H
        self.K = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=mode)
        f = lambda c=1: np.zeros((c, alloc) + shape[1:])
        self.ub, self.zb, self.x, self.sh = [f(), f()], [f(3), f(3)], f(), f()
        self.cur = 0
        self.ub[0][0, 1:-1] = y[a:b]
        self.x[0, 1:-1] = y[a:b]
        self.sh[0, 1:-1] = -y[a:b]

    def ptr(self, t, comp, plane):
        return C.c_void_p(t.ctypes.data + 8 * ((comp * t.shape[1] + self.H + plane) * self.plane))

    def desc(self, p0, p1):
        lo = 1 if (p0 > 0 or self.has_lo) else 0
        hi = 1 if (p1 < self.n0 or self.has_hi) else 0
        return self.K._desc(1, K.F64, slab=K.Slab(lo, hi, self.H, self.n0 + 2 * self.H), shape0=p1 - p0)

    def iterate(self, p0, p1, prm, form, chunk):
        tau, sigma, rho, lam = prm
        d = self.desc(p0, p1)
        P = E.pds_params(tau, sigma, rho, gspec=(K.PROX_POS, 0, 0), fkind=K.F_SQL2, alpha=0.5, hkind=K.DUAL_L21, lam=lam)
        P.f.shift = self.sh.ctypes.data + 8 * (self.H + p0) * self.plane
        P.f.shift_period = self.sh.shape[1] * self.plane
        s, t = self.cur, 1 - self.cur
        fn = E.lib().emu_tv_iter if form == "direct" else E.lib().emu_tv_iter_tma
        rc = fn(K.ALGO_PD3O, C.byref(d), C.byref(P), self.ptr(self.ub[s], 0, p0), self.ptr(self.zb[s], 0, p0), self.ptr(self.ub[t], 0, p0),
                self.ptr(self.zb[t], 0, p0), self.ptr(self.x, 0, p0), None, None, chunk)
        assert rc == 0, rc


def _exchange_iter(ranks, idx, periodic=False):
    """u, z_0..2 first owned plane -> lower neighbour's upper ghost; z_0 last owned plane -> upper neighbour's lower ghost
    (periodic: a ring, as HaloExchanger(periodic=True) for 'wrap' along z)"""
    W = len(ranks)
    for r, rk in enumerate(ranks):
        if r > 0 or (periodic and W > 1):
            lo = ranks[(r - 1) % W]
            lo.ub[idx][0][-1] = rk.ub[idx][0][1]
            for c in range(3):
                lo.zb[idx][c][-1] = rk.zb[idx][c][1]
        if r + 1 < W or (periodic and W > 1):
            ranks[(r + 1) % W].zb[idx][0][0] = rk.zb[idx][0][rk.n0]


# 'constant' (the form SlabPD3OTV runs by default) and folding modes: in-plane folds on every rank, folds along z on the
# closed sides of the end ranks, 'wrap' along z as a ring of open sides
SLAB_ITER_MODES = ["constant", ("constant", "reflect", "wrap"), ("reflect", "symmetric", "edge"), ("edge", "wrap", "reflect"),
                   ("wrap", "reflect", "symmetric")]


@pytest.mark.parametrize("mode", SLAB_ITER_MODES, ids=lambda m: m if isinstance(m, str) else "-".join(m))
@pytest.mark.parametrize("form", ["direct", "tma"])
@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("scheme", ["forward"])
@pytest.mark.parametrize("shape", [(13, 6, 8), (9, 16, 64)], ids=["ragged", "fulltiles"])  # full tiles: folded rims served from the tile's boxes
def test_slab_single_kernel_iteration_equals_single_domain(world, form, scheme, mode, shape):
    n_iter, lam = (10 if shape[2] == 8 else 4), 0.08
    tau = sigma = 0.28
    rho = 1.3
    y = np.random.default_rng(1).random(shape)
    periodic = (mode if isinstance(mode, str) else mode[0]) == "wrap" and world > 1
    zfold = (mode if isinstance(mode, str) else mode[0]) not in ("constant", "wrap")
    # single-domain reference: two-sweep generic bodies
    Kop = pxo.Gradient(arg_shape=shape, scheme=scheme, mode=mode)
    shift = np.ascontiguousarray(-y.reshape(-1))
    P = E.pds_params(tau, sigma, rho, gspec=(K.PROX_POS, 0, 0), fkind=K.F_SQL2, alpha=0.5, shift=shift, hkind=K.DUAL_L21, lam=lam)
    d = Kop._desc(1, K.F64)
    x_ref = y.reshape(-1).copy()
    z_ref = E.gradient_run(Kop, x_ref, False)
    u_ref, w = x_ref.copy(), np.empty_like(x_ref)
    for _ in range(n_iter):
        E.lib().emu_pds_primal(K.ALGO_PD3O, C.byref(d), C.byref(P), E.p(u_ref), E.p(z_ref), None, E.p(x_ref), E.p(w), None)
        E.lib().emu_pds_dual(C.byref(d), C.byref(P), E.p(w), E.p(z_ref), None)

    parts = partition(shape[0], world)
    ranks = [_RankIter(shape, y, a, b, r, world, scheme, mode, periodic) for r, (a, b) in enumerate(parts)]
    prm = (tau, sigma, rho, lam)
    for r, rk in enumerate(ranks):  # ghost planes of x0 and of the shift; z0 = K x0
        if r > 0 or periodic:
            ranks[(r - 1) % world].x[0][-1] = rk.x[0][1]
            ranks[(r - 1) % world].sh[0][-1] = rk.sh[0][1]
        if r + 1 < world or periodic:
            ranks[(r + 1) % world].x[0][0] = rk.x[0][rk.n0]
            ranks[(r + 1) % world].sh[0][0] = rk.sh[0][rk.n0]
    for rk in ranks:
        dd = rk.desc(0, rk.n0)
        E.lib().emu_gradient(C.byref(dd), 0, rk.ptr(rk.x, 0, 0), rk.ptr(rk.zb[0], 0, 0))
        # the vectorised Gradient body (what pxb_gradient_apply launches for this descriptor) agrees on the slab: open sides
        # read their ghost planes, closed ones fold onto the slab's own planes
        zv = np.zeros_like(rk.zb[0])
        assert E.lib().emu_tv_grad(2, 0, C.byref(dd), rk.ptr(rk.x, 0, 0), rk.ptr(zv, 0, 0)) == 0
        assert np.allclose(zv[:, 1:-1], rk.zb[0][:, 1:-1], rtol=1e-14, atol=1e-15)
    _exchange_iter(ranks, 0, periodic)
    for it in range(n_iter):
        for rk in ranks:  # boundary chunks first, then the interior (same order as SlabPD3OTV._step_fused)
            e = 2
            lo, hi = (e if rk.has_lo else 0), (rk.n0 - e if rk.has_hi else rk.n0)
            if zfold:  # a fold along z must not meet a sub-range cut: whole-slab launches (SlabPD3OTV.overlap is off)
                lo, hi = 0, rk.n0
            if lo:
                rk.iterate(0, lo, prm, form, 3)
            if hi < rk.n0:
                rk.iterate(hi, rk.n0, prm, form, 3)
            if lo < hi:
                rk.iterate(lo, hi, prm, form, 2)
        for rk in ranks:
            rk.cur = 1 - rk.cur
        _exchange_iter(ranks, ranks[0].cur, periodic)
    x = np.concatenate([rk.x[0, 1:-1] for rk in ranks], axis=0).reshape(-1)
    z = np.concatenate([rk.zb[rk.cur][:, 1:-1] for rk in ranks], axis=1).reshape(-1)
    assert np.allclose(x, x_ref, rtol=1e-13, atol=1e-15) and np.allclose(z, z_ref, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("mode", ["constant", ("reflect", "symmetric", "edge"), ("wrap", "reflect", "symmetric")], ids=lambda m: m if isinstance(m, str) else "-".join(m))
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("shape,chunk", [((26, 6, 8), 3), ((41, 16, 64), 0), ((12, 16, 64), 0), ((96, 16, 64), 2)],
                         ids=["ragged-chunk3", "fulltiles", "one-chunk-slabs", "many-chunks"])
@pytest.mark.parametrize("order", [1, 2], ids=["edges-first", "edges-interleaved"])
def test_slab_iteration_with_the_exchange_fused_into_the_kernel(world, mode, shape, chunk, order):
    """pxb_pds_iter_p2p: ONE launch per rank and iteration; the kernel body stores the new boundary planes into the neighbours'
    ghost planes (here: the other simulated ranks' arrays), bumps their counters, and checks the neighbours' counters before
    reading its own ghost planes.  No exchange between iterations; result = the single-domain iteration, bit for bit in
    what is exchanged, to rounding in the rest (the single-domain reference runs the generic two-sweep bodies)."""
    n_iter, lam = 6, 0.08
    tau = sigma = 0.28
    rho = 1.3
    assert E.lib().emu_set_edge_first(order) == 0  # block order of the launch: the edge work items first (the library's default) / one in four
    y = np.random.default_rng(3).random(shape)
    m0 = mode if isinstance(mode, str) else mode[0]
    periodic = m0 == "wrap" and world > 1
    x_ref, z_ref = _single_domain(shape, y, mode, n_iter, tau, sigma, rho, lam)
    parts = partition(shape[0], world)
    ranks = [_RankIter(shape, y, a, b, r, world, "forward", mode, periodic) for r, (a, b) in enumerate(parts)]
    flags = [np.zeros(4, dtype=np.uint32) for _ in ranks]
    for r, rk in enumerate(ranks):  # what the engine's set-up exchanges once over NCCL: ghost planes of x0 and of the shift; z0 = K x0
        if r > 0 or periodic:
            ranks[(r - 1) % world].x[0][-1] = rk.x[0][1]
            ranks[(r - 1) % world].sh[0][-1] = rk.sh[0][1]
        if r + 1 < world or periodic:
            ranks[(r + 1) % world].x[0][0] = rk.x[0][rk.n0]
            ranks[(r + 1) % world].sh[0][0] = rk.sh[0][rk.n0]
    for rk in ranks:
        dd = rk.desc(0, rk.n0)
        E.lib().emu_gradient(C.byref(dd), 0, rk.ptr(rk.x, 0, 0), rk.ptr(rk.zb[0], 0, 0))
    _exchange_iter(ranks, 0, periodic)
    fn = E.lib().emu_tv_iter_tma_p2p
    fn.argtypes = E.lib().emu_tv_iter_tma.argtypes + [C.POINTER(K.Peer)]
    addr = lambda t, comp, plane_index, rk: t.ctypes.data + 8 * ((comp * t.shape[1] + plane_index) * rk.plane)
    for it in range(n_iter):
        for r, rk in enumerate(ranks):
            s, t = rk.cur, 1 - rk.cur
            pr = K.Peer()
            if rk.has_lo:
                lo = ranks[(r - 1) % world]
                pr.dn_u = addr(lo.ub[t], 0, rk.H + lo.n0, lo)
                pr.dn_z = addr(lo.zb[t], 0, rk.H + lo.n0, lo)
                pr.dn_zvol = lo.zb[t].shape[1] * lo.plane
                pr.dn_flag = flags[(r - 1) % world].ctypes.data + 4
                pr.lo_wait = flags[r].ctypes.data
            if rk.has_hi:
                hi = ranks[(r + 1) % world]
                pr.up_z0 = addr(hi.zb[t], 0, rk.H - 1, hi)
                pr.up_flag = flags[(r + 1) % world].ctypes.data
                pr.hi_wait = flags[r].ctypes.data + 4
            pr.epoch = it
            d = rk.desc(0, rk.n0)
            P = E.pds_params(tau, sigma, rho, gspec=(K.PROX_POS, 0, 0), fkind=K.F_SQL2, alpha=0.5, hkind=K.DUAL_L21, lam=lam)
            P.f.shift = rk.sh.ctypes.data + 8 * rk.H * rk.plane
            P.f.shift_period = rk.sh.shape[1] * rk.plane
            rc = fn(K.ALGO_PD3O, C.byref(d), C.byref(P), rk.ptr(rk.ub[s], 0, 0), rk.ptr(rk.zb[s], 0, 0), rk.ptr(rk.ub[t], 0, 0), rk.ptr(rk.zb[t], 0, 0),
                    rk.ptr(rk.x, 0, 0), None, None, chunk, C.byref(pr))
            assert rc == 0, (rc, r, it)
        for rk in ranks:
            rk.cur = 1 - rk.cur
    x = np.concatenate([rk.x[0, 1:-1] for rk in ranks], axis=0)
    z = np.concatenate([rk.zb[rk.cur][:, 1:-1] for rk in ranks], axis=1)
    assert np.allclose(x, x_ref, rtol=1e-13, atol=1e-15) and np.allclose(z, z_ref, rtol=1e-13, atol=1e-15)
    # every edge thread block signalled once per iteration: counters = iterations x tiles per plane
    tiles = -(-shape[1] // 8) * -(-shape[2] // 64)
    for r, rk in enumerate(ranks):
        assert flags[r][0] == (n_iter * tiles if rk.has_lo else 0) and flags[r][1] == (n_iter * tiles if rk.has_hi else 0)
    E.lib().emu_set_edge_first(1)


def _worker_many(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n0, h = 4, 1
        mk = lambda v: torch.full((n0 + 2 * h, 2, 3), float(v), dtype=torch.float64)
        u, z0, z1 = mk(10 + rank), mk(20 + rank), mk(30 + rank)
        hx = HaloExchanger()
        for r in hx.exchange_many([(u, False, True), (z0, True, True), (z1, False, True)], h, n0):
            r.wait()
        ok = True
        if rank + 1 < world:  # upper ghosts hold the upper neighbour's first planes
            ok &= bool((u[-1] == 10 + rank + 1).all() and (z0[-1] == 20 + rank + 1).all() and (z1[-1] == 30 + rank + 1).all())
        if rank > 0:  # lower ghost: only z0 travels up
            ok &= bool((z0[0] == 20 + rank - 1).all() and (u[0] == 10 + rank).all() and (z1[0] == 30 + rank).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_exchange_many_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 3
    procs = [ctx.Process(target=_worker_many, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


# ---- the z-slab path behind Solver.fit(), over gloo, on the emulated device ---------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_slab_worker_on_the_emulated_device(world):
    """tests/slab_worker.py -- the script the GPU test launches under torchrun -- with gloo and the device emulated
    (tests/emu_device.py): PD3O(...).fit() / CondatVu(...).fit() on every rank, decomposed into z-slabs (constant / folding modes /
    ring wrap, overlapped sub-range launches, single-kernel and two-sweep forms, Stencil data term, ShardedArray I/O), against
    fixtures of the real reference and the NumPy oracle."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PXB_SLAB_WORKER_DEVICE="cpu", OMP_NUM_THREADS="1", PYXU_B200_DENSE3D_MARCH="1")  # (the dense-PSF case included)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "tests", "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = 13 if world <= 3 else 12
    assert f"{n}/{n} cases OK" in r.stdout and "FAIL" not in r.stdout
