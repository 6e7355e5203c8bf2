"""torchrun worker: the z-slab decomposed solves -- PD3O(f, g, h, K).fit() / CondatVu(...).fit() called on every rank of
torch.distributed -- must reproduce the REFERENCE's results (tests/golden/slabs.npz, solvers.npz: produced by the real
pyxu, see tests/golden/make_golden.py), the NumPy oracle on a larger fp32 problem, and the iteration count of the
reference's default stopping criterion.

Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/slab_worker.py

PXB_SLAB_WORKER_DEVICE=cpu runs the same script without GPUs (tests/test_slab_cpu.py): gloo instead of NCCL, and the device
emulated by tests/emu_device.py (CPU tensors, the kernel bodies of tests/emu behind the C ABI) -- the solver / slab engines' own
Python (buffer layout, launch order, exchanges, norms, gathers) is then what is under test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main(dev="cuda"):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if dev == "cuda":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    import types

    import cases
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    px = types.SimpleNamespace(operator=pxo, solver=pxs, stop=pxst)
    from pyxu_b200 import slab
    from oracle import pyxu_oracle as orc

    gs = np.load(os.path.join(ROOT, "tests", "golden", "slabs.npz"))
    g0 = np.load(os.path.join(ROOT, "tests", "golden", "solvers.npz"))
    results = []

    def report(name, slv, errs, tol, extra=True):
        assert slv._astate.get("error") is None, slv._astate.get("error")
        eng = slv._slab
        assert eng is not None and eng.world == world, "the solve did not take the z-slab path"
        good = all(e < tol for e in errs) and bool(extra)
        results.append(good)
        if rank == 0:
            form = type(eng).__name__ + ("/single-kernel" if getattr(eng, "fused", True) else "/two-sweep")
            form += ", exchange: " + ("fused into the kernel (peer memory)" if getattr(eng, "p2p", None) is not None else "NCCL send/recv")
            print(f"[slab] world={world} {name} ({form}, overlap={eng.overlap}): rel.err vs reference " + ", ".join(f"{e:.2e}" for e in errs)
                  + f" {'OK' if good else 'FAIL'}", flush=True)

    def check_fixture(name, slv, g, prefix, tol=1e-10):
        data, hist = slv.stats()
        steps_ok = all(abs(float(slv._mstate[k]) - float(g[f"{prefix}/{k}"])) < 1e-8 for k in ("tau", "sigma", "rho"))
        report(name, slv, [relerr(data["x"], g[f"{prefix}/x"]), relerr(data["z"], g[f"{prefix}/z"])], tol,
               steps_ok and len(hist) == int(g[f"{prefix}/n_hist"]))

    always = dict(distributed=True)  # world 1 included: the engines then run without neighbours
    shape = (32, 12, 16)
    y = gs["y"]
    # -- PD3O-TV against the reference (constant boundaries, rho != 1) ---------------------------------------------
    slv = cases.build_tv_denoise(px, y, shape, lam=0.08)
    slv.fit(x0=y.reshape(-1).copy(), stop_crit=pxst.MaxIter(25), rho=1.2, **always)
    check_fixture("pd3o_tv3d 32x12x16 fp64", slv, gs, "pd3o_tv3d")
    # -- the reference's default criterion RelError[x] & RelError[z]: same iteration count on every world size --------
    slv = cases.build_tv_denoise(px, y, shape, lam=0.3)
    slv.fit(x0=y.reshape(-1).copy(), **always)
    check_fixture("pd3o_tv3d default stop", slv, gs, "pd3o_tv3d/default_stop", tol=1e-9)
    # -- folding boundary modes: ring along z ('wrap'), folds on the end ranks, in-plane folds ---------------------
    for tag, mode in (("ring", ("wrap", "reflect", "edge")), ("fold", ("reflect", "symmetric", "wrap")), ("edge", ("edge", "constant", "symmetric"))):
        slv = cases.build_tv_denoise(px, y, shape, lam=0.08, mode=mode)
        slv.fit(x0=y.reshape(-1).copy(), stop_crit=pxst.MaxIter(20), tuning_strategy=3, **always)
        check_fixture(f"pd3o_tv3d modes={mode}", slv, gs, f"pd3o_tv3d/{tag}")
    # ... and through the two-sweep form (pxb_pds_primal + pxb_pds_dual per slab), which serves whatever pxb_pds_iter declines
    from pyxu_b200 import _cabi

    _cabi.lib().pxb_set_iter_modes(0)
    for tag, mode in (("ring", ("wrap", "reflect", "edge")), ("fold", ("reflect", "symmetric", "wrap"))):
        slv = cases.build_tv_denoise(px, y, shape, lam=0.08, mode=mode)
        slv.fit(x0=y.reshape(-1).copy(), stop_crit=pxst.MaxIter(20), tuning_strategy=3, **always)
        assert slv._slab.fused is False
        check_fixture(f"pd3o_tv3d modes={mode}", slv, gs, f"pd3o_tv3d/{tag}")
    _cabi.lib().pxb_set_iter_modes(-1)
    # -- CondatVu TV denoising (pointwise data term, no g) ---------------------------------------------------------
    slv = cases.build_tv_denoise(px, y, shape, lam=0.08, solver="CondatVu", positivity=False)
    slv.fit(x0=np.zeros(y.size), stop_crit=pxst.MaxIter(25), **always)
    check_fixture("cv_tv3d", slv, gs, "cv_tv3d")
    # -- CondatVu deblurring, separable PSF reaching 3 planes across a cut (configs[4] in miniature) ----------------
    yb = gs["cv_deblur3d/y"]
    for tag, taps, cen, kw in (("cv_deblur3d", (7, 1.2), (3, 2, 3), dict(rho=0.9)), ("cv_deblur3d/even", (6, 1.2), (2, 2, 3), {})):
        psf = [cases.gaussian_1d(*taps), cases.gaussian_1d(5, 1.0), cases.gaussian_1d(7, 1.5)]
        slv, Aop = cases.build_tv_deblur(px, yb, shape, psf, cen, lam=0.02, positivity=True)
        slv.fit(x0=np.zeros(yb.size), stop_crit=pxst.MaxIter(15), **kw, **always)
        check_fixture(tag, slv, gs, tag)
    # -- the same with a DENSE 5x5x5 PSF of full rank (reaches 2 planes across a cut): the dense marching kernel per slab ---------
    from pyxu_b200.operator.linop import stencil as st_mod

    if st_mod.DENSE3D_MARCH:
        gd = np.load(os.path.join(ROOT, "tests", "golden", "dense3d.npz"))
        yd = gd["cv_deblur3d_dense/y"]
        slv, Aop = cases.build_tv_deblur(px, yd, shape, gd["cv_deblur3d_dense/psf"], (2, 2, 2), lam=0.02, positivity=True)
        slv.fit(x0=np.zeros(yd.size), stop_crit=pxst.MaxIter(15), rho=0.9, **always)
        assert slv._slab.dense
        check_fixture("cv_deblur3d_dense", slv, gd, "cv_deblur3d_dense")
    # -- the small fixtures of the single-GPU suite, when their slabs are thick enough --------------------------------
    if world <= 3:
        y3 = g0["pd3o_tv3d/y"]
        slv = cases.build_tv_denoise(px, y3, (10, 12, 14), lam=0.08)
        slv.fit(x0=y3.reshape(-1).copy(), stop_crit=pxst.MaxIter(50), **always)
        check_fixture("pd3o_tv3d 10x12x14 (solvers.npz)", slv, g0, "pd3o_tv3d")
    # -- ShardedArray in -> ShardedArray out: every rank holds only its planes; device tensors; fp32 vs the NumPy oracle -----
    shape32, lam, n_iter = (64, 48, 64), 0.08, 25
    N = int(np.prod(shape32))
    y32 = np.random.default_rng(7).random(shape32).astype(np.float32)
    a, b = slab.partition(shape32[0], world)[rank]
    y_loc = torch.from_numpy(y32[a:b].copy()).to(dev)
    sh = lambda t: slab.ShardedArray(t, shape32, rank=rank, world=world)
    f = 0.5 * px.operator.SquaredL2Norm(dim=N).argshift(sh(-y_loc))
    Kop = px.operator.Gradient(arg_shape=shape32, dtype=np.float32)
    h = lam * px.operator.L21Norm(arg_shape=(3, *shape32), l2_axis=(0,))
    slv = pxs.PD3O(f=f, g=px.operator.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False, final_writeback=False)
    crit = pxst.MaxIter(n_iter) | pxst.RelError(eps=1e-30, var="x") | pxst.RelError(eps=1e-30, var="z")
    slv.fit(x0=sh(y_loc.clone()), stop_crit=crit)
    xs = slv.solution()
    assert isinstance(xs, slab.ShardedArray) and (xs.start, xs.stop) == (a, b) and torch.is_tensor(xs.local)
    prob = orc.tv_problem(y32.astype(np.float64).reshape(-1), shape32, lam)
    tau, sigma, rho = (slv._mstate[k] for k in ("tau", "sigma", "rho"))
    st = orc.pd3o_init(y32.astype(np.float64).reshape(-1), prob["K"])
    x_prev = None
    for _ in range(n_iter):
        x_prev = st["x"].copy()
        orc.pd3o_step(st, tau, sigma, rho, prob["grad_f"], prob["prox_g"], prob["prox_h"], prob["K"], prob["KT"])
    x_ref = st["x"].reshape(shape32)[a:b]
    _, hist = slv.stats()
    rel_x = float(np.linalg.norm(st["x"] - x_prev) / np.linalg.norm(x_prev))
    e_rx = abs(float(hist[-1]["RelError[x]"]) - rel_x) / rel_x
    report("pd3o_tv3d 64x48x64 fp32, ShardedArray device tensors, vs oracle", slv, [relerr(xs.local.cpu().numpy(), x_ref), e_rx * 1e-2], 1e-4)

    ok = all(results)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[slab] world={world}: {sum(results)}/{len(results)} cases OK", flush=True)
    del slv
    slab.release_pool()  # drop the mappings of the neighbours' buffers before anybody's process ends
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    if os.environ.get("PXB_SLAB_WORKER_DEVICE", "cuda") == "cpu":
        from emu_device import emulated_device

        with emulated_device(cuda_runtime=True):
            main("cpu")
    else:
        main()
