"""torchrun worker: z-slab PD3O-TV on WORLD_SIZE GPUs must reproduce the single-GPU PD3O solver.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/slab_worker.py

PXB_SLAB_WORKER_DEVICE=cpu runs the same script without GPUs (tests/test_slab_cpu.py): gloo instead of NCCL, and the device
emulated by tests/emu_device.py (CPU tensors, the kernel bodies of tests/emu behind the C ABI) -- the slab classes' own Python
(buffer layout, launch order, exchanges, norms) is then what is under test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(dev="cuda"):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if dev == "cuda":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst
    from pyxu_b200.slab import SlabPD3OTV

    ok = True
    for shape, mode, overlap, dtype, tol in [((37, 24, 32), "constant", True, torch.float64, 1e-13),
                                             ((37, 24, 32), "constant", False, torch.float64, 1e-13),
                                             ((40, 20, 28), ("reflect", "wrap", "edge"), True, torch.float64, 1e-13),
                                             ((32, 16, 24), "wrap", True, torch.float64, 1e-13),
                                             ((64, 48, 64), "constant", True, torch.float32, 1e-5)]:
        n_iter, lam = 25, 0.08
        gen = torch.Generator(device=dev).manual_seed(7)
        y = torch.rand(shape, device=dev, dtype=dtype, generator=gen)
        slab = SlabPD3OTV(shape, y_full=y, lam=lam, positivity=True, dtype=dtype, mode=mode, overlap=overlap, rho=1.2)
        v = None
        for i in range(n_iter):  # norms only now and then: x is then rebuilt from the previous iterate when needed
            v = slab.step(want_norms=(i % 3 == 2 or i == n_iter - 1))
        x_slab = slab.gather_x().reshape(-1)
        # single-GPU reference through the public solver, same step sizes
        N = int(np.prod(shape))
        f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y.reshape(-1))
        Kop = pxo.Gradient(arg_shape=shape, mode=mode, dtype=np.float64 if dtype == torch.float64 else np.float32)
        h = lam * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
        slv = pxs.PD3O(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False, final_writeback=False)
        sc = pxst.MaxIter(n_iter) | pxst.RelError(eps=1e-30, var="x") | pxst.RelError(eps=1e-30, var="z")
        slv.fit(x0=y.reshape(-1), stop_crit=sc, tau=slab.tau, sigma=slab.sigma, rho=1.2)
        assert slv._astate.get("error") is None, slv._astate.get("error")
        x_ref = slv._mstate["x"]
        err = float((x_slab - x_ref).norm() / x_ref.norm())
        _, hist = slv.stats()
        rx, rz = slab.rel_errors(v)
        e_rx = abs(rx - float(hist[-1]["RelError[x]"])) / max(rx, 1e-300)
        e_rz = abs(rz - float(hist[-1]["RelError[z]"])) / max(rz, 1e-300)
        good = err < tol and e_rx < 1e-6 and e_rz < 1e-6
        ok &= good
        if rank == 0:
            print(f"[slab] world={world} shape={shape} mode={mode} overlap={slab.overlap} {dtype}: rel.err={err:.2e} "
                  f"relerr-norms dev=({e_rx:.1e},{e_rz:.1e}) {'OK' if good else 'FAIL'}", flush=True)
    # ---- CondatVu TV deblurring with a separable 3-D PSF (configs[4] in miniature) -------------------------------
    from pyxu_b200.slab import SlabCondatVuDeblur

    def gauss(n, s):
        t_ = np.arange(n) - (n - 1) / 2
        k = np.exp(-0.5 * (t_ / s) ** 2)
        return k / k.sum()

    for shape, psf, cen, dtype, tol, ovl in [((41, 24, 32), [gauss(7, 1.2), gauss(5, 1.0), gauss(7, 1.5)], (3, 2, 3), torch.float64, 1e-12, True),
                                             ((41, 24, 32), [gauss(6, 1.2), gauss(5, 1.0), gauss(7, 1.5)], (2, 2, 3), torch.float64, 1e-12, True),
                                             ((72, 40, 64), [gauss(7, 1.2), gauss(7, 1.2), gauss(7, 1.2)], (3, 3, 3), torch.float32, 2e-5, True),
                                             ((48, 40, 64), [gauss(7, 1.2), gauss(7, 1.2), gauss(7, 1.2)], (3, 3, 3), torch.float32, 2e-5, False)]:
        n_iter, lam = 15, 0.05
        npdt = np.float64 if dtype == torch.float64 else np.float32
        gen = torch.Generator(device=dev).manual_seed(11)
        y = torch.rand(shape, device=dev, dtype=dtype, generator=gen)
        slab = SlabCondatVuDeblur(shape, psf, cen, y_full=y, lam=lam, positivity=True, dtype=dtype, rho=0.9, overlap=ovl)
        v = None
        for i in range(n_iter):
            v = slab.step(want_norms=(i == n_iter - 1))
        x_slab = slab.gather_x().reshape(-1)
        N = int(np.prod(shape))
        Aop = pxo.Stencil(arg_shape=shape, kernel=[np.asarray(k, dtype=npdt) for k in psf], center=cen, mode="constant")
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y.reshape(-1))) * Aop
        Kop = pxo.Gradient(arg_shape=shape, dtype=npdt)
        h = lam * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
        slv = pxs.CondatVu(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, beta=float(Aop.lipschitz) ** 2, show_progress=False, final_writeback=False)
        sc = pxst.MaxIter(n_iter) | pxst.RelError(eps=1e-30, var="x") | pxst.RelError(eps=1e-30, var="z")
        slv.fit(x0=y.reshape(-1), stop_crit=sc, tau=slab.tau, sigma=slab.sigma, rho=0.9)
        assert slv._astate.get("error") is None, slv._astate.get("error")
        x_ref = slv._mstate["x"]
        err = float((x_slab - x_ref).norm() / x_ref.norm())
        _, hist = slv.stats()
        rx = float(np.sqrt(v[0]) / np.sqrt(v[1]))
        e_rx = abs(rx - float(hist[-1]["RelError[x]"])) / max(rx, 1e-300)
        good = err < tol and e_rx < 1e-5
        ok &= good
        if rank == 0:
            print(f"[slab-deblur] world={world} shape={shape} {dtype} overlap={slab.overlap} single_pass={slab.single_pass}: rel.err={err:.2e} "
                  f"relerr-norm dev={e_rx:.1e} {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    if os.environ.get("PXB_SLAB_WORKER_DEVICE", "cuda") == "cpu":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from emu_device import emulated_device

        with emulated_device(cuda_runtime=True):
            main("cpu")
    else:
        main()
