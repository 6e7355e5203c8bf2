#!/usr/bin/env python
"""
Secondary workloads of BASELINE.json (configs[1], configs[2]) through the public solver API, device-resident.
(bench.py stays the headline PD3O-TV 1024^3 line; these numbers go to DESIGN.md / profiles.)

    python tools/bench_configs.py --workload deblur2d [--size 8192] [--steps 20]
    python tools/bench_configs.py --workload fista   [--batch 256] [--size 1024]     (torchrun: the batch is dealt out to the ranks)
    torchrun ... tools/bench_configs.py --workload deblur3d [--shape 2048,2048,1024]  (z-slabs over the ranks)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import pyxu_b200.operator as pxo
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst
from pyxu_b200 import _cabi
from pyxu_b200.abc import Mode
from pyxu_b200.slab import split_batch


def gauss(n, s):
    t = np.arange(n) - (n - 1) / 2
    k = np.exp(-0.5 * (t / s) ** 2)
    return (k / k.sum()).astype(np.float32)


def timed(step, K, W, world):
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _cabi.launch_count()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, _cabi.launch_count() - l0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["deblur2d", "fista", "deblur3d", "tv2d"])
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dense", action="store_true", help="dense instead of separable blur kernel")
    ap.add_argument("--shape", default="2048,2048,1024")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif args.workload == "deblur3d":  # the slab class talks to torch.distributed even with one rank
        import socket

        s_ = socket.socket()
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
        s_.close()
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", local))
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    K, W = args.steps, max(3, args.warmup)
    if args.workload == "tv2d":
        # configs[0]: the reference's own CPU-runnable case -- 2-D TV denoising 512x512 float64, PD3O, 200 iterations,
        # HOST arrays in and out through Solver.fit() (wall clock, transfers and Python included)
        import time

        n = args.size or 512
        shape, N = (n, n), n * n
        rng = np.random.default_rng(0)
        y = rng.random(N)
        def solve():
            f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)
            Kop = pxo.Gradient(arg_shape=shape)
            h = 0.1 * pxo.L21Norm(arg_shape=(2, *shape), l2_axis=(0,))
            slv = pxs.PD3O(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, show_progress=False, final_writeback=False)
            slv.fit(x0=y.copy(), stop_crit=pxst.MaxIter(200))
            return slv.solution()
        solve()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = solve()
        dt = time.perf_counter() - t0
        assert isinstance(x, np.ndarray) and x.dtype == np.float64
        print(json.dumps({"workload": f"2-D TV denoising {n}x{n} float64, PD3O, 200 iterations, host arrays through Solver.fit()", "n_gpus": 1,
                          "seconds": dt, "iterations_per_s": 200 / dt, "gvoxel_iter_per_s": N * 200 / dt / 1e9}))
        return
    if args.workload == "deblur2d":
        n = args.size or 8192
        shape, N = (n, n), n * n
        g1 = gauss(9, 1.7)
        kern = np.outer(g1, g1) if args.dense else [g1, g1]
        A = pxo.Stencil(arg_shape=shape, kernel=kern, center=(4, 4), mode="constant")
        y = torch.rand(N, device="cuda", generator=gen)
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)) * A
        Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
        h = 0.05 * pxo.L21Norm(arg_shape=(2, *shape), l2_axis=(0,))
        slv = pxs.CondatVu(f=f, g=None, h=h, K=Kop, beta=float(A.lipschitz) ** 2, show_progress=False)
        slv.fit(x0=y, mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
        assert slv._plan.kind == "fused"
        ms, launches = timed(slv.m_step, K, W, world)
        assert slv._plan.iter_ok is True and A._tiled_ok is True
        nvox, name = N, f"2-D TV deblurring {n}x{n} fp32, CondatVu, {'dense' if args.dense else 'separable'} 9x9 Gaussian Stencil blur + L21 o Gradient"
        bpv = 12 + 8 + 28  # A x - y (read x, y; write r) + A^T r + single-kernel CV iteration (read x, grad f, z0, z1; write x, z0, z1)
    elif args.workload == "deblur3d":
        from pyxu_b200.slab import ShardedArray, partition

        shape = tuple(int(v) for v in args.shape.split(","))
        N = int(np.prod(shape))
        a, b = partition(shape[0], world)[rank]
        sh = lambda t: ShardedArray(t, shape, rank=rank, world=world)
        y_local = torch.rand((b - a, *shape[1:]), device="cuda", generator=gen)
        g7 = gauss(7, 1.2)
        A = pxo.Stencil(arg_shape=shape, kernel=[g7, g7, g7], center=(3, 3, 3), mode="constant")
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(sh(-y_local))) * A
        Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
        h = 0.05 * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
        slv = pxs.CondatVu(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=Kop, beta=float(A.lipschitz) ** 2, show_progress=False)
        slv.fit(x0=sh(y_local), mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), distributed=True)
        del y_local
        ms, launches = timed(slv.m_step, K, W, world)
        nvox = N
        name = f"3-D TV deblurring {'x'.join(map(str, shape))} fp32, CondatVu, separable 7x7x7 Stencil PSF + positivity + L21 o Gradient, {world} z-slab(s)"
        # 2 alpha (A x + c) (read x, c; write r) | A^T r | single-kernel CV iteration with grad f array     (two-pass stencils: 8+12+8+8+36)
        bpv = (12 + 8 + 36) if slv._slab.single_pass else (8 + 12 + 8 + 8 + 36)
    else:
        n = args.size or 1024
        shape, N = (n, n), n * n
        lo, hi = split_batch(args.batch, world)[rank]
        B = hi - lo
        g1 = gauss(5, 1.0)
        psf = np.outer(g1, g1)
        if args.dense:  # a PSF that is not an outer product: the register-blocked dense instance
            psf = psf + 0.02 * np.eye(5, dtype=np.float32)
        A = pxo.Stencil(arg_shape=shape, kernel=psf, center=(2, 2), mode="constant")
        y = torch.rand(B, N, device="cuda", generator=gen)
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)) * A
        g = 0.02 * pxo.L1Norm(dim=N)
        slv = pxs.PGD(f=f, g=g, show_progress=False)
        slv.fit(x0=y, mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), tau=1.0 / float(A.lipschitz) ** 2)
        ms, launches = timed(slv.m_step, K, W, world)
        assert slv._fused is not None, "the two-pass tiled FISTA form did not apply"
        nvox, name = args.batch * N, f"batch of {args.batch} {n}x{n} images, PGD (FISTA) L1 deconvolution, 5x5 Stencil ({'full-rank PSF: dense instance' if args.dense else 'Gaussian PSF given as a 2-D array: rank 1, separable passes'}), batch split over {world} GPU(s)"
        bpv = 16 + 16  # r = A y - b: read x, x_prev, b, write r;  x_new = prox(y - tau A^T r): read r, x, x_prev, write x_new
    if rank == 0:
        per = ms / K
        print(json.dumps({"workload": name, "n_gpus": world, "ms_per_iter": per, "gvoxel_iter_per_s": nvox / per / 1e6,
                          "algorithmic_bytes_per_voxel": bpv, "achieved_GBps_per_gpu": bpv * nvox / world / per / 1e6,
                          "launches_per_iter": launches / K}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
