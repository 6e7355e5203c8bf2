#!/usr/bin/env python
"""
bench.py -- headline benchmark of the hot path (BASELINE.json: "PD3O-TV Gvoxel-iter/s").

Workload (configs[3] of BASELINE.json, the configuration the metric is quoted on): 3-D TV denoising of
a 1024^3 fp32 synthetic phantom with PD3O  (f = 1/2||x - y||^2, g = positivity, h = lam*L21 o Gradient).
One "step" = one PD3O iteration over the whole volume.  With N GPUs the volume is cut into N z-slabs
(strong scaling: total work fixed) with one-plane halo exchange per half-iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over ranks);
`e2e` is the same metric through the public API with HOST buffers (Solver.fit(x0=<pinned host array>)
-> solution() back on the host, the per-iteration stopping metric read back every step);
`roofline` is algorithmic bytes / measured kernel time of the dominant kernel vs the measured HBM peak;
`cpu_baseline` times the oracle's C/OpenMP port of the same iteration on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "pd3o_tv_gvoxel_iter_per_s", "Gvoxel-iter/s"
LAM = 0.08
# algorithmic HBM bytes per voxel per launch (fp32; DESIGN.md "Kernels"):
#   single-kernel iteration: read u, y, z0, z1, z2 + write u, z0, z1, z2 = 9 floats (x is materialised on demand)
#   two-sweep form:  primal: read u, z0, z1, z2, y + write x, w, u = 8 floats;   dual: read w, z0..2 + write z0..2 = 7 floats
BYTES_PER_VOXEL = {"pxb_pds_iter": 9 * 4, "pxb_pds_primal": 8 * 4, "pxb_pds_dual": 7 * 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=1024, help="cube edge of the volume (default: the named 1024^3 workload)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-size", type=int, default=320, help="cube edge of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--two-sweep", action="store_true", help="force the two-kernel form of the iteration (A/B comparison)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle C/OpenMP port (oracle/tv_oracle.c) on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------
def cpu_port():
    from pyxu_b200 import _build

    path = _build.build_oracle()
    h = ctypes.CDLL(path)
    vp, i, i64, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    h.tv_pd3o_f32.argtypes = [vp, vp, vp, vp, i, i64, i64, i64, d, d, d, d, d, i, i]
    h.tv_num_threads.restype = i
    h.tv_set_threads.argtypes = [i]
    return h


def phantom_np(n, seed=0):
    import numpy as np

    rng = np.random.default_rng(seed)
    c = max(1, n // 16)
    coarse = rng.random((16, 16, 16)).astype(np.float32)
    x = np.repeat(np.repeat(np.repeat(coarse, c, 0), c, 1), c, 2)[:n, :n, :n]
    x = np.ascontiguousarray(x)
    return x + 0.1 * rng.standard_normal(x.shape, dtype=np.float32)


def step_sizes(ndim=3):
    # PD3O defaults for beta = 1, ||K|| <= sqrt(4*ndim) (same rule as PD3O._set_step_sizes)
    import math

    L = math.sqrt(4.0 * ndim)
    t = min(0.5 * (math.log(0.99) - 2 * math.log(L)), 0.0)
    return math.exp(t), math.exp(t), 1.0


_BEST_THREADS = None


def run_cpu(n, steps, warmup):
    """Returns (Gvoxel-iter/s, threads, seconds) of the C/OpenMP port on an n^3 fp32 volume.

    The thread count is the one that runs fastest on this host (all hardware threads unless the container's CPU
    quota makes fewer threads faster), found once on a small probe volume."""
    global _BEST_THREADS
    import numpy as np

    h = cpu_port()
    tau, sigma, rho = step_sizes()
    p = lambda a: a.ctypes.data
    if _BEST_THREADS is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        hw = os.cpu_count() or 1
        m = 96
        yy = phantom_np(m).reshape(-1)
        best = (float("inf"), 1)
        t = hw
        while t >= 1:
            h.tv_set_threads(t)
            xx, uu, zz = yy.copy(), yy.copy(), np.zeros(3 * yy.size, dtype=np.float32)
            h.tv_pd3o_f32(p(yy), p(xx), p(uu), p(zz), 3, m, m, m, 0.5, tau, sigma, rho, LAM, 1, 1)
            t0 = time.perf_counter()
            h.tv_pd3o_f32(p(yy), p(xx), p(uu), p(zz), 3, m, m, m, 0.5, tau, sigma, rho, LAM, 1, 2)
            best = min(best, (time.perf_counter() - t0, t))
            t //= 2
        _BEST_THREADS = best[1]
    h.tv_set_threads(_BEST_THREADS)
    y = phantom_np(n).reshape(-1)
    x, u = y.copy(), y.copy()
    z = np.zeros(3 * y.size, dtype=np.float32)
    if warmup:
        h.tv_pd3o_f32(p(y), p(x), p(u), p(z), 3, n, n, n, 0.5, tau, sigma, rho, LAM, 1, warmup)
    t0 = time.perf_counter()
    h.tv_pd3o_f32(p(y), p(x), p(u), p(z), 3, n, n, n, 0.5, tau, sigma, rho, LAM, 1, steps)
    dt = time.perf_counter() - t0
    return (n**3) * steps / dt / 1e9, int(h.tv_num_threads()), dt


def reference_arm(args, rank):
    if rank != 0:
        return
    n = args.cpu_size
    val, threads, dt = run_cpu(n, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"3-D TV denoising {args.size}^3 fp32, PD3O (SquaredL2Norm + L21Norm o Gradient + PositiveOrthant)",
                   "sample": f"{n}^3 sub-volume, {args.steps} iterations"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"oracle/tv_oracle.c (C/OpenMP pass-by-pass port of the reference's PD3O iteration) on a {n}^3 fp32 phantom, {args.steps} iterations"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in ln.split(",")]))

    def mark(self):
        """host time stamp; summary(t0, t1) keeps the samples taken between two marks"""
        return time.time()

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        rows = [r for ts, r in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.11)]
        in_window = len(rows)
        if not rows and self.rows:  # region shorter than the sampling period: the sample nearest to it
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - mid))[1]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": in_window}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, nvox):
    """dram bytes/launch of `kernel` from the committed ncu --set full capture, rescaled per voxel (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)[kernel]
        return float(t["dram_bytes_per_voxel"]) * nvox
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst
    from pyxu_b200 import _cabi
    from pyxu_b200.abc import Mode

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.size
    shape = (n, n, n)
    nvox = n**3
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # synthetic phantom, generated on the device (not timed)
    gen = torch.Generator(device=dev).manual_seed(1234)
    c = max(1, n // 16)
    coarse = torch.rand((16, 16, 16), device=dev, dtype=torch.float32, generator=gen)
    y = coarse.repeat_interleave(c, 0).repeat_interleave(c, 1).repeat_interleave(c, 2)[:n, :n, :n].contiguous()
    y += 0.1 * torch.randn(y.shape, device=dev, dtype=torch.float32, generator=gen)
    del coarse
    y_host = shift_host = None
    if world == 1 and not args.no_e2e:  # host copies for the end-to-end run (made outside every timed region)
        y_host = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
        y_host.copy_(y.reshape(-1))
        shift_host = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
        torch.neg(y_host, out=shift_host)


    if world == 1:
        N = nvox
        f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y.reshape(-1))
        Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
        h = LAM * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
        g = pxo.PositiveOrthant(dim=N)
        slv = pxs.PD3O(f=f, g=g, h=h, K=Kop, show_progress=False)
        slv.fit(x0=y.reshape(-1), mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
        assert slv._plan.kind == "fused"
        if args.two_sweep:
            slv._plan.iter_ok = False
        step = slv.m_step
        pending = []

        def probe(tag):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            pending.append((tag, ev))

        slv._probe = probe
        local_vox = nvox
    else:
        from pyxu_b200.slab import SlabPD3OTV

        slv = SlabPD3OTV(shape, y_full=y, lam=LAM, positivity=True, dtype=torch.float32)
        del y
        step = slv.step
        pending = slv.event_log
        local_vox = slv.local_voxels

    def run_steps(k):
        for _ in range(k):
            step()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:  # nvidia-smi is started before the warm-up so that it is sampling when the timed region begins
        run_steps(W)
        pending.clear()
        barrier()
        l0 = _cabi.launch_count()
        t_begin = clk.mark()
        e0.record()
        run_steps(K)
        e1.record()
        barrier()
        t_end = clk.mark()
        l1 = _cabi.launch_count()
    ms = e0.elapsed_time(e1)
    launches = int(l1 - l0)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = nvox * K / (ms * 1e-3) / 1e9

    # per-kernel durations from the events recorded inside the timed region
    tags = {}
    for (t0, a), (t1, b) in zip(pending[:-1], pending[1:]):
        if t0 == "primal_begin" and t1 == "primal_end":
            tags.setdefault("pxb_pds_primal", []).append(a.elapsed_time(b))
        if t0 == "primal_end" and t1 == "dual_end":
            tags.setdefault("pxb_pds_dual", []).append(a.elapsed_time(b))
        if t0 == "iter_begin" and t1 == "iter_end":
            tags.setdefault("pxb_pds_iter", []).append(a.elapsed_time(b))
    peak, peak_src = measured_peak()
    roof = {}
    for kname, ts in tags.items():
        avg = sum(ts) / len(ts)
        bytes_alg = BYTES_PER_VOXEL[kname] * local_vox
        ach = bytes_alg / (avg * 1e-3) / 1e9
        roof[kname] = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "traffic": ncu_traffic(kname, local_vox), "avg_ms": avg, "algorithmic_bytes": bytes_alg,
                       "peak_source": peak_src, "launches_timed": len(ts)}
    dominant = max(roof.values(), key=lambda r: r["avg_ms"] * r["launches_timed"]) if roof else None

    # ---- end to end through the public API with HOST buffers (single GPU) -----------------------------
    e2e = None
    if world == 1 and not args.no_e2e:
        slv._probe = None
        del slv, f, h, g, step
        torch.cuda.empty_cache()
        shift_np, x0_np = shift_host.numpy(), y_host.numpy()  # pinned host memory seen as NumPy arrays
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f2 = 0.5 * pxo.SquaredL2Norm(dim=nvox).argshift(shift_np)
        slv2 = pxs.PD3O(f=f2, g=pxo.PositiveOrthant(dim=nvox), h=LAM * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,)),
                        K=Kop, show_progress=False, final_writeback=False)
        # K iterations; the RelError metric is read back from the device every iteration (eps tiny: never triggers)
        t1 = time.perf_counter()
        slv2.fit(x0=x0_np, stop_crit=pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x"))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        x_host = slv2.solution()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_parts = {"build_s": t1 - t0, "fit_s": t2 - t1, "solution_s": t0 + dt - t2}
        assert slv2._astate.get("error") is None, slv2._astate.get("error")
        assert slv2._plan.kind == "fused" and "_fused_norms" in slv2._mstate
        assert isinstance(x_host, np.ndarray) and x_host.shape == (nvox,) and np.isfinite(x_host[:: max(1, nvox // 1000)]).all()
        e2e = {"value": nvox * K / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(2 * 4 * nvox / K),
               "d2h_bytes_per_step": int(4 * nvox / K + 16), "seconds": dt, "parts": e2e_parts,
               "what": "PD3O(...).fit(x0=<host array>, stop_crit=MaxIter(K)|RelError) + solution(): H2D of x0 and of the data y, "
                       "K fused iterations with the RelError scalars read back every step, D2H of x"}
        del slv2
    elif world > 1 and not args.no_e2e:
        # multi-GPU end to end: every rank uploads its z-slab of the data from pinned host memory, K iterations with the
        # RelError sums all-reduced and read back every step, the slab of x copied back to a host array
        from pyxu_b200 import _array as A_
        from pyxu_b200.slab import SlabPD3OTV as _Slab, partition as _part

        a_, b_ = _part(n, world)[rank]
        y_host = torch.empty((b_ - a_, n, n), dtype=torch.float32, pin_memory=True)
        y_host.copy_(slv.shift_h[0, slv.HALO : slv.HALO + slv.n0].neg())  # (the phantom's slab, staged outside the timed region)
        del slv, step
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        y_dev = y_host.to(dev, non_blocking=True)
        slv2 = _Slab(shape, y_local=y_dev, lam=LAM, positivity=True, dtype=torch.float32)
        for _ in range(K):
            v = slv2.step(want_norms=True)
        slv2.materialize_x()
        x_host = A_.restore(slv2.x[0, slv2.HALO : slv2.HALO + slv2.n0].contiguous().reshape(-1), A_.HOST)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert isinstance(x_host, np.ndarray) and np.isfinite(x_host[:: max(1, x_host.size // 1000)]).all() and np.isfinite(v).all()
        e2e = {"value": nvox * K / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(4 * nvox / K), "d2h_bytes_per_step": int(4 * nvox / K + 32 * world),
               "seconds": dt,
               "what": "per rank: H2D of its z-slab of the data from pinned host memory, SlabPD3OTV set-up (halo exchange, z0 = K x0), K single-kernel "
                       "iterations with the RelError sums all-reduced and read back every step, D2H of its slab of x; wall clock, max over ranks"}
        del slv2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ncpu = args.cpu_size
        v1, thr, dt1 = run_cpu(ncpu, 2, 1)
        iters = int(max(3, min(400, 12.0 / max(dt1 / 2, 1e-3))))  # ~12 s of CPU work
        v, thr, dtc = run_cpu(ncpu, iters, 0)
        cpu = {"value": v, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"oracle/tv_oracle.c (C/OpenMP pass-by-pass port of the reference's PD3O iteration) on a {ncpu}^3 fp32 phantom, {iters} iterations, {dtc:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3-D TV denoising {n}^3 fp32, PD3O (SquaredL2Norm + L21Norm o Gradient + PositiveOrthant)",
                       "decomposition": "single GPU" if world == 1 else f"{world} z-slabs, boundary planes of the new iterate (5 planes per interface) exchanged by NCCL send/recv while the interior is computed",
                       "l2_policy": f"inputs larger than L2: {4 * nvox / world / 2**20:.0f} MiB per field per GPU vs 126 MB L2",
                       "iterations_per_step": 1},
            "clocks": clk.summary(t_begin, t_end), "e2e": e2e, "gpu_launches": launches,
            "roofline": dominant, "roofline_all": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
