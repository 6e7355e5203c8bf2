#!/usr/bin/env python
"""
bench.py -- headline benchmark of the hot path (BASELINE.json: "PD3O-TV Gvoxel-iter/s").

Workload (configs[3] of BASELINE.json, the configuration the metric is quoted on): 3-D TV denoising of
a 1024^3 fp32 synthetic phantom with PD3O  (f = 1/2||x - y||^2, g = positivity, h = lam*L21 o Gradient).
One "step" = one PD3O iteration over the whole volume.  The SAME `pxs.PD3O(f, g, h, K)` object is built at every
N; with N > 1 ranks `fit()` decomposes the volume into N z-slabs (strong scaling: total work fixed) with boundary
planes exchanged over NCCL while the interior is computed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Prints ONE JSON line (rank 0):
  value         device-resident throughput of K iterations (CUDA events, max over ranks), no stopping criterion
  roofline      algorithmic bytes / measured kernel time of the dominant kernel vs the measured HBM peak;
  roofline_all  ... plus the same iteration driven by the reference's default criterion RelError[x] & RelError[z]
                (norms fused into the kernel, x written every iteration, 32-byte readback per step: 44 B/voxel)
  e2e           the same metric through the public API with HOST buffers: PD3O(...).fit(x0=<pinned host array>) +
                solution() back on the host, RelError read back every step; `parts` = where the time goes
  parity        a fixture of the real reference replayed on the N ranks before anything is timed
  configs       the other BASELINE.json configurations, each with ms/iteration, algorithmic bytes and roofline fraction
  cpu_baseline  the reference itself (oracle/_ref, NumPy/Numba) or the oracle's C/OpenMP port on the host cores
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
#   single-kernel iteration: read u, y, z0, z1, z2 + write u, z0, z1, z2 = 9 floats (x is materialised on demand);
#   with the RelError criterion: + write x, + read the previous x = 11 floats
#   two-sweep form:  primal: read u, z0, z1, z2, y + write x, w, u = 8 floats;   dual: read w, z0..2 + write z0..2 = 7 floats
BYTES_PER_VOXEL = {"pxb_pds_iter": 9 * 4, "pxb_pds_iter+criterion": 11 * 4, "pxb_pds_primal": 8 * 4, "pxb_pds_dual": 7 * 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=1024, help="cube edge of the volume (default: the named 1024^3 workload)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-size", type=int, default=320, help="cube edge of the bounded CPU sample (C/OpenMP port)")
    ap.add_argument("--ref-size", type=int, default=160, help="cube edge of the bounded sample the real reference is timed on")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE.json configurations")
    ap.add_argument("--two-sweep", action="store_true", help="force the two-kernel form of the iteration (A/B comparison)")
    return ap.parse_args()


def workload_name(n):
    return f"3-D TV denoising {n}^3 fp32, PD3O (SquaredL2Norm + L21Norm o Gradient + PositiveOrthant)"


# ---------------------------------------------------------------------------------------------
# CPU arms: the real reference staged under oracle/_ref (NumPy/Numba), and the oracle's C/OpenMP port
# ---------------------------------------------------------------------------------------------
def cpu_port():
    from oracle import build as obuild

    h = ctypes.CDLL(obuild.build_port())
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


_BEST_THREADS = None


def run_port(n, steps, warmup):
    """(Gvoxel-iter/s, threads, seconds) of the C/OpenMP port on an n^3 fp32 volume, PD3O's default step sizes.

    The thread count is the one that runs fastest on this host (all hardware threads unless the container's CPU
    quota makes fewer threads faster), found once on a small probe volume."""
    global _BEST_THREADS
    import math

    import numpy as np

    h = cpu_port()
    L = math.sqrt(12.0)  # ||Gradient|| bound in 3-D; same rule as PD3O._set_step_sizes for beta = 1
    tau = sigma = math.exp(min(0.5 * (math.log(0.99) - 2 * math.log(L)), 0.0))
    rho = 1.0
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


def run_reference(n, steps, warmup, budget_s=None):
    """(Gvoxel-iter/s, threads, seconds, iterations) of the REAL reference (pyxu.opt.solver.PD3O, NumPy + Numba stencils,
    single precision like our arm) on an n^3 volume: pds.py:747-761 driven through m_step().  Raises when the staged
    reference cannot be imported here."""
    import numpy as np

    from oracle import build as obuild

    ns = obuild.load_ref()
    import pyxu.runtime as pxrt

    N, shape = n**3, (n, n, n)
    y = phantom_np(n).reshape(-1)
    with pxrt.Precision(pxrt.Width.SINGLE):
        pxo = ns.operator
        f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y)
        K = pxo.Gradient(arg_shape=shape, dtype=np.float32)
        h = LAM * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
        slv = ns.solver.PD3O(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=K, show_progress=False)
        slv.fit(x0=y.copy(), mode=ns.abc.Mode.MANUAL, stop_crit=ns.stop.ManualStop())
        for _ in range(max(2, warmup)):  # Numba compiles the stencils on the first calls
            slv.m_step()
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            slv.m_step()
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        assert slv._mstate["x"].dtype == np.float32 and np.isfinite(slv._mstate["x"][:: max(1, N // 1000)]).all()
    try:
        import numba

        threads = int(numba.get_num_threads())
    except Exception:
        threads = os.cpu_count() or 1
    return N * done / dt / 1e9, threads, dt, done


def reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 2)
    n = args.ref_size
    try:
        val, threads, dt, done = run_reference(n, K, W, budget_s=150.0)
        kind, why = "reference", None
        sample = (f"the real pyxu.opt.solver.PD3O (staged under oracle/_ref; NumPy + Numba stencils, pxrt.Width.SINGLE) on a {n}^3 fp32 "
                  f"phantom, {done} iterations through m_step() after {W} warm-up iterations (JIT), {dt:.1f} s")
    except Exception as e:  # the staged reference is missing or cannot run here: the oracle's port, reason stated
        n = args.cpu_size
        why = f"{type(e).__name__}: {e}"
        v1, threads, dt1 = run_port(n, 2, 2)
        done = int(max(K, min(400, 6.0 / max(dt1 / 2, 1e-3))))  # >= 5 s of CPU work: the ratio must not swing with CPU noise
        val, threads, dt = run_port(n, done, 0)
        kind = "port"
        sample = f"oracle/tv_oracle.c (C/OpenMP pass-by-pass port of the reference's PD3O iteration) on a {n}^3 fp32 phantom, {done} iterations, {dt:.1f} s"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "sample": f"{n}^3 sub-volume, {done} iterations"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if why:
        line["cpu_baseline"]["reference_unavailable"] = why
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
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # nvidia-smi takes a moment to start: be sampling before anything is timed
                time.sleep(0.02)
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
class Env:
    """rank / world / device, the barrier, max-over-ranks reductions."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def wrap(self, shape, comps=1):
        """How an array enters the solver: as it is on one GPU, as this rank's planes (ShardedArray) on several."""
        if self.world == 1:
            return lambda t: t
        from pyxu_b200.slab import ShardedArray

        return lambda t: ShardedArray(t, shape, comps=comps, rank=self.rank, world=self.world)

    def planes(self, n0):
        from pyxu_b200.slab import partition

        return partition(n0, self.world)[self.rank]


def tv_solver(shape, neg_y, positivity=True, lam=LAM, dtype=None):
    """The object under test -- the same at every N:  PD3O(f = 1/2||x - y||^2, g = i_+, h = lam*L21, K = Gradient)."""
    import numpy as np

    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs

    N = int(np.prod(shape))
    f = 0.5 * pxo.SquaredL2Norm(dim=N).argshift(neg_y)
    Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32 if dtype is None else dtype)
    h = lam * pxo.L21Norm(arg_shape=(len(shape), *shape), l2_axis=(0,))
    g = pxo.PositiveOrthant(dim=N) if positivity else None
    return pxs.PD3O(f=f, g=g, h=h, K=Kop, show_progress=False, final_writeback=False)


def local_phantom(env, n):
    """This rank's planes of the synthetic phantom (blocks of a 16^3 random field + noise), generated on the device."""
    torch = env.torch
    a, b = env.planes(n)
    c = max(1, n // 16)
    coarse = torch.rand((16, 16, 16), device=env.dev, dtype=torch.float32, generator=torch.Generator(device=env.dev).manual_seed(1234))
    idx = torch.arange(a, b, device=env.dev) // c
    y = coarse[idx.clamp_(max=15)].repeat_interleave(c, 1).repeat_interleave(c, 2)[:, :n, :n].contiguous()
    y += 0.1 * torch.randn(y.shape, device=env.dev, dtype=torch.float32, generator=torch.Generator(device=env.dev).manual_seed(99 + env.rank))
    return y


def timed_steps(env, step, K, W, clk=None):
    """W untimed + K timed calls of `step` between barriers; (ms max over ranks, launches, host marks)."""
    from pyxu_b200 import _cabi

    torch = env.torch
    for _ in range(W):
        step()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _cabi.launch_count()
    t_begin = clk.mark() if clk else None
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    env.barrier()
    t_end = clk.mark() if clk else None
    return env.max_over_ranks(e0.elapsed_time(e1)), int(_cabi.launch_count() - l0), (t_begin, t_end)


def parity_check(env):
    """A fixture of the REAL reference (tests/golden/slabs.npz: PD3O-TV 32x12x16 fp64, 25 iterations, rho = 1.2) replayed through
    the same solver path on the N ranks before anything is timed; fp32 replay against the float64 result as well."""
    import numpy as np

    import pyxu_b200.opt.stop as pxst

    g = np.load(os.path.join(ROOT, "tests", "golden", "slabs.npz"))
    y, shape = g["y"], (32, 12, 16)
    out = {"case": "PD3O-TV 32x12x16, 25 iterations, rho=1.2 -- tests/golden/slabs.npz:pd3o_tv3d, produced by the real reference",
           "ranks": env.world}
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))
    for tag, dt, tol in (("f64", np.float64, 1e-10), ("f32", np.float32, 1e-4)):
        slv = tv_solver(shape, -y.reshape(-1).astype(dt), dtype=dt)
        slv.fit(x0=y.reshape(-1).astype(dt), stop_crit=pxst.MaxIter(25), rho=1.2, distributed=(True if env.world > 1 else None))
        assert slv._astate.get("error") is None, slv._astate.get("error")
        assert (slv._slab is not None) == (env.world > 1)
        data, hist = slv.stats()
        ex, ez = rel(data["x"], g["pd3o_tv3d/x"]), rel(data["z"], g["pd3o_tv3d/z"])
        out[tag] = {"rel_err_x": ex, "rel_err_z": ez, "tol": tol, "iterations": int(len(hist)) - 1}
        assert ex <= tol and ez <= tol and len(hist) == int(g["pd3o_tv3d/n_hist"]), (tag, ex, ez, len(hist))
    out["ok"] = True
    return out


def _cabi_launches():
    from pyxu_b200 import _cabi

    return int(_cabi.launch_count())


def secondary_configs(env, peak, K=10, W=3):
    """BASELINE.json configs[0], [1], [2], [4] through the public solver API: ms / iteration, algorithmic bytes, fraction of the
    measured HBM peak.  [0] and [1] are single-GPU problems (reported at N = 1); [2] deals its batch out to the ranks; [4] is
    slab-decomposed (its full 2048x2048x1024 volume needs the memory of 8 GPUs: at N < 8 the z extent is 128 planes per rank)."""
    import numpy as np

    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst
    from pyxu_b200.abc import Mode
    from pyxu_b200.slab import split_batch

    torch, world, rank = env.torch, env.world, env.rank
    gen = torch.Generator(device=env.dev).manual_seed(7 + rank)
    out = {}

    def gauss(n, s):
        t = np.arange(n) - (n - 1) / 2
        k = np.exp(-0.5 * (t / s) ** 2)
        return (k / k.sum()).astype(np.float32)

    def entry(name, ms, launches, nvox, bpv, n_gpus, **kw):
        per = ms / K
        gbs = bpv * nvox / n_gpus / per / 1e6
        return dict(workload=name, n_gpus=n_gpus, ms_per_iter=per, gvoxel_iter_per_s=nvox / per / 1e6, algorithmic_bytes_per_voxel=bpv,
                    achieved_GBps_per_gpu=gbs, frac=gbs / peak, launches_per_iter=launches / K, **kw)

    if world == 1:
        # configs[0]: the reference's own CPU-runnable case, HOST arrays in and out through fit() + solution(), wall clock
        n = 512
        N = n * n
        y = np.random.default_rng(0).random(N)

        def solve():
            slv = tv_solver((n, n), -y, lam=0.1, dtype=np.float64)
            slv.fit(x0=y.copy(), stop_crit=pxst.MaxIter(200))
            return slv, slv.solution()

        solve()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slv, x = solve()
        dt = time.perf_counter() - t0
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and len(slv.stats()[1]) == 201
        # device time of the 200 iterations as fit() issues them (BLOCK mode, MaxIter: one iteration, a timed batch of 8, then the
        # rest in ONE cooperative launch with grid-wide barriers -- k_tv_tile2d_loop), CUDA events around fit() on device arrays
        yd = torch.from_numpy(y).to(env.dev)
        l0 = _cabi_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slv.fit(x0=yd, stop_crit=pxst.MaxIter(200))
        e1.record()
        torch.cuda.synchronize()
        ms_fit, launches_fit = e0.elapsed_time(e1), _cabi_launches() - l0
        slv.fit(x0=yd, mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
        ms, launches, _ = timed_steps(env, slv.m_step, 200, 20)
        out["configs[0]"] = dict(workload="2-D TV denoising 512x512 float64, PD3O, 200 iterations, host arrays through fit() + solution()",
                                 n_gpus=1, e2e_seconds=dt, e2e_iterations_per_s=200 / dt, e2e_gvoxel_iter_per_s=N * 200 / dt / 1e9,
                                 fit_device_us_per_iter=1e3 * ms_fit / 200, fit_launches=launches_fit,
                                 device_us_per_iter=1e3 * ms / 200, algorithmic_bytes_per_voxel=56,
                                 achieved_GBps_per_gpu=56 * N / (ms_fit / 200) / 1e6, launches_per_iter=launches / 200,
                                 note="2 MiB per field: the state lives in L2; latency-bound, not HBM-bound.  fit_device_us_per_iter: the iterations as fit() "
                                      "issues them (persistent cooperative kernel, grid barrier per iteration); device_us_per_iter: one launch per "
                                      "iteration driven from Python (Mode.MANUAL)")
        del slv
        # configs[1]: 2-D TV deblurring 8192^2 fp32, CondatVu, separable 9x9 Gaussian blur
        n = 8192
        N = n * n
        g9 = gauss(9, 1.7)
        Aop = pxo.Stencil(arg_shape=(n, n), kernel=[g9, g9], center=(4, 4), mode="constant")
        yd = torch.rand(N, device=env.dev, generator=gen)
        f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yd)) * Aop
        h = 0.05 * pxo.L21Norm(arg_shape=(2, n, n), l2_axis=(0,))
        slv = pxs.CondatVu(f=f, g=None, h=h, K=pxo.Gradient(arg_shape=(n, n), dtype=np.float32), beta=float(Aop.lipschitz) ** 2, show_progress=False)
        slv.fit(x0=yd, mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
        ms, launches, _ = timed_steps(env, slv.m_step, K, W)
        assert slv._plan.iter_ok is True and Aop._tiled_ok is True
        # 2 alpha (A x - y) (read x, y; write r) + A^T r (read r; write grad f) + CV iteration (read x, grad f, z0, z1; write x, z0, z1)
        out["configs[1]"] = entry("2-D TV deblurring 8192x8192 fp32, CondatVu, separable 9x9 Gaussian Stencil blur + L21 o Gradient", ms, launches, N, 12 + 8 + 28, 1)
        del slv, f, yd, Aop
        torch.cuda.empty_cache()

    # configs[2]: batch of 256 1024^2 images, FISTA, 5x5 PSF, batch split over the ranks (no communication)
    n, B = 1024, 256
    N = n * n
    lo, hi = split_batch(B, world)[rank]
    g5 = gauss(5, 1.0)
    Aop = pxo.Stencil(arg_shape=(n, n), kernel=np.outer(g5, g5), center=(2, 2), mode="constant")
    yd = torch.rand(hi - lo, N, device=env.dev, generator=gen)
    f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-yd)) * Aop
    slv = pxs.PGD(f=f, g=0.02 * pxo.L1Norm(dim=N), show_progress=False)
    slv.fit(x0=yd, mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), tau=1.0 / float(Aop.lipschitz) ** 2)
    ms, launches, _ = timed_steps(env, slv.m_step, K, W)
    assert slv._fused is not None, "the two-pass tiled FISTA form did not apply"
    # r = A y_k - b: read x, x_prev, b, write r;  x_new = prox(y_k - tau A^T r): read r, x, x_prev, write x_new
    out["configs[2]"] = entry(f"batch of 256 1024x1024 images, PGD (FISTA) L1 deconvolution, 5x5 Stencil, batch split over {world} GPU(s)", ms, launches, B * N, 32, world)
    del slv, f, yd, Aop
    torch.cuda.empty_cache()

    # configs[4]: 3-D TV deblurring, CondatVu, separable 7x7x7 PSF + positivity, z-slabs
    nz = 1024 if world >= 8 else 128 * world
    shape = (nz, 2048, 2048)
    N = int(np.prod(shape))
    a, b = env.planes(nz)
    sh = env.wrap(shape)
    y_loc = torch.rand((b - a) * 2048 * 2048, device=env.dev, generator=gen)
    g7 = gauss(7, 1.2)
    Aop = pxo.Stencil(arg_shape=shape, kernel=[g7, g7, g7], center=(3, 3, 3), mode="constant")
    f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(sh(-y_loc))) * Aop
    h = 0.05 * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,))
    slv = pxs.CondatVu(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=pxo.Gradient(arg_shape=shape, dtype=np.float32), beta=float(Aop.lipschitz) ** 2,
                       show_progress=False)
    slv.fit(x0=sh(y_loc), mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), distributed=True if world > 1 else None)
    del y_loc
    ms, launches, _ = timed_steps(env, slv.m_step, K, W)
    full = "the named 2048x2048x1024 volume" if nz == 1024 else f"z extent reduced to {nz} planes (128 per rank: the full volume needs 8 GPUs)"
    out["configs[4]"] = entry(f"3-D TV deblurring 2048x2048x{nz} fp32, CondatVu, separable 7x7x7 Stencil PSF + positivity, {world} z-slab(s); {full}",
                              ms, launches, N, 12 + 8 + 36, world)
    del slv, f, Aop
    torch.cuda.empty_cache()

    if world == 1:
        # configs[4] with a DENSE 7x7x7 PSF of full rank (a measured PSF; the reference takes any dense kernel, stencil.py:356-461): the two
        # stencil passes are then 343 FMAs per voxel each (pxb_stencil3d_dense_apply), the iteration is FMA- instead of HBM-bound.
        # An extra entry: a failure here is reported in the entry and leaves the line intact.
        try:
            rng = np.random.default_rng(11)
            ax = np.arange(7) - 3.0
            q = (ax[:, None, None] / 2.4) ** 2 + ((ax[None, :, None] - 0.6 * ax[:, None, None]) / 2.1) ** 2 + ((ax[None, None, :] + 0.6 * ax[None, :, None]) / 2.8) ** 2
            psf = np.exp(-0.5 * q) * (1 + 0.1 * rng.standard_normal((7, 7, 7)))
            psf = (psf / psf.sum()).astype(np.float32)
            Aop = pxo.Stencil(arg_shape=shape, kernel=psf, center=(3, 3, 3), mode="constant")
            y_loc = torch.rand(N, device=env.dev, generator=gen)
            f = (0.5 * pxo.SquaredL2Norm(dim=N).argshift(-y_loc)) * Aop
            slv = pxs.CondatVu(f=f, g=pxo.PositiveOrthant(dim=N), h=h, K=pxo.Gradient(arg_shape=shape, dtype=np.float32), beta=float(Aop.lipschitz) ** 2,
                               show_progress=False)
            slv.fit(x0=y_loc, mode=Mode.MANUAL, stop_crit=pxst.ManualStop())
            ms, launches, _ = timed_steps(env, slv.m_step, K, W)
            flops = 2 * 2 * 343 * N  # two dense passes per iteration
            out["configs[4] dense PSF"] = entry(f"3-D TV deblurring 2048x2048x{nz} fp32, CondatVu, DENSE 7x7x7 Stencil PSF of full rank + positivity, one GPU",
                                                ms, launches, N, 12 + 8 + 36, 1, stencil_path="marching kernel" if Aop._march3d_ok else "per-plane tiled passes",
                                                fp32_TFLOPs_per_s=flops / (ms / K) / 1e9,
                                                note="FMA-bound: 2 x 343 FMA per voxel and iteration in the two stencil passes; frac (HBM) is not the bound here")
            del slv, f, Aop, y_loc
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["configs[4] dense PSF"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args, int(os.environ.get("RANK", "0")))

    import numpy as np

    env = Env()
    torch, dist, rank, world, dev = env.torch, env.dist, env.rank, env.world, env.dev

    import pyxu_b200.opt.stop as pxst
    from pyxu_b200 import _array as A_
    from pyxu_b200.abc import Mode

    n = args.size
    shape = (n, n, n)
    nvox = n**3
    K, W = args.steps, max(args.warmup, 3)
    a0, b0 = env.planes(n)
    local_vox = (b0 - a0) * n * n
    sh = env.wrap(shape)
    dist_kw = dict(distributed=True) if world > 1 else {}

    parity = parity_check(env)  # replay a fixture of the real reference on these ranks before timing anything

    y = local_phantom(env, n).reshape(-1)
    y_host = shift_host = None
    if not args.no_e2e:  # host copies for the end-to-end run (made outside every timed region)
        y_host = torch.empty(local_vox, dtype=torch.float32, pin_memory=True)
        y_host.copy_(y)
        shift_host = torch.empty(local_vox, dtype=torch.float32, pin_memory=True)
        torch.neg(y_host, out=shift_host)
        A_.reserve_host_results(4 * local_vox)  # solution() lands in a pinned buffer reserved ahead of time (a serving loop does this once)

    pending = []

    def probe(tag):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        pending.append((tag, ev))

    def durations(kinds):
        tags = {}
        for (t0, a), (t1, b) in zip(pending[:-1], pending[1:]):
            for (ta, tb), name in kinds.items():
                if t0 == ta and t1 == tb:
                    tags.setdefault(name, []).append(a.elapsed_time(b))
        return tags

    peak, peak_src = measured_peak()

    def roof_entry(kname, ts, note=None):
        avg = sum(ts) / len(ts)
        bytes_alg = BYTES_PER_VOXEL[kname] * local_vox
        ach = bytes_alg / (avg * 1e-3) / 1e9
        r = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": ncu_traffic(kname, local_vox), "avg_ms": avg, "algorithmic_bytes": bytes_alg,
             "peak_source": peak_src, "launches_timed": len(ts)}
        if note:
            r["note"] = note
        return r

    # ---- (1) device-resident iterations, no stopping criterion: the headline `value` ----------------------------------
    slv = tv_solver(shape, sh(-y))
    slv.fit(x0=sh(y), mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), **dist_kw)
    assert slv._plan.kind == "fused" and (slv._slab is not None) == (world > 1)
    if args.two_sweep:
        slv._plan.iter_ok = False
        if slv._slab is not None:
            slv._slab.fused = False
    exchange = None
    if slv._slab is not None:
        exchange = ("stored into the neighbours' ghost planes by the iteration kernel itself (peer memory over NVLink, pxb_pds_iter_p2p), edge chunks first"
                    if slv._slab.p2p is not None else "NCCL send/recv on a high-priority stream while the interior is computed")
    slv._probe = probe
    with ClockSampler(env.local) as clk:  # nvidia-smi is started before the warm-up so that it is sampling when the timed region begins
        for _ in range(W):
            slv.m_step()
        pending.clear()
        ms, launches, (t_begin, t_end) = timed_steps(env, slv.m_step, K, 0, clk)
    value = nvox * K / (ms * 1e-3) / 1e9
    kinds = {("primal_begin", "primal_end"): "pxb_pds_primal", ("primal_end", "dual_end"): "pxb_pds_dual", ("iter_begin", "iter_end"): "pxb_pds_iter"}
    roof = {k: roof_entry(k, ts) for k, ts in durations(kinds).items()}
    dominant = max(roof.values(), key=lambda r: r["avg_ms"] * r["launches_timed"]) if roof else None
    slv._probe = None
    del slv
    torch.cuda.empty_cache()

    # ---- (2) the same iteration driven by the reference's default criterion RelError[x] & RelError[z] (pds.py default_stop_crit):
    #          Solver._step() = criterion test on the fused norms (32-byte readback, all-reduced on N ranks) + history + m_step -----
    crit = pxst.MaxIter(10**9) | (pxst.RelError(eps=1e-30, var="x") & pxst.RelError(eps=1e-30, var="z"))
    slv = tv_solver(shape, sh(-y))
    slv.fit(x0=sh(y), mode=Mode.MANUAL, stop_crit=crit, **dist_kw)
    assert "_fused_norms" in slv._mstate
    slv._probe = probe
    for _ in range(W):
        slv._step()
    pending.clear()
    ms_c, launches_c, _ = timed_steps(env, slv._step, K, 0)
    if not args.two_sweep:
        ts = durations({("iter_begin", "iter_end"): "pxb_pds_iter+criterion"}).get("pxb_pds_iter+criterion")
        if ts:
            roof["pxb_pds_iter+criterion"] = roof_entry("pxb_pds_iter+criterion", ts, note="default stopping criterion RelError[x] & RelError[z] fused: x written and the previous x re-read every iteration")
            roof["pxb_pds_iter+criterion"].update(step_ms=ms_c / K, gvoxel_iter_per_s=nvox * K / (ms_c * 1e-3) / 1e9,
                                                  what="Solver._step(): criterion on the fused sums (one 32-byte readback per step) + history record + m_step, CUDA events, max over ranks")
    slv._probe = None
    del slv
    torch.cuda.empty_cache()

    # ---- (3) end to end through the public API with HOST buffers -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del y
        torch.cuda.empty_cache()
        shift_np, x0_np = shift_host.numpy(), y_host.numpy()  # pinned host memory seen as NumPy arrays
        runs = []
        for rep in range(2):  # [0] pays the one-off costs (first use of the copy engines, pools, NCCL channels); [1] is the reported one
            env.barrier()
            t0 = time.perf_counter()
            slv2 = tv_solver(shape, sh(shift_np))
            t1 = time.perf_counter()
            # K iterations; the RelError metric is read back from the device every iteration (eps tiny: never triggers)
            slv2.fit(x0=sh(x0_np), stop_crit=pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x"), **dist_kw)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            x_host = slv2.solution()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            env.barrier()
            dt = env.max_over_ranks(time.perf_counter() - t0)
            assert slv2._astate.get("error") is None, slv2._astate.get("error")
            assert slv2._plan.kind == "fused" and "_fused_norms" in slv2._mstate and sum(len(h_) for h_ in slv2._astate["history"]) == K + 1
            x_loc = getattr(x_host, "local", x_host)
            assert isinstance(x_loc, np.ndarray) and x_loc.size == local_vox and np.isfinite(x_loc[:: max(1, local_vox // 1000)]).all()
            tm = slv2._astate["timing"]
            runs.append({"seconds": dt, "build_s": t1 - t0, "fit_s": t2 - t1, "fit_m_init_s": tm["m_init_s"], "fit_iterations_s": tm["run_s"],
                         "solution_s": t3 - t2, "final_barrier_s": time.perf_counter() - t3})
            hist_e2e = np.concatenate(slv2._astate["history"])["RelError[x]"].copy()
            streamed = getattr(slv2, "_slab", None) is not None and world == 1
            if rep == 0:
                del x_host, x_loc
            del slv2
        # Full-size check of the end-to-end result (outside every timed region): the same K iterations on DEVICE-resident arrays
        # through the ordinary one-launch-per-iteration / batched path; result and RelError history must agree with what the
        # host-array run (streamed wavefront on one GPU) brought back.
        yd = torch.from_numpy(x0_np).to(dev)
        slvc = tv_solver(shape, sh(-yd))
        slvc.fit(x0=sh(yd), stop_crit=pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x"), **dist_kw)
        xc = slvc.solution()
        xc = getattr(xc, "local", xc)
        xe = torch.from_numpy(x_loc).to(dev)
        num, den = torch.linalg.vector_norm((xe - xc.reshape(-1)).double()), torch.linalg.vector_norm(xc.double())
        rel_chk = env.max_over_ranks(float(num / den))
        hist_c = np.concatenate(slvc._astate["history"])["RelError[x]"]
        hist_dev = float(np.max(np.abs(hist_e2e[1:] - hist_c[1:]) / np.maximum(np.abs(hist_c[1:]), 1e-30)))
        assert rel_chk < 1e-5 and hist_dev < 1e-3, (rel_chk, hist_dev)
        e2e_check = {"rel_err_x_vs_device_resident_run": rel_chk, "max_rel_dev_of_RelError_history": hist_dev, "tol": [1e-5, 1e-3],
                     "streamed_wavefront": bool(streamed), "voxels_compared": int(local_vox)}
        del slvc, xc, xe, yd, x_host, x_loc
        torch.cuda.empty_cache()
        dt = runs[1]["seconds"]
        e2e = {"value": nvox * K / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(2 * 4 * nvox / K),
               "d2h_bytes_per_step": int(4 * nvox / K + 16 * world), "seconds": dt, "parts": runs[1], "first_call": runs[0], "check": e2e_check,
               "what": "per rank: PD3O(...).fit(x0=<pinned host array>, stop_crit=MaxIter(K)|RelError[x]) + solution(): H2D of x0 and of the data y "
                       "(this rank's planes), K fused iterations with RelError[x] tested after every one of them (1 GPU: x0 and y travel in 16-plane "
                       "z-chunks and the iterations are queued as a wavefront behind the chunks, the result travels back behind the wave, the "
                       "criterion is replayed from the sums every iteration left on the device -- SlabTV.run_streamed; N ranks: one upload, "
                       "sums all-reduced and read back every step), D2H of x into a "
                       "pinned result buffer reserved beforehand (reserve_host_results); wall clock between barriers, max over ranks; second of two "
                       "calls (the first one, which also pays one-off initialisation, is `first_call`)"}
        A_.release_host_results()
        # What a caller gets who hands over ordinary (pageable) NumPy arrays and has reserved nothing: uploads through the pinned staging
        # buffers (host threads), the result into a fresh pageable array.  One call, reported beside the pinned number.
        x0_pg, shift_pg = np.array(x0_np, copy=True), np.array(shift_np, copy=True)
        env.barrier()
        t0 = time.perf_counter()
        slv3 = tv_solver(shape, sh(shift_pg))
        slv3.fit(x0=sh(x0_pg), stop_crit=pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x"), **dist_kw)
        x_pg = slv3.solution()
        torch.cuda.synchronize()
        env.barrier()
        dt_pg = env.max_over_ranks(time.perf_counter() - t0)
        x_pg = getattr(x_pg, "local", x_pg)
        assert isinstance(x_pg, np.ndarray) and np.isfinite(x_pg[:: max(1, local_vox // 1000)]).all()
        e2e["pageable"] = {"value": nvox * K / dt_pg / 1e9, "unit": UNIT, "seconds": dt_pg,
                           "what": "the same call with pageable NumPy arrays in and out and no reserved result buffer (staged copies, host threads)"}
        del slv3, x_pg, x0_pg, shift_pg
        del y_host, shift_host, shift_np, x0_np
        torch.cuda.empty_cache()

    configs = None
    if not args.no_configs:
        if world > 1:
            from pyxu_b200 import slab

            slab.release_pool()  # the headline volume's pooled slab buffers
            env.barrier()
        configs = secondary_configs(env, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ncpu = args.cpu_size
        v1, thr, dt1 = run_port(ncpu, 2, 1)
        iters = int(max(3, min(400, 12.0 / max(dt1 / 2, 1e-3))))  # ~12 s of CPU work
        v, thr, dtc = run_port(ncpu, iters, 0)
        port = {"value": v, "unit": UNIT, "cores": thr, "kind": "port",
                "sample": f"oracle/tv_oracle.c (C/OpenMP pass-by-pass port of the reference's PD3O iteration) on a {ncpu}^3 fp32 phantom, {iters} iterations, {dtc:.1f} s"}
        try:
            rv, rthr, rdt, rdone = run_reference(args.ref_size, 12, 2, budget_s=20.0)
            cpu = {"value": rv, "unit": UNIT, "cores": rthr, "kind": "reference",
                   "sample": f"the real pyxu.opt.solver.PD3O (oracle/_ref, NumPy + Numba, single precision) on a {args.ref_size}^3 fp32 phantom, {rdone} iterations, {rdt:.1f} s",
                   "port": port}
        except Exception as e:
            cpu = dict(port, reference_unavailable=f"{type(e).__name__}: {e}")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "solver": "pyxu_b200.opt.solver.PD3O(f, g, h, K).fit() -- the same object at every N",
                       "decomposition": "single GPU" if world == 1 else f"{world} z-slabs behind Solver.fit(); boundary planes of the new iterate (5 planes per interface): {exchange}",
                       "l2_policy": f"inputs larger than L2: {4 * nvox / world / 2**20:.0f} MiB per field per GPU vs 126 MB L2",
                       "iterations_per_step": 1},
            "clocks": clk.summary(t_begin, t_end), "e2e": e2e, "gpu_launches": launches, "parity": parity,
            "roofline": dominant, "roofline_all": roof, "configs": configs, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        from pyxu_b200 import slab

        slab.release_pool()  # drop the mappings of the neighbours' buffers before anybody's process ends
        env.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
