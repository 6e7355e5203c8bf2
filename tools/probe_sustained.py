#!/usr/bin/env python
"""Burst vs sustained (power-capped) rate of a plain device copy and of the single-kernel PD3O-TV iteration on one GPU, with the SM clock
sampled meanwhile: is the iteration's sustained slow-down the memory system's or the SMs'?   python tools/probe_sustained.py"""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pyxu_b200.operator as pxo
from pyxu_b200 import _cabi as K

lib = K.lib()
samples = []
stop = threading.Event()


def sampler():
    while not stop.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout
            a, b = o.strip().split(",")
            samples.append((time.perf_counter(), float(a), float(b)))
        except Exception:
            pass
        time.sleep(0.05)


threading.Thread(target=sampler, daemon=True).start()


def run(fn, seconds):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # burst: 10 calls after a pause
    time.sleep(1.0)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 10
    # sustained: back to back for `seconds`, timing the last quarter
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 0.75 * seconds:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    ta = time.perf_counter()
    e0.record()
    m = 0
    while time.perf_counter() - ta < 0.25 * seconds:
        for _ in range(20):
            fn()
        m += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    tb = time.perf_counter()
    sm = [s for (t, s, p) in samples if ta <= t <= tb]
    pw = [p for (t, s, p) in samples if ta <= t <= tb]
    return burst, e0.elapsed_time(e1) / m, (float(np.median(sm)) if sm else None), (float(np.median(pw)) if pw else None)


out = {"PXB_EXP": os.environ.get("PXB_EXP")}
if os.environ.get("PXB_SKIP_COPY"):
    pass
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
bu, su, sm, pw = run(lambda: b.copy_(a), 6.0)
nbytes = 2 * a.numel() * 2
out["copy"] = {"burst_GBps": nbytes / bu / 1e6, "sustained_GBps": nbytes / su / 1e6, "sm_mhz_sustained": sm, "power_w": pw}
del a, b
shape = (512, 1024, 1024)
N = int(np.prod(shape))
y = torch.rand(N, device="cuda")
shift = -y
P = K.PdsParams()
P.tau, P.sigma, P.rho = 0.28, 0.28, 1.0
P.g = K.ProxSpec(K.PROX_POS, 0, 0.0, 0.0)
f = K.FTerm()
f.kind, f.alpha, f.shift, f.shift_period = K.F_SQL2, 0.5, shift.data_ptr(), shift.numel()
P.f = f
P.hkind, P.lam = K.DUAL_L21, 0.08
u0, u1 = y.clone(), torch.empty_like(y)
z0, z1 = torch.zeros(3 * N, device="cuda"), torch.empty(3 * N, device="cuda")
d = pxo.Gradient(arg_shape=shape, dtype=np.float32)._desc(1, K.F32)


def pair():
    K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u0.data_ptr(), z0.data_ptr(), u1.data_ptr(), z1.data_ptr(), None, None, None, None), "iter")
    K.check(lib.pxb_pds_iter(K.ALGO_PD3O, C.byref(d), C.byref(P), u1.data_ptr(), z1.data_ptr(), u0.data_ptr(), z0.data_ptr(), None, None, None, None), "iter")


bu, su, sm, pw = run(pair, 6.0)
out["pd3o_iteration_512x1024x1024"] = {"burst_GBps": 2 * 36 * N / bu / 1e6, "sustained_GBps": 2 * 36 * N / su / 1e6, "burst_ms": bu / 2, "sustained_ms": su / 2,
                                       "sm_mhz_sustained": sm, "power_w": pw}
stop.set()
print(json.dumps(out))
