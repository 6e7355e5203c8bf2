#!/usr/bin/env python
"""
Host <-> device facts behind the end-to-end numbers (run on the GPU box; prints one JSON object):
pinned / pageable copy bandwidths, the cost of pinning, cudaMalloc of multi-GiB blocks, first touch of a fresh
NumPy array, and the timing breakdown of PD3O.fit(x0=<host>) + solution() at the headline size.

    python tools/probe_e2e.py [--size 1024] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def t_(fn, sync=True):
    if sync:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    if sync:
        torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--gib", type=float, default=4.0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    nbytes = int(args.gib * (1 << 30))
    gb = nbytes / 1e9

    dt, d = t_(lambda: torch.empty(nbytes, dtype=torch.uint8, device=dev))
    out["cudaMalloc_s_per_GiB(first)"] = dt / args.gib
    dt, d2 = t_(lambda: torch.empty(nbytes, dtype=torch.uint8, device=dev))
    out["cudaMalloc_s_per_GiB"] = dt / args.gib
    del d2

    dt, pin = t_(lambda: torch.empty(nbytes, dtype=torch.uint8, pin_memory=True))
    out["cudaHostAlloc_GBps(first)"] = gb / dt
    dt, pin2 = t_(lambda: torch.empty(nbytes, dtype=torch.uint8, pin_memory=True))
    out["cudaHostAlloc_GBps"] = gb / dt
    del pin2
    dt, pin2 = t_(lambda: torch.empty(nbytes, dtype=torch.uint8, pin_memory=True))
    out["cudaHostAlloc_GBps(cached block)"] = gb / dt

    for tag, fn in (("h2d_pinned", lambda: d.copy_(pin, non_blocking=True)), ("d2h_pinned", lambda: pin.copy_(d, non_blocking=True))):
        fn()
        best = min(t_(fn)[0] for _ in range(3))
        out[f"{tag}_GBps"] = gb / best

    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d.copy_(pin, non_blocking=True)
        with torch.cuda.stream(s2):
            pin2.copy_(d, non_blocking=True)

    both()
    best = min(t_(both)[0] for _ in range(3))
    out["h2d+d2h_concurrent_GBps_each"] = gb / best

    # chunked pinned copies (64 MiB) to see per-copy overheads
    ch = 64 << 20

    def chunked():
        for o in range(0, nbytes, ch):
            d[o : o + ch].copy_(pin[o : o + ch], non_blocking=True)

    chunked()
    out["h2d_pinned_64MiB_chunks_GBps"] = gb / min(t_(chunked)[0] for _ in range(2))

    # pageable
    dt, pg = t_(lambda: np.empty(nbytes, dtype=np.uint8), sync=False)
    dt, _ = t_(lambda: pg.fill(1), sync=False)
    out["first_touch_GBps"] = gb / dt
    dt, _ = t_(lambda: pg.fill(2), sync=False)
    out["memset_warm_GBps(1 thread)"] = gb / dt
    tp = torch.from_numpy(pg)
    dt, _ = t_(lambda: d.copy_(tp))
    out["h2d_pageable_torch_GBps"] = gb / dt
    dt, _ = t_(lambda: tp.copy_(d))
    out["d2h_pageable_torch_GBps"] = gb / dt
    # cudaHostRegister of a pre-faulted array
    cudart = torch.cuda.cudart()
    dt, rc = t_(lambda: cudart.cudaHostRegister(pg.ctypes.data, nbytes, 0))
    out["cudaHostRegister_GBps(prefaulted)"] = gb / dt
    out["cudaHostRegister_rc"] = int(rc)
    if int(rc) == 0:
        dt, _ = t_(lambda: d.copy_(tp, non_blocking=True))
        out["h2d_registered_GBps"] = gb / dt
        dt, _ = t_(lambda: cudart.cudaHostUnregister(pg.ctypes.data))
        out["cudaHostUnregister_s"] = dt
    fresh = np.empty(nbytes, dtype=np.uint8)
    dt, rc = t_(lambda: cudart.cudaHostRegister(fresh.ctypes.data, nbytes, 0))
    out["cudaHostRegister_GBps(fresh pages)"] = gb / dt
    if int(rc) == 0:
        tf = torch.from_numpy(fresh)
        dt, _ = t_(lambda: tf.copy_(d, non_blocking=True))
        out["d2h_registered_fresh_GBps"] = gb / dt
        cudart.cudaHostUnregister(fresh.ctypes.data)
    del fresh, pg, tp, d, pin, pin2
    torch.cuda.empty_cache()

    # pipelined pageable paths of the library
    from pyxu_b200 import _array as A

    n = args.size
    nvox = n**3
    y = np.random.default_rng(0).random(nvox, dtype=np.float32)
    for rep in range(2):
        dt, yd = t_(lambda: A.asdevice(y)[0])
        out[f"A.asdevice(pageable {nvox * 4 / 2**30:.0f} GiB)_GBps[{rep}]"] = nvox * 4 / 1e9 / dt
        dt, yh = t_(lambda: A.restore(yd, A.HOST))
        out[f"A.restore(-> pageable)_GBps[{rep}]"] = nvox * 4 / 1e9 / dt
        assert np.array_equal(yh[:: nvox // 997], y[:: nvox // 997])
        del yh
    del yd
    torch.cuda.empty_cache()

    # end-to-end breakdown at the headline size, twice (cold / warm pools)
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

    shape = (n, n, n)
    yp = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
    yp.copy_(torch.from_numpy(y))
    sp = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
    torch.neg(yp, out=sp)
    Kop = pxo.Gradient(arg_shape=shape, dtype=np.float32)
    K = args.steps
    for rep in range(3):
        for src_kind, (x0, sh) in (("pinned", (yp.numpy(), sp.numpy())), ("pageable", (y, -y))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f2 = 0.5 * pxo.SquaredL2Norm(dim=nvox).argshift(sh)
            slv = pxs.PD3O(f=f2, g=pxo.PositiveOrthant(dim=nvox), h=0.08 * pxo.L21Norm(arg_shape=(3, *shape), l2_axis=(0,)), K=Kop,
                           show_progress=False, final_writeback=False)
            t1 = time.perf_counter()
            slv.fit(x0=x0, stop_crit=pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x"))
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            xh = slv.solution()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            out[f"e2e[{src_kind},{rep}]"] = {"build": t1 - t0, "fit": t2 - t1, "solution": t3 - t2, "total": t3 - t0,
                                             "Gvoxel_iter_s": nvox * K / (t3 - t0) / 1e9}
            del slv, f2, xh
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
