#!/usr/bin/env python
"""Anatomy of the streamed host-array fit (SlabTV.run_streamed) on one GPU: when, relative to the first queued copy, the upload ends,
the last chunk task ends and the last chunk of the result has reached the host -- for several chunk heights.

    python tools/probe_stream.py [--size 1024] [--iters 20] [--planes 8,16,32] [--crit 1]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import pyxu_b200.opt.solver as pxs
import pyxu_b200.opt.stop as pxst
from pyxu_b200 import _array as A_

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--planes", default="8,16,32")
ap.add_argument("--crit", type=int, default=1)
args = ap.parse_args()
n, K = args.size, args.iters
shape, N = (n, n, n), n**3
y = torch.rand(N, dtype=torch.float32, pin_memory=True)
shift = torch.empty(N, dtype=torch.float32, pin_memory=True)
torch.neg(y, out=shift)
A_.reserve_host_results(4 * N)
PDS = pxs.PD3O.__mro__[1]
rows = []
for planes in [int(v) for v in args.planes.split(",")]:
    PDS._STREAM_PLANES = planes
    for rep in range(2):
        slv = bench.tv_solver(shape, shift.numpy())
        slv._probe = lambda tag: None  # switches the marks on
        crit = pxst.MaxIter(K) | pxst.RelError(eps=1e-30, var="x") if args.crit else pxst.MaxIter(K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        real = type(slv._slab).run_streamed if slv._slab is not None else None
        slv.fit(x0=y.numpy(), stop_crit=crit)
        eng = slv._slab
        x = slv.solution()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        m = dict(eng._stream_marks)
        row = {"planes": planes, "rep": rep, "wall_s": dt, "host_issue_s": slv._astate["timing"]["run_s"]}
        for k in ("upload_end", "compute_end", "download_end"):
            if k in m:
                row[k + "_ms"] = m["begin"].elapsed_time(m[k])
        rows.append(row)
        print(json.dumps(row), flush=True)
        del slv, x, eng
print(json.dumps({"size": n, "iters": K, "criterion": bool(args.crit), "rows": rows}))
