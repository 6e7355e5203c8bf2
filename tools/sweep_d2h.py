#!/usr/bin/env python
"""Times pyxu_b200._array.restore (pipelined D2H of a 4 GiB result) over host thread counts / chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyxu_b200 import _array as A

t = torch.rand(1 << 30, device="cuda")
print("cpus", os.cpu_count(), flush=True)
for chunk in (16, 32, 64):
    for th in (4, 8, 12, 16, 24):
        os.environ["PXB_D2H_THREADS"] = str(th)
        A._CHUNK, A._STAGE = chunk << 20, None
        A.restore(t[: 1 << 26], A.HOST)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = A.restore(t, A.HOST)
        dt = time.perf_counter() - t0
        print(f"chunk {chunk} MiB threads {th:2d}: {dt:.3f} s  {4.295 / dt:.1f} GB/s", flush=True)
        del out
torch.cuda.synchronize(); t0 = time.perf_counter(); o = t.cpu(); print("plain .cpu():", time.perf_counter() - t0)
