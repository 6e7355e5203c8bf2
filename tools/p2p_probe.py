#!/usr/bin/env python
"""
Peer-memory plumbing check (2+ GPUs): every rank exports a buffer (CUDA IPC through torch's tensor sharing), maps its
neighbours' (pyxu_b200.slab.PeerBuffers), then a kernel launched on ITS device (pxb_lincomb) writes into the upper neighbour's
buffer; after a barrier every rank checks what its lower neighbour wrote.  Prints one line per rank.

    torchrun --nproc-per-node 2 tools/p2p_probe.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pyxu_b200 import _array as A
from pyxu_b200 import _cabi as K
from pyxu_b200 import slab


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 20
    lo, hi = (rank - 1) % world, (rank + 1) % world
    pb = slab.PeerBuffers(lambda: dict(buf=torch.zeros(n, device="cuda"), flags=torch.zeros(4, dtype=torch.int32, device="cuda")), rank, world, lo, hi)
    mine = torch.full((n,), float(rank + 1), device="cuda")
    torch.cuda.synchronize()
    dist.barrier()
    dst = pb.hi["buf"]
    print(f"[p2p] rank {rank}: my buf {pb.bufs['buf'].data_ptr():#x} on {pb.bufs['buf'].device}, upper neighbour's mapped at {dst.data_ptr():#x} (tensor says {dst.device})", flush=True)
    # out = 1.0 * mine, written by a kernel of THIS device into the neighbour's memory
    rc = K.lib().pxb_lincomb(K.F32, n, C.c_void_p(dst.data_ptr()), 1.0, A.ptr(mine), 0.0, None, 0, 0.0, None, 0, A.stream())
    K.check(rc, "pxb_lincomb")
    torch.cuda.synchronize()
    dist.barrier()
    got = pb.bufs["buf"]
    ok = bool((got == float(lo + 1)).all())
    print(f"[p2p] rank {rank}: lower neighbour {lo} wrote {float(got[0])} (expected {lo + 1}.0) {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
