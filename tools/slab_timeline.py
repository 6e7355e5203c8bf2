#!/usr/bin/env python
"""
Where an iteration of the z-slab decomposed PD3O-TV solve spends its time on every rank (CUDA events on the streams the work
runs on + host clocks), to name what separates N-GPU scaling from ideal:

    torchrun --nproc-per-node N tools/slab_timeline.py [--size 1024] [--steps 30]

Per rank and averaged over the steps: the two boundary launches, the interior launch, the NCCL exchange on the side stream
(start relative to the end of the boundary launches, duration), the whole step on the device, and the host time it takes to
ISSUE a step (if that exceeds the device time the loop is host-bound).  Rank 0 prints one JSON object with every rank's row.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    env = bench.Env()
    import pyxu_b200.opt.stop as pxst
    from pyxu_b200.abc import Mode

    n, K = args.size, args.steps
    shape = (n, n, n)
    sh = env.wrap(shape)
    y = bench.local_phantom(env, n).reshape(-1)
    slv = bench.tv_solver(shape, sh(-y))
    slv.fit(x0=sh(y), mode=Mode.MANUAL, stop_crit=pxst.ManualStop(), **(dict(distributed=True) if env.world > 1 else {}))
    log = []

    def probe(tag):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()  # on the stream that is current where the engine ticks (main, or the side stream inside the exchange)
        log.append((tag, ev))

    for _ in range(5):
        slv.m_step()
    env.barrier()
    slv._probe = probe
    host = []
    t_all0 = time.perf_counter()
    for _ in range(K):
        t0 = time.perf_counter()
        slv.m_step()
        host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t_all0
    rows = {}
    steps = []
    cur = {}
    for tag, ev in log:
        if tag == "iter_begin":
            cur = {}
        cur[tag] = ev
        if tag == "iter_end":
            steps.append(cur)
    def avg(a, b):
        v = [s[a].elapsed_time(s[b]) for s in steps if a in s and b in s]
        return (sum(v) / len(v)) if v else None
    row = {"rank": env.rank, "planes": slv._slab.n0 if slv._slab is not None else n,
           "step_ms": avg("iter_begin", "iter_end"), "edges_ms": avg("iter_begin", "edges_end"), "interior_ms": avg("edges_end", "interior_end"),
           "exchange_start_after_edges_ms": avg("edges_end", "exchange_begin"), "exchange_ms": avg("exchange_begin", "exchange_end"),
           "tail_after_interior_ms": avg("interior_end", "iter_end"),
           "host_issue_ms_per_step": 1e3 * float(np.median(host)), "wall_ms_per_step": 1e3 * t_all / K}
    if env.world > 1:
        allrows = [None] * env.world
        dist.all_gather_object(allrows, row)
    else:
        allrows = [row]
    if env.rank == 0:
        print(json.dumps({"workload": bench.workload_name(n), "n_gpus": env.world, "steps": K,
                          "exchange": "peer memory (pxb_pds_iter_p2p)" if (slv._slab is not None and slv._slab.p2p is not None) else "NCCL send/recv",
                          "ranks": allrows}))
    if env.world > 1:
        from pyxu_b200 import slab

        del slv
        slab.release_pool()
        env.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
