"""GPU tests of the z-slab path behind Solver.fit() (tests/slab_worker.py: every case is checked against fixtures of the real
reference or the NumPy oracle): world_size 1 always (buffer layout with ghost planes, sub-range descriptors); world_size 2 / 4 / 8
(NCCL halo exchange over NVLink) when the box has the GPUs.  The worker's report is kept under gpurun_out/ when that exists."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, "tests", "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"slab_worker_world{world}.log"), "w") as fh:
            fh.write(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    from pyxu_b200.operator.linop import stencil as st_mod

    n = (12 if world <= 3 else 11) + int(st_mod.DENSE3D_MARCH)  # (+ the dense-PSF case when the marching kernel is selected)
    assert f"{n}/{n} cases OK" in r.stdout and "FAIL" not in r.stdout


def test_slab_world1():
    _run(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slab_world2():
    _run(2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_slab_world4():
    _run(4)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs")
def test_slab_world8():
    _run(8)
