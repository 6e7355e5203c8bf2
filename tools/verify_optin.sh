#!/bin/bash
# The paths of this library that are switched off until they have run on a GPU (sections 3b and 4a): their GPU
# tests and A/B timings in one go.  On the GPU box:   bash tools/verify_optin.sh   (one B200; add a 2-GPU run for the slab form)
#   PYXU_B200_STENCIL_PADDED=1    Stencil with a folding boundary mode through Pad -> tiled stencil / tiled stencil -> Pad^T
#   PYXU_B200_SLAB_FUSED_MODES=1  SlabPD3OTV: single-kernel iteration with folding boundary modes on slabs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYXU_B200_STENCIL_PADDED=1 PYXU_B200_SLAB_FUSED_MODES=1
python -m pytest tests/test_gpu_zz_stencil_padded.py tests/test_gpu_zz_iter_modes.py tests/test_gpu_slab.py tests/test_gpu_solvers.py tests/test_gpu_operators.py \
    -m gpu -q --runxfail 2>&1 | tee gpurun_out/optin_pytest.log | tail -5   # --runxfail: the opt-in tests count as ordinary tests here
python tools/bench_modes.py 2>&1 | tee gpurun_out/optin_bench_modes.log
python tools/bench_stencil.py 2>&1 | tee gpurun_out/optin_bench_stencil.log
