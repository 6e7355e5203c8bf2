# one-GPU verification of the dense marching stencil kernel: GPU suite, timings against the per-plane passes and the gather kernel, one ncu capture
#   gpurun --timeout 200 -- 'bash tools/gpu_call_dense3d.sh'
mkdir -p gpurun_out
(timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_dense3d.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_dense3d.log)
tail -3 gpurun_out/pytest_gpu_dense3d.log
timeout 60 python tools/bench_stencil3d.py --dense-only > gpurun_out/bench_dense3d.txt 2>&1; echo "bench rc=$?"; grep -v "^{" gpurun_out/bench_dense3d.txt | tail -20
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_stencil3d_dense -c 1 -o gpurun_out/dense3d_7 python tools/bench_stencil3d.py --dense-only --taps 7 --reps 1 --only march --no-tiled > gpurun_out/ncu_dense3d.log 2>&1; echo "ncu rc=$?"
