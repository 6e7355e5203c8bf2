# one-GPU verification of the dense marching stencil kernel: GPU suite, A/B timings, one ncu capture
mkdir -p gpurun_out
(timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/r02zz_pytest_gpu_march.log 2>&1; echo "rc=$?" >> gpurun_out/r02zz_pytest_gpu_march.log)
tail -3 gpurun_out/r02zz_pytest_gpu_march.log
timeout 60 python tools/bench_stencil3d.py --dense-only --only march > gpurun_out/r02zz_bench_dense3d_b.txt 2>&1; echo "bench rc=$?"; grep -v "^{" gpurun_out/r02zz_bench_dense3d_b.txt | tail -20
timeout 75 ncu --set full --clock-control none --import-source on -k regex:k_stencil3d_dense -c 1 -o gpurun_out/r02zz_dense3d_7b python tools/bench_stencil3d.py --dense-only --taps 7 --reps 1 --only march > gpurun_out/r02zz_ncu_b.log 2>&1; echo "ncu rc=$?"
