python -m pytest tests/test_gpu_iter.py tests/test_gpu_solvers.py tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -2
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --kernel-name regex:k_tv_iter_tma --launch-skip 4 --launch-count 1 python tools/bench_criterion.py --n0 256 --reps 2 --only plain 2>&1 | grep -E "inst_executed|duration"
python tools/probe_sustained.py | tail -1 | cut -c1-600
