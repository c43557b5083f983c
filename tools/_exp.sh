set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "first_iteration or generate_to_convergence or one_million or golden" 2>&1 | tail -3
timeout 120 python tools/profile_run.py --sites 1000000 --brief --repeat 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1000 --csv --log-file gpurun_out/r1b_launches_1M.csv python bench.py --steps 1 --warmup 0 --max-iter 70 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
