set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1b_pytest_gpu.txt; cat gpurun_out/r1b_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r1b_bench_1M.json 2> gpurun_out/r1b_bench_1M.err; cut -c1-700 gpurun_out/r1b_bench_1M.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_reference.json 2>&1; cut -c1-400 gpurun_out/r1b_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 600 --csv --log-file gpurun_out/r1b_launches_1M.csv python tools/profile_run.py --sites 1000000 --brief --max-iter 420 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
