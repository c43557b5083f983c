#!/bin/bash
# one GPU visit: raster timing + ncu capture, interp GPU tests, bench, smoke, then the full GPU suite if time is left
mkdir -p gpurun_out
echo "== profile_raster"; timeout 90 python tools/profile_raster.py 250000 2048 3 2>&1 | tee gpurun_out/r1c_raster_timing.txt
echo "== profile_raster 1M/4096"; timeout 120 python tools/profile_raster.py 1000000 4096 3 2>&1 | tee -a gpurun_out/r1c_raster_timing.txt
echo "== ncu raster"; timeout 150 ncu --set full --clock-control none --import-source on -k k_nn_raster -c 1 -f -o gpurun_out/r1c_raster python tools/profile_raster.py 250000 2048 1 > gpurun_out/r1c_ncu_raster.log 2>&1; tail -2 gpurun_out/r1c_ncu_raster.log
echo "== pytest interp gpu"; timeout 200 python -m pytest tests/test_interp.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r1c_pytest_gpu_interp.txt
echo "== bench"; timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/r1c_bench_1M.json 2> gpurun_out/r1c_bench_1M.err; tail -c 3000 gpurun_out/r1c_bench_1M.json; tail -3 gpurun_out/r1c_bench_1M.err
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r1c_smoke.txt
echo "== pytest gpu all"; timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r1c_pytest_gpu.txt
