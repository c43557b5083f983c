set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_capi.py -m gpu -x -q -k "first_iteration or generate_to_convergence or one_million or golden or flood or options or capi or host" 2>&1 | tail -3
FASTLEM_TRACE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v5.json')); print(d['value'], d['e2e'])"; grep "set_graph\|destroy\|flood: alloc" gpurun_out/bench_v5.err | tail -8
