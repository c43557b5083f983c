set -x
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/profile_run.py --sites 1000000 --brief
python tools/flow_timeline.py 1000000 300
