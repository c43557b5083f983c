set -x
python tools/profile_run.py --sites 1000000 --brief --max-iter 5 > /dev/null   # builds the workload cache
python tools/flow_stats.py 1000000
python tools/flow_stats.py 1000000 incremental=0
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 600 --csv --log-file gpurun_out/incr_launches.csv python tools/profile_run.py --sites 1000000 --brief --max-iter 420
