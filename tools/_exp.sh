python tools/profile_run.py --sites 1000000 --brief --max-iter 3 > /dev/null
FASTLEM_TRACE=1 python tools/_trace.py 1000000 2>&1 | grep "rooting\|rep"
python tools/profile_run.py --sites 1000000 --brief --repeat 2
python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt rebuild_growth=8
python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt rebuild_growth=16
python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt rebuild_growth=2
