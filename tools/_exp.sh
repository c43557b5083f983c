set -x
FASTLEM_TRACE=1 timeout 800 python tools/profile_run.py --sites 16000000 --lattice --brief --repeat 2 2>&1 | grep -v "set_graph: free\|copies"
