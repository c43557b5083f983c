timeout 100 python tools/profile_run.py --sites 1000000 --brief --repeat 2 | tail -1
timeout 100 python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt park_after=4 | tail -1
timeout 100 python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt park_after=16 | tail -1
