python tools/profile_run.py --sites 1000000 --brief --max-iter 3 > /dev/null
for o in "rebuild_height=150" "rebuild_height=200" "rebuild_height=300" "rebuild_height=300 --opt rebuild_growth=8" "rebuild_height=500 --opt rebuild_growth=8" "rebuild_height=500 --opt rebuild_growth=16"; do
timeout 120 python tools/profile_run.py --sites 1000000 --brief --repeat 2 --opt $o | tail -1
done
