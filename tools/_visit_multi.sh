N=$1; shift
mkdir -p gpurun_out
R=r2v
echo "== bench N=$N $@"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 2 --warmup 3 "$@" > gpurun_out/${R}_bench_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; grep "^{" gpurun_out/${R}_bench_${N}gpu.json | head -c 1200; echo; tail -3 gpurun_out/${R}_bench_${N}gpu.err
