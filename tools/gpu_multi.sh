cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
