#!/bin/bash
# one ncu --set full capture of selected kernels of a short bench run.  usage: gpu_ncu_kernel.sh <kernel regex> <tag> [bench args...]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KRE=$1; TAG=$2; shift 2
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 gpurun_out/build.log; }
timeout 600 $CMD > gpurun_out/bench_pre_$TAG.json 2> gpurun_out/bench_pre_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/bench_pre_$TAG.err; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 6 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
