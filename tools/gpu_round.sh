#!/bin/bash
# Runs on the B200 box under gpurun: parity tests (isolated pytest processes so one CUDA fault cannot poison the rest),
# a short bench with each projection implementation, and -- only if the plain bench exited 0 -- the ncu launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || echo "BUILD FAILED"
run() { # name, env, pytest args...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 900 python -m pytest "$@" -q -m gpu -x --no-header -p no:cacheprovider > $OUT/test_$name.log 2>&1
  echo "== $name: exit $? :: $(tail -1 $OUT/test_$name.log)"
}
run plan "TECGAT_PROJ=ffma" tests/test_gpu_gatv2.py -k "plan"
run proj_ffma "TECGAT_PROJ=ffma" tests/test_gpu_gatv2.py -k "projection and ffma"
run proj_tc "TECGAT_PROJ=tc" tests/test_gpu_gatv2.py -k "projection and tc"
run fused_ffma "TECGAT_PROJ=ffma" tests/test_gpu_gatv2.py -k "not plan and not projection"
run fused_tc "TECGAT_PROJ=tc" tests/test_gpu_gatv2.py -k "not plan and not projection"
run graph "TECGAT_PROJ=tc" tests/test_gpu_graph.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "== smoke: exit $? :: $(tail -1 $OUT/smoke.log)"
TECGAT_PROJ=ffma timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_ffma.json 2> $OUT/bench_ffma.err; echo "== bench ffma: exit $?"; tail -c 1500 $OUT/bench_ffma.json
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_tc.json 2> $OUT/bench_tc.err; rc=$?; echo "== bench tc: exit $rc"; tail -c 1500 $OUT/bench_tc.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "== bench reference: exit $?"
if [ "${1:-}" = "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline > $OUT/bench_small.json 2> $OUT/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  echo "== ncu launch list: exit $?"
fi
