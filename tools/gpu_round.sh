#!/bin/bash
# Runs on the B200 box under gpurun: parity tests (isolated pytest processes so one CUDA fault cannot poison the rest),
# benches, and -- only after the plain command exited 0 -- ncu captures.   usage: gpu_round.sh [tests] [bench] [ncu] [ncufull]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
WHAT=" ${*:-tests bench} "
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 $OUT/build.log; }
run() { # name, env, pytest args...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 1200 python -m pytest "$@" -q -m gpu --no-header -p no:cacheprovider > $OUT/test_$name.log 2>&1
  echo "== $name: exit $? :: $(tail -1 $OUT/test_$name.log)"
  grep -E "^(FAILED|ERROR)" $OUT/test_$name.log | head -20
}
if [[ "$WHAT" == *" tests "* ]]; then
  run proj "TECGAT_PROJ=tc" tests/test_gpu_gatv2.py -k "plan or projection"
  run fused_ffma "TECGAT_PROJ=ffma" tests/test_gpu_gatv2.py -k "not plan and not projection"
  run fused_tc "TECGAT_PROJ=tc" tests/test_gpu_gatv2.py -k "not plan and not projection"
  run graph "TECGAT_PROJ=tc" tests/test_gpu_graph.py
  run round2 "TECGAT_PROJ=tc" tests/test_gpu_round2.py -s
  grep -E "unpinned|bf16 training" $OUT/test_round2.log | head -20
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "== smoke: exit $? :: $(tail -1 $OUT/smoke.log)"
fi
if [[ "$WHAT" == *" bench "* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_tc.json 2> $OUT/bench_tc.err; echo "== bench tc: exit $?"
  python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_tc.json"))
    print("value %.3e edge-msgs/s  ms/step %.3f  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d["phases"].items(): print("  %-9s %7.3f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
    print("  clocks", d["clocks"], "cpu", d.get("cpu_baseline", {}).get("value"), "launches/step", d.get("gpu_launches_per_step_per_rank"))
    print("  other", json.dumps(d.get("other_configs")))
    print("  graph_build", json.dumps(d.get("graph_build")))
except Exception as e: print("bench parse failed", e)
PY
  timeout 600 python bench.py --steps 10 --warmup 3 --autocast --no-cpu-baseline > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "== bench bf16: exit $?"
  python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_bf16.json"))
    print("bf16: value %.3e  ms/step %.3f" % (d["value"], d["ms_per_step"]))
    for k, v in d["phases"].items(): print("  %-9s %7.3f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
except Exception as e: print("bench parse failed", e)
PY
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "== bench reference: exit $?"
fi
if [[ "$WHAT" == *" configs "* ]]; then
  for c in dense300h4 global64k; do
    timeout 900 python bench.py --config $c --steps 5 --warmup 3 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "== bench $c: exit $?"
    python - $c <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
    print("%s: value %.3e  ms/step %.3f  whole-path frac %.3f  graph_build %.2f ms" % (sys.argv[1], d["value"], d["ms_per_step"], d["whole_path_frac_of_roofline"], d["graph_build"]["ms"]))
    for k, v in d["phases"].items(): print("  %-9s %7.3f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
except Exception as e: print("bench parse failed", e)
PY
  done
fi
SMALL="python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline"
if [[ "$WHAT" == *" ncu "* ]]; then
  timeout 600 $SMALL > $OUT/bench_small.json 2> $OUT/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv $SMALL > $OUT/ncu_launches.log 2>&1
  echo "== ncu launch list: exit $?"
fi
if [[ "$WHAT" == *" ncufull "* ]]; then
  timeout 600 $SMALL > $OUT/bench_small.json 2> $OUT/bench_small.err &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"edge_|project_" -s 18 -c 5 -f -o $OUT/prof $SMALL > $OUT/ncu_full.log 2>&1
  echo "== ncu full: exit $?"
fi
