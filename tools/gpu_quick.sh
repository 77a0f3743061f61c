#!/bin/bash
# quick GPU check of selected test files + the default bench line.   usage: gpu_quick.sh "<pytest args>" [bench]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest $1 -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/test_quick.log 2>&1
echo "== tests: exit $? :: $(tail -1 gpurun_out/test_quick.log)"
grep -E "^(FAILED|ERROR)|bf16 training|Error" gpurun_out/test_quick.log | head -40
if [[ "${2:-}" == "bench" ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "== bench: exit $?"
  tail -3 gpurun_out/bench_tc.err
  python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_tc.json"))
    print("value %.3e edge-msgs/s  ms/step %.3f  e2e %.3e (%.2f ms/step, h2d %.0f MB)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e6))
    for k, v in d["phases"].items(): print("  %-9s %7.3f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
    print("  other", json.dumps(d.get("other_configs")))
except Exception as e: print("bench parse failed", e)
PY
fi
