#!/bin/bash
# multi-GPU session: DDP parity test (2 ranks), then weak- and strong-scaling bench lines.   usage: gpu_multi.sh N [configs]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_round2.py -k "ddp" -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_ddp.log 2>&1
echo "== ddp test: exit $? :: $(tail -1 gpurun_out/test_ddp.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/test_ddp.log | head
run() { # tag, args...
  local tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
  echo "== $tag rc=$?"; python - gpurun_out/bench_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("  value %.3e  ms/step %.3f  e2e %.3e (%.2f ms)  allreduce %.1f us (%.1f %% of step) check %s  launches/step/rank %.1f" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], 1e3 * d["allreduce"]["ms_per_step"],
        100 * d["allreduce"]["share_of_step"], d["allreduce"]["check"], d["gpu_launches_per_step_per_rank"]))
except Exception as e:
    print("  parse failed", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run ${N}gpu_weak_B128 --steps 10 --warmup 3
run ${N}gpu_strong_B128 --steps 10 --warmup 3 --scaling strong --batch 128
run ${N}gpu_strong_B8 --steps 50 --warmup 5 --scaling strong --batch 8
if [[ "${2:-}" == "configs" ]]; then
  run ${N}gpu_dense300h4 --steps 5 --warmup 3 --config dense300h4
  run ${N}gpu_global64k --steps 5 --warmup 3 --config global64k
fi
