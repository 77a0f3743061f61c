#!/bin/bash
# launch list + one ncu --set full capture of the edge kernels for BASELINE configs 4 and 5; summaries are made ON the box (the
# reports exceed what gpurun copies back)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for c in dense300h4 global64k; do
  CMD="python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/bench_pre_$c.json 2> gpurun_out/bench_pre_$c.err || { echo "plain $c failed"; continue; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$c.csv $CMD > gpurun_out/ncu_launches_$c.log 2>&1; echo "launch list $c rc=$?"
  timeout 1200 ncu --set full --clock-control none -k regex:"edge_bwd|edge_fwd" -s 6 -c 2 -f -o /tmp/prof_$c $CMD > gpurun_out/ncu_full_$c.log 2>&1; echo "ncu full $c rc=$?"
  python tools/ncu_summary.py /tmp/prof_$c.ncu-rep gpurun_out/ncu_full_$c | tail -1
done
