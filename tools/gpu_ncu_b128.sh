set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 gpurun_out/build.log; }
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/bench_b128_pre.json 2> gpurun_out/bench_b128_pre.err || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"edge_|project_|reduce_columns|seed_advance|embed_" -s 30 -c 12 -f -o /tmp/prof_b128 $CMD > gpurun_out/ncu_b128.log 2>&1
echo "ncu full rc=$?"
# the report exceeds what gpurun copies back: summarise it on the box
python tools/ncu_summary.py /tmp/prof_b128.ncu-rep gpurun_out/r02_ncu_full_B128 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b128.csv $CMD > gpurun_out/ncu_launches_b128.log 2>&1
echo "ncu launches rc=$?"
