#!/bin/bash
# Runs on the GPU box: links every edge_bwd variant object under gpurun_scratch/variants/ (built here with
# -DTG_TUNE_DEFAULT_ONLY plus whatever -D switch is being compared, seconds each) against the in-tree objects and times the default workload with each
# (TECGAT_LIB).  ptxas schedules the two edge loops of edge_bwd differently after ANY change to the kernel (+-4 %).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
OBJS=$(ls tec_mollm_b200/build/*.o | grep -v "/edge_bwd.o")
for v in gpurun_scratch/variants/edge_bwd_t*.o; do
  t=$(basename $v .o | sed 's/edge_bwd_t//')
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/lib_t$t.so $OBJS $v -cudart static 2>/dev/null || { echo "link $t failed"; continue; }
  for rep in 1 2; do
    TECGAT_LIB=/tmp/lib_t$t.so timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > /tmp/b_$t.json 2> /tmp/b_$t.err || { echo "bench $t failed"; tail -2 /tmp/b_$t.err; continue; }
    python - $t <<'PY'
import json, sys
d = json.load(open("/tmp/b_%s.json" % sys.argv[1]))
print("TG_TUNE=%-4s step %.3f ms  edge_bwd %.3f  edge_fwd %.3f" % (sys.argv[1], d["ms_per_step"], d["phases"]["edge_bwd"]["ms"], d["phases"]["edge_fwd"]["ms"]), flush=True)
PY
  done
done 2>&1 | tee gpurun_out/tune_edge_bwd.log
