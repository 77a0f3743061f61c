#!/bin/bash
# A/B harness for edge_bwd source variants on the default workload (one B200).  Build the variant objects HERE (5 s each: only the
# default workload's kernel is instantiated), e.g.
#   mkdir -p gpurun_scratch/variants
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
#        -I tec_mollm_b200/csrc -I include -DTG_TUNE_DEFAULT_ONLY -D<SWITCH> -c tec_mollm_b200/csrc/edge_bwd.cu \
#        -o gpurun_scratch/variants/edge_bwd_t<name>.o
# then `gpurun -- bash tools/tune_edge_bwd.sh`: every object is linked against the in-tree objects on the box and timed twice
# through bench.py (TECGAT_LIB selects the library).  ptxas compiles each kernel on its own, so the trimmed object's SASS is
# identical to the full build's (checked).  Findings: profiles/r02_item_schedule.md.
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
