#!/bin/bash
# graph builder session: parity tests, timing of both grids, one ncu --set full capture of the count kernel on the global grid
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "BUILD FAILED"; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_graph.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_graph.log 2>&1
echo "== graph tests: exit $? :: $(tail -1 gpurun_out/test_graph.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/test_graph.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $? :: $(tail -1 gpurun_out/smoke.log)"
cat > /tmp/gb.py <<'PY'
import json, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
for name in ("default", "dense300h4", "global64k"):
    ei, leg = bench.graph_build_leg(bench.CONFIGS[name], dev, with_cpu=(name != "global64k") or True)
    print(json.dumps({"config": name, **leg}), flush=True)
PY
timeout 900 python /tmp/gb.py > gpurun_out/graph_build.jsonl 2> gpurun_out/graph_build.err; echo "== graph_build rc=$?"; cat gpurun_out/graph_build.jsonl | cut -c1-900; tail -2 gpurun_out/graph_build.err
cat > /tmp/gb1.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from tec_mollm_b200 import graph
lat, lon = np.linspace(-89.5, 89.5, 180), np.linspace(-179.5, 179.5, 360)
for _ in range(2):
    ei, ew = graph.build_graph(lat, lon, 150.0, device="cuda:0")
torch.cuda.synchronize(); print(ei.shape)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edges_kernel -s 2 -c 2 -f -o gpurun_out/prof_haversine python /tmp/gb1.py > gpurun_out/ncu_haversine.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_haversine.log
