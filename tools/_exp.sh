cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['phases'].items()})" "$TAG"; }
TAG="bf16 rt"; run --autocast
TAG="bf16 tc"; TECGAT_PROJ_BWD=tc run --autocast
TAG="fp32 tc"; TECGAT_PROJ_BWD=tc run
