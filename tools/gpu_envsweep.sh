cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  step %.3f ms  ' % d['ms_per_step'] + '  '.join('%s %.3f' % (k, v['ms']) for k,v in d['phases'].items()))"; }
for cfg in "$@"; do run $cfg; done
