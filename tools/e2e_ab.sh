# A/B of an environment switch on the headline bench (value / e2e):  bash tools/e2e_ab.sh <ENV_VAR> <v1> <v2> ...
var=$1; shift
for v in "$@"; do echo "== $var=$v"; env $var=$v python bench.py --steps ${STEPS:-8} --warmup 3 --no-ess --no-cpu-baseline --no-pixels --blocks "" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f launches %s  fwd/bwd sampled ms %.1f / %.1f  sm %s MHz' % (d['value'], d['e2e']['value'], d['gpu_launches'], r['group_ms']['fwd_sampled'], r['group_ms']['bwd_sampled'], d['clocks']['sm_mhz']))"; done
