# A/B of a persistent-minibatch-kernel switch on BASELINE config 3:  bash tools/cfg3_ab.sh <ENV_VAR> <v1> <v2> ...   (STEPS = headline steps, default 1)
var=$1; shift
for v in "$@"; do echo "== $var=$v"; env $var=$v BHMC_PROF=${PROF:-1} python bench.py --blocks cfg3 --steps ${STEPS:-1} --warmup 1 --no-pixels --no-e2e --no-ess --no-cpu-baseline 2> gpurun_out/cfg3_$var$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['cfg3']
print(b.get('error') or {k:(round(b[k]['value']), round(1e3*b[k]['ms_per_step'],1), b[k]['steps']) for k in ('sgld','sghmc')})"; grep "prof persist\] 1" gpurun_out/cfg3_$var$v.err | tail -1; done
