BHMC_TAG=p1 tools/gpu_session.sh tests "persistent or sgld or sgd or cfg3"
for v in 1 0; do echo "== BHMC_PERSIST_PAIR=$v"; BHMC_PERSIST_PAIR=$v BHMC_PROF=1 python bench.py --blocks cfg3 --steps 1 --warmup 1 --no-pixels --no-e2e --no-ess --no-cpu-baseline 2> gpurun_out/cfg3_pair$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['cfg3']
print(b.get('error') or {k:(b[k]['value'], b[k]['ms_per_step']) for k in ('sgld','sghmc')})"; grep "prof persist" gpurun_out/cfg3_pair$v.err | tail -2; done
