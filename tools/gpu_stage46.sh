#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4
for c in 8 16 32 64; do for v in 0 1; do echo -n "XA_BLOCKED=$v chains=$c: "; BHMC_XA_BLOCKED=$v python tools/profile_grad.py --chains $c --evals 6 2>&1 | tail -1; done; done
B="python bench.py --steps 20 --warmup 3 --no-ess --no-cpu-baseline --no-pixels"
for rep in 1 2; do for v in 0 1; do
BHMC_XA_BLOCKED=$v $B 2>$O/b46_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('XA_BLOCKED=$v value=%.0f e2e=%.0f ms/step=%.2f warm=%s clocks=%s'%(d['value'], d['e2e']['value'], d['ms_per_step'], {k:(round(v,1) if isinstance(v,float) else v) for k,v in r['warmup_group_ms'].items() if k!='per'}, d['clocks']['sm_mhz']))"
done; done
