#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2 3; do for v in 0 1; do
env BHMC_BN_LAST=$v $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; w=r['warmup_group_ms']; print('BN_LAST=$v value=%.0f total=%.0f  warmup fwd=%.1f bwd=%.1f clocks=%s'%(d['value'], r['group_ms']['step_total'], w['fwd'], w['bwd'], d['clocks']['sm_mhz']))"
done; done
for c in 20 36 52; do for v in 0 1; do echo -n "BN_LAST=$v chains=$c: "; BHMC_BN_LAST=$v python tools/profile_grad.py --chains $c --evals 5 2>&1 | tail -1; done; done
