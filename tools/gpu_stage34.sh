#!/bin/bash
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2; do for f in 1 0; do
env BHMC_FWD2=$f $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; w=r['warmup_group_ms']; print('FWD2=$f value=%.0f total=%.0f  warmup fwd=%.1f bwd=%.1f clocks=%s'%(d['value'], r['group_ms']['step_total'], w['fwd'], w['bwd'], d['clocks']['sm_mhz']))"
done; done
