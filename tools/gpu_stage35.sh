#!/bin/bash
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2; do for sl in 0 100 500 2000; do
env BHMC_EPI_SLEEP=$sl $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; w=r['warmup_group_ms']; print('EPI_SLEEP=$sl value=%.0f total=%.0f  warmup fwd=%.1f bwd=%.1f clocks=%s'%(d['value'], r['group_ms']['step_total'], w['fwd'], w['bwd'], d['clocks']['sm_mhz']))"
done; done
