#!/bin/bash
# parity suite, then A/B of the fused reduce + streaming update + next operand preparation launch
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2; do for v in 0 1; do
env BHMC_FUSED_STREAM=$v $B 2>gpurun_out/b38_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('FUSED_STREAM=$v value=%.0f launches=%s ms/step=%.2f group_ms=%s clocks=%s'%(d['value'], d.get('gpu_launches'), d['ms_per_step'], r.get('group_ms'), d['clocks']['sm_mhz']))"
done; done
