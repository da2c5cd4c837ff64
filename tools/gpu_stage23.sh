#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -m gpu -q -x --timeout 600 2>&1 | tail -5
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-ess --no-cpu-baseline --schedule lockstep"
for z in 0 1 0 1; do
env BHMC_ZCACHE=$z $B 2>gpurun_out/b23.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ZCACHE=$z lockstep value=%.0f total=%.0f'%(d['value'], r['group_ms']['step_total']), r['warmup_group_ms'])" || tail -5 gpurun_out/b23.err
done
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-ess --no-cpu-baseline --path-mode shared"
for z in 0 1; do
env BHMC_ZCACHE=$z $B 2>gpurun_out/b23.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ZCACHE=$z shared value=%.0f total=%.0f'%(d['value'], r['group_ms']['step_total']), r['warmup_group_ms'])" || tail -5 gpurun_out/b23.err
done
