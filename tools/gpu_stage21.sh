#!/bin/bash
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2 3 4 5 6; do
env BHMC_PROF_HOST=1 $B 2>gpurun_out/b21.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value=%.0f total=%.0f frac=%.3f'%(d['value'], r['group_ms']['step_total'], r['frac']))" || tail -5 gpurun_out/b21.err
grep "prof host" gpurun_out/b21.err | tail -2 | cut -c1-150
done
