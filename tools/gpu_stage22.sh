#!/bin/bash
B="python bench.py --workload cfg2-small --steps 40 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for cfg in "" "BENCH_NO_KTIMING=1"; do
env BHMC_PROF_HOST=1 $cfg $B 2>gpurun_out/b22.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$cfg] value=%.0f total=%.0f launches=%d'%(d['value'], r['group_ms']['step_total'], d['gpu_launches']))"; grep "prof host" gpurun_out/b22.err
done
