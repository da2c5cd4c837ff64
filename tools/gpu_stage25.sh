#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -m gpu -q -x --timeout 600 2>&1 | tail -15
for z in 0 1; do for k in 3 20; do
env BHMC_ZCACHE=$z python bench.py --steps $k --warmup 3 --no-e2e --no-ess --no-cpu-baseline 2>gpurun_out/b25.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ZCACHE=$z K=$k value=%.0f total=%.0f frac=%.3f'%(d['value'], r['group_ms']['step_total'], r['frac']), d['config']['schedule'], r['warmup_group_ms'])" || tail -5 gpurun_out/b25.err
done; done
