#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -m gpu -q -x --timeout 600 2>&1 | tail -15
for k in 3 10 20; do for sc in lockstep streaming; do
python bench.py --steps $k --warmup 3 --schedule $sc --no-cpu-baseline --no-e2e --no-ess > gpurun_out/bench_${sc}_$k.json 2>gpurun_out/bench_${sc}_$k.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${sc}_$k.json')); r=d['roofline']; print('$sc K=$k value=%.0f ms/step=%.1f frac=%.3f chains/launch=%.1f'%(d['value'],d['ms_per_step'],r['frac'],r['chains_per_launch_mean']), {k:round(v,1) for k,v in r['group_ms'].items()}, d['clocks']['sm_mhz'], d['config']['schedule'])" || tail -5 gpurun_out/bench_${sc}_$k.err
done; done
