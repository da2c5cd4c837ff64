#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -n 6 gpurun_out/t_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?" >> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --steps 3 --warmup 3 --path-mode shared --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg2_shared.json 2> gpurun_out/bench_cfg2_shared.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_cfg2.json','gpurun_out/bench_cfg2_shared.json','gpurun_out/bench_ref.json']:
    try:
        d=json.load(open(f)); r=d.get('roofline',{})
        print(f, 'value=%.0f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'launched', d['config'].get('grad_evals_launched_incl_masked'), r.get('group_ms'), d.get('clocks'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
