#!/bin/bash
# What the driver runs at round end, in one go: GPU tests, smoke, reference arm, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_ref.json','gpurun_out/bench_n1.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f, 'value=%.1f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'frac', r.get('frac'), r.get('group_ms'), d.get('clocks'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'ess', (d.get('ess') or {}).get('ess_median_per_s'), 'launches', d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
