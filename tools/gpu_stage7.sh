#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -s 2>&1 | grep -E "chain [0-9]+: max err|passed|failed|Error|error" | head -20
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_cfg2.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f, 'value=%.0f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), r.get('group_ms'), d.get('clocks'), d.get('ess'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
