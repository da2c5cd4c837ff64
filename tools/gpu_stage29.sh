#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
for w in cfg5 cfg5-half cfg2; do python tools/profile_grad.py --workload $w --evals 5 2>&1 | tail -1; done
python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value=%.0f ms/step=%.1f frac=%.3f ess=%.0f'%(d['value'],d['ms_per_step'],r['frac'],d['ess']['ess_min_per_s']), r['warmup_group_ms'])"
python tools/bench_extra.py sgld --epochs 30 | tail -1 | cut -c90-250
