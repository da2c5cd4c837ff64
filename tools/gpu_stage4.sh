#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -n 12 gpurun_out/t_all.log
for pair in 1 0; do
BHMC_PAIR=$pair timeout 600 python bench.py --steps 2 --warmup 3 --path-mode shared --no-cpu-baseline --no-e2e > gpurun_out/bench_shared_pair$pair.json 2> gpurun_out/bench_shared_pair$pair.err
BHMC_PAIR=$pair timeout 600 python bench.py --steps 2 --warmup 3 --path-mode shared --no-cpu-baseline --no-e2e --precision bf16 > gpurun_out/bench_bf16_pair$pair.json 2> gpurun_out/bench_bf16_pair$pair.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_pair*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, 'value=%.0f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'dom_avg_ms=%.4f'%r['avg_launch_ms'], r['group_ms'], d['clocks'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
