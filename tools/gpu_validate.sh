#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the bench line of both arms
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_final.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['pixel_data']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-160
