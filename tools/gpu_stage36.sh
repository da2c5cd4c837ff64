#!/bin/bash
for e in 1e-7 2e-7 4e-7; do for L in 4 8 15 30; do
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --ess-eps $e --ess-L $L --ess-steps 120 --ess-burnin 40 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['ess']; print('eps=$e L=$L ess_min/s=%.1f med/s=%.1f accept=%.3f secs=%.2f'%(e['ess_min_per_s'],e['ess_median_per_s'],e['mean_accept_prob'],e['seconds']))"
done; done
