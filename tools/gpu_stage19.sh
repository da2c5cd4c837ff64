#!/bin/bash
echo "== ESS sweep"
for e in 5e-8 1e-7 2e-7 3e-7; do for L in 100 30; do
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --ess-eps $e --ess-L $L --ess-steps 60 --ess-burnin 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['ess']; print('eps=$e L=$L ess_min/s=%.1f med/s=%.1f accept=%.3f secs=%.2f'%(e['ess_min_per_s'],e['ess_median_per_s'],e['mean_accept_prob'],e['seconds']))"
done; done
echo "== sgld launch list"
timeout 300 python tools/bench_extra.py sgld --epochs 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 100 --csv --log-file gpurun_out/launches_sgld.csv python tools/bench_extra.py sgld --epochs 1 > gpurun_out/ncu_sgld.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_sgld.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4][:40]; agg.setdefault(k,[]).append(float(r[-1]))
tot=0
for k,v in agg.items():
    print('%-42s n=%3d avg=%.2f us'%(k,len(v),sum(v)/len(v)/1e3)); tot+=sum(v)
print('sum per step (5 launches): %.1f us'%(tot/1e3/(len(rows)/5.0)))
PY
