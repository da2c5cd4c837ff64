#!/bin/bash
# pair backward (cta_group::2) on the exact-operand path: parity, then A/B on pixel data
O=gpurun_out
BHMC_BWD2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "exact or pixels or full_size or ragged" 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline --no-pixels --data pixels"
for rep in 1 2; do for v in 0 1; do
BHMC_BWD2=$v $B 2>$O/b50_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('BWD2=$v value=%.0f ms/step=%.2f warm=%s clocks=%s'%(d['value'], d['ms_per_step'], {k:(round(v,1) if isinstance(v,float) else v) for k,v in r['warmup_group_ms'].items() if k!='per'}, d['clocks']['sm_mhz']))"
done; done
python tools/bench_extra.py sgld 2>>$O/extra50.err | cut -c1-260
python tools/bench_extra.py sgld 2>>$O/extra50.err | cut -c1-260
