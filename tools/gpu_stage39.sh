#!/bin/bash
# exact-operand (2-MMA) path: parity suite, then dense vs 8-bit-pixel data
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline"
for rep in 1 2; do for v in dense pixels; do
$B --data $v 2>gpurun_out/b39_$v.err | tee gpurun_out/b39_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('data=$v value=%.0f ms/step=%.2f group_ms=%s warm=%s clocks=%s x=%s'%(d['value'], d['ms_per_step'], r.get('group_ms'), {k:(round(v,1) if isinstance(v,float) else v) for k,v in r['warmup_group_ms'].items() if k!='per'}, d['clocks']['sm_mhz'], d['config']['x_operand']))"
done; done
tail -3 gpurun_out/b39_pixels.err
