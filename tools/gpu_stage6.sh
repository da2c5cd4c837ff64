#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check.log
grep -E "row-sharded|chain-sharded|MULTI_GPU|rc=|Error|error" gpurun_out/multi_check.log | head -20
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -n 4 gpurun_out/t_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ['gpurun_out/bench_n1.json','gpurun_out/bench_n2.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f, 'value=%.0f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), r.get('group_ms'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
