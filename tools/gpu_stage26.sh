#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1
grep -E "row-sharded|chain-sharded|MULTI_GPU|Error|error" gpurun_out/multi_check.log | head
timeout 600 python tools/bench_extra.py rows --steps 3 2>gpurun_out/extra_rows_g1.err | tail -1 > gpurun_out/extra_rows_g1.json; cat gpurun_out/extra_rows_g1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_extra.py rows --steps 3 2> gpurun_out/extra_rows_g2.err | tail -1 > gpurun_out/extra_rows_g2.json; cat gpurun_out/extra_rows_g2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('2-GPU value=%.0f e2e=%.0f ms/step=%.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'], d.get('ess',{}).get('ess_min_per_s'))"
timeout 300 python tools/bench_extra.py sgld --epochs 30 | tail -1 > gpurun_out/extra_sgld.json; cat gpurun_out/extra_sgld.json
timeout 300 python tools/bench_extra.py mlp --chains 16 | tail -1 > gpurun_out/extra_mlp.json; cat gpurun_out/extra_mlp.json
