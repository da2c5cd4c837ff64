#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_extra.py rows --steps 3 > gpurun_out/extra_rows_g1.json 2> gpurun_out/extra_rows_g1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_extra.py rows --steps 3 > gpurun_out/extra_rows_g2.json 2> gpurun_out/extra_rows_g2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1
grep -E "row-sharded|chain-sharded|MULTI_GPU" gpurun_out/multi_check.log
tail -n 2 gpurun_out/extra_rows_g1.json gpurun_out/extra_rows_g2.json; tail -n 5 gpurun_out/extra_rows_g1.err gpurun_out/extra_rows_g2.err
