#!/bin/bash
# 2 GPUs: chain-sharded bench line (pixel_data block has collectives on every rank) + row-shard parity check
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n2_r01f.json 2> $O/bench_n2_r01f.err; echo "bench n2 rc=$?"
cut -c1-300 $O/bench_n2_r01f.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_n2_ref.json 2>> $O/bench_n2_r01f.err; echo "ref n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py > $O/multi_check_f.log 2>&1; echo "multi_gpu_check rc=$?"; tail -5 $O/multi_check_f.log
