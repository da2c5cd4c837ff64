#!/bin/bash
mkdir -p gpurun_out
echo "== cfg5 on one GPU: full vs half rows"; 
timeout 300 python tools/profile_grad.py --workload cfg5 --evals 4 2>&1 | tail -3
timeout 300 python tools/profile_grad.py --workload cfg5-half --evals 4 2>&1 | tail -3
echo "== MLP"; timeout 300 python tools/bench_extra.py mlp --chains 16 2>&1 | tail -1
timeout 300 python tools/bench_extra.py sgld 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -c 1500 gpurun_out/bench_r01e.json
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_fwd2|k_tc_gemm|k_tc_reduce|k_tc_prep" -s 4 -c 4 -o gpurun_out/prof_tc_r01e python tools/profile_grad.py --evals 4 > gpurun_out/ncu_full_e.log 2>&1; tail -2 gpurun_out/ncu_full_e.log
