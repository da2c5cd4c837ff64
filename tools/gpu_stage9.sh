#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-ess --no-cpu-baseline --path-mode shared"
$B > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches_r01_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
python tools/profile_grad.py --evals 4 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tc_gemm|k_tc_reduce|k_tc_prep" -s 4 -c 4 -o gpurun_out/prof_tc_r01c python tools/profile_grad.py --evals 4 > gpurun_out/ncu_full.log 2>&1
python tools/bench_extra.py update > gpurun_out/extra_update.json 2>/dev/null &&
ncu --set full --clock-control none -k regex:"k_hmc_update|k_sgld|k_accept|k_hmc_begin" -s 3 -c 12 -o gpurun_out/prof_update_r01 python tools/bench_extra.py update > gpurun_out/ncu_upd.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01_bench.csv; tail -3 gpurun_out/ncu_launch.log
