#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "rc=$?" >> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --steps 2 --warmup 3 --path-mode shared --no-cpu-baseline > gpurun_out/bench_cfg2_shared.json 2> gpurun_out/bench_cfg2_shared.err
timeout 600 python bench.py --steps 2 --warmup 3 --path-mode shared --no-cpu-baseline --no-e2e --precision bf16 > gpurun_out/bench_cfg2_bf16.json 2> gpurun_out/bench_cfg2_bf16.err
timeout 600 python bench.py --steps 1 --warmup 1 --path-mode shared --no-cpu-baseline --no-e2e --precision fp32 > gpurun_out/bench_cfg2_fp32.json 2> gpurun_out/bench_cfg2_fp32.err
timeout 300 python tools/profile_grad.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_grad.py > gpurun_out/ncu_launch.log 2>&1
timeout 300 python tools/profile_grad.py > gpurun_out/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 2 -c 2 -o gpurun_out/prof_tc_r01 python tools/profile_grad.py > gpurun_out/ncu_full.log 2>&1
for f in t_all.log bench_cfg2.json bench_cfg2.err bench_cfg2_shared.json bench_cfg2_bf16.json bench_cfg2_fp32.json profile_plain.log; do echo "== $f"; tail -n 6 gpurun_out/$f; done
