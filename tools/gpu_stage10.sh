#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for ew in 16 24; do echo "EW=$ew"; BHMC_EPI_WARPS=$ew python tools/profile_grad.py --evals 6 | tail -2; BHMC_EPI_WARPS=$ew python tools/profile_grad.py --evals 6 --precision bf16 | tail -1; done
for ew in 16 24; do BHMC_EPI_WARPS=$ew python bench.py --steps 3 --warmup 3 --path-mode shared --no-cpu-baseline --no-e2e --no-ess > gpurun_out/bench_ew$ew.json 2>gpurun_out/bench_ew$ew.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ew$ew.json')); print('EW=$ew value=%.0f'%d['value'], d['roofline']['group_ms'], d['clocks'])"; done
