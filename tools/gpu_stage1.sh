#!/bin/bash
# First GPU visit: checker path first, tensor-core path in its own process (a trap there must not
# hide the fp32 results). Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "fp32 or mvn or philox or nan" > gpurun_out/t_fp32.log 2>&1
echo "fp32 rc=$?" >> gpurun_out/t_fp32.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "bf16" > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/t_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/t_fp32.log gpurun_out/t_tc.log gpurun_out/smoke.log
