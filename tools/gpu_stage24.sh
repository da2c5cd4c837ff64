#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-ess --no-cpu-baseline --path-mode shared"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_softmax_from_z|k_tc_fwd2" -s 4 -c 8 --csv --log-file gpurun_out/launches_zc.csv $B > gpurun_out/ncu_zc.log 2>&1
grep -E "from_z|fwd2" gpurun_out/launches_zc.csv | cut -d, -f5,15- | cut -c1-40,80-
