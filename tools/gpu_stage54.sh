#!/bin/bash
# programmatic dependent launch on the GEMM kernels (opt-in): parity under BHMC_PDL=1, then A/B
O=gpurun_out
BHMC_PDL=1 timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -2
for rep in 1 2; do for v in 0 1; do
echo -n "PDL=$v sgld: "; BHMC_PDL=$v timeout 120 python tools/bench_extra.py sgld 2>/dev/null | cut -c100-200
done; done
for v in 0 1; do
BHMC_PDL=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-ess --no-cpu-baseline --no-pixels 2>$O/b54_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL=$v value=%.0f ms/step=%.2f clocks=%s'%(d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
done
