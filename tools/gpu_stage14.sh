#!/bin/bash
for h in 0 1; do echo "== L2HINT=$h"; for w in cfg5 cfg5-half cfg2; do BHMC_L2HINT=$h timeout 300 python tools/profile_grad.py --workload $w --evals 5 2>&1 | tail -1; done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
