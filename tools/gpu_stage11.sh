#!/bin/bash
# cta_group::2 forward kernel: parity first (bounded), then A/B timing against the multicast-pair kernel
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 2>&1 | tail -5
echo "pytest rc=$?"
for f in 0 1; do for prec in bf16x3 bf16; do echo "FWD2=$f $prec"; BHMC_FWD2=$f timeout 120 python tools/profile_grad.py --evals 6 --precision $prec | tail -2; done; done
