#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for b in 0 1; do echo "BWD2=$b"; BHMC_BWD2=$b python tools/profile_grad.py --evals 6 2>&1 | tail -1; BHMC_BWD2=$b BHMC_PROF=1 python tools/profile_grad.py --evals 2 2>&1 | grep "prof bwd" | tail -1 | cut -c1-420; done
for w in cfg5-half; do for b in 0 1; do echo -n "BWD2=$b $w: "; BHMC_BWD2=$b python tools/profile_grad.py --workload $w --evals 4 2>&1 | tail -1; done; done
for c in 48 16 4; do for b in 0 1; do echo -n "BWD2=$b chains=$c: "; BHMC_BWD2=$b python tools/profile_grad.py --chains $c --evals 5 2>&1 | tail -1; done; done
