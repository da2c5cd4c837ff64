#!/bin/bash
run() { echo -n "XT=$1 DM=$2 PAD=$3 chains=$4: "; BHMC_XT_SLAB=$1 BHMC_DM_SLAB=$2 BHMC_SLAB_PAD=$3 python tools/profile_grad.py --evals 6 --chains $4 2>&1 | tail -1 | sed 's/kernel-only (CUDA events around the launches)://'; }
for c in 64 16; do
run 8192 8192 64 $c
run 64 64 0 $c
run 64 8192 0 $c
run 8192 64 0 $c
run 512 512 0 $c
run 64 64 64 $c
done
BHMC_XT_SLAB=64 BHMC_DM_SLAB=64 BHMC_SLAB_PAD=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
