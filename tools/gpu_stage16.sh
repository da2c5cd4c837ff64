#!/bin/bash
nvidia-smi --query-gpu=name,serial,power.limit,clocks.max.sm --format=csv,noheader
for rep in 1 2; do for sl in 8192 100000000 65536 2048; do echo "== DM_SLAB=$sl"; for w in cfg2 cfg5-half; do BHMC_DM_SLAB=$sl timeout 300 python tools/profile_grad.py --workload $w --evals 6 2>&1 | tail -1; done; done; done
