#!/bin/bash
# A/B of the backward GEMM kernels at cfg2 size: row-slab kernels (BHMC_BWD_SK=0) vs swapped roles + stream-K (=2),
# kernel-only times per chains-per-launch, the in-kernel counters of the MMA thread, and the half-item weight sweep.
for sk in 0 2; do for c in ${CHAINS:-64 52 48 40 32 26 24}; do
  echo -n "BWD_SK=$sk chains=$c: "; BHMC_BWD_SK=$sk python tools/profile_grad.py --chains $c --evals 8 2>&1 | tail -1
done; done
for wh in ${WHS:-5 7 10}; do echo -n "BWD_SK=2 WH=$wh chains=64: "; BHMC_BWD_SK=2 BHMC_SK_WH=$wh python tools/profile_grad.py --chains 64 --evals 8 2>&1 | tail -1; done
for sk in 0 2; do BHMC_BWD_SK=$sk BHMC_PROF=1 python tools/profile_grad.py --chains 64 --evals 2 2>&1 | grep "prof bwd" | tail -1; done
