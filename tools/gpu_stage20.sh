#!/bin/bash
for rep in 1 2; do for f in 0 1; do for cpt in 0 8; do echo "FUSED=$f CPT=$cpt"; BHMC_FUSED_STEP=$f BHMC_CPT=$cpt timeout 300 python tools/bench_extra.py sgld --epochs 30 2>&1 | tail -1 | cut -c100-260; done; done; done
