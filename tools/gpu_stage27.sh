#!/bin/bash
for c in 64 56 48 40 32 24 16 8 4 1; do echo -n "chains=$c  "; BHMC_PROF=1 python tools/profile_grad.py --evals 4 --chains $c 2>&1 | grep -E "prof bwd|kernel-only" | tail -2 | sed -e 's/.*n_split/n_split/' -e 's/| MMA.*per chunk:/per chunk:/' | tr '\n' ' '; echo; done
