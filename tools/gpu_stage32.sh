#!/bin/bash
for pf in 0 2 4 8 16; do echo -n "PF=$pf: "; BHMC_PF=$pf python tools/profile_grad.py --evals 6 2>&1 | tail -1; BHMC_PF=$pf BHMC_PROF=1 python tools/profile_grad.py --evals 2 2>&1 | grep "prof bwd" | tail -1 | sed 's/.*mean over/mean over/' | cut -c1-230; done
for c in 32 40; do for b in 0 2; do echo -n "BWD2=$b chains=$c: "; BHMC_BWD2=$b python tools/profile_grad.py --chains $c --evals 5 2>&1 | tail -1; done; done
