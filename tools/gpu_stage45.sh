#!/bin/bash
for c in 4 8 16 24 32 48 64; do echo -n "chains=$c: "; python tools/profile_grad.py --chains $c --evals 6 2>&1 | tail -1; done
