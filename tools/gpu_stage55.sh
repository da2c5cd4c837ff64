#!/bin/bash
BHMC_PDL_MLP=1 timeout 60 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x --timeout 50 2>&1 | tail -1
for v in 0 1; do echo -n "PDL_MLP=$v: "; BHMC_PDL_MLP=$v timeout 40 python tools/bench_extra.py mlp --chains 16 --steps 10 2>/dev/null | cut -c118-200; done
