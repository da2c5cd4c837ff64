#!/bin/bash
# tensor-map cache: parity (maps reused across evaluations, shapes and windows), then A/B on the launch-bound workloads
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
for rep in 1 2; do for v in 0 1; do
echo -n "MAP_CACHE=$v sgld: "; BHMC_MAP_CACHE=$v python tools/bench_extra.py sgld 2>/dev/null | cut -c100-200
done; done
for v in 0 1; do echo -n "MAP_CACHE=$v mlp16: "; BHMC_MAP_CACHE=$v python tools/bench_extra.py mlp --chains 16 --steps 20 2>/dev/null | cut -c118-220; done
