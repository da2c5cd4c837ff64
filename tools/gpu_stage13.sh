#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4
echo "== cfg5 / cfg5-half / cfg2"
timeout 300 python tools/profile_grad.py --workload cfg5 --evals 4 2>&1 | tail -2
timeout 300 python tools/profile_grad.py --workload cfg5-half --evals 4 2>&1 | tail -2
timeout 300 python tools/profile_grad.py --evals 6 2>&1 | tail -2
