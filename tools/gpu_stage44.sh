#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x --timeout 600 2>&1 | tail -2
BHMC_BG_BN=auto timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for rep in 1 2; do
echo "default:";            python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/mlp44.err | tee $O/mlp44.json
echo "SPLIT_TR=32:";        BHMC_SPLIT_TR=32 python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/mlp44.err
echo "BG_BN=auto:";         BHMC_BG_BN=auto python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/mlp44.err
echo "BG_BN=256:";          BHMC_BG_BN=256 python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/mlp44.err
done
