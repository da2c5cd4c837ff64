#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5
python tools/bench_extra.py mlp --chains 16 --steps 20 2>$O/mlp43.err | tee $O/mlp43.json
BHMC_MLP_SKINNY=0 python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/mlp43.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file $O/launches_mlp16b.csv python tools/bench_extra.py mlp --chains 16 --steps 6 > $O/ncu_mlp16b.log 2>&1
python tools/summarise_launches.py $O/launches_mlp16b.csv
