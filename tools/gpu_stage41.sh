#!/bin/bash
O=gpurun_out
python tools/bench_extra.py mlp --chains 16 --steps 20 2>$O/mlp41.err | tee $O/mlp41.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file $O/launches_mlp16.csv python tools/bench_extra.py mlp --chains 16 --steps 6 > $O/ncu_mlp16.log 2>&1
python tools/summarise_launches.py $O/launches_mlp16.csv
