#!/bin/bash
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 90 --csv --log-file $O/launches_sgld.csv python tools/bench_extra.py sgld --epochs 2 > $O/ncu_sgld.log 2>&1; echo rc=$?
python tools/summarise_launches.py $O/launches_sgld.csv
