#!/bin/bash
for b in 0 1; do
BHMC_BWD2=$b timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 10 --csv --log-file gpurun_out/l31_$b.csv python tools/profile_grad.py --evals 5 > /dev/null 2>&1
echo "BWD2=$b"; grep -E "^\"[0-9]" gpurun_out/l31_$b.csv | awk -F'","' '{printf "%-50s %s\n", substr($5,1,50), $NF}' | tr -d '"'
done
