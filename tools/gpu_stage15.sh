#!/bin/bash
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for w in cfg5 cfg5-half; do
timeout 900 ncu --metrics $M --clock-control none -k regex:"k_tc_gemm|k_tc_fwd2" -s 2 -c 2 --csv --log-file gpurun_out/ncu_$w.csv python tools/profile_grad.py --workload $w --evals 3 > gpurun_out/ncu_$w.log 2>&1
cat gpurun_out/ncu_$w.csv | grep -v "^==" | cut -d, -f5,13- | head -30
done
