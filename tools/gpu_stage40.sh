#!/bin/bash
# final-state measurements of the round: bench lines (ours dense / pixels, reference arm), launch lists, ncu --set full
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench_r01f.json 2> $O/bench_r01f.err; echo "bench rc=$?"; cut -c1-400 $O/bench_r01f.json
python bench.py --steps 20 --warmup 3 --data pixels --no-cpu-baseline > $O/bench_r01f_pixels.json 2> $O/bench_r01f_pixels.err; echo "bench pixels rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r01f_ref.json 2> $O/bench_r01f_ref.err; echo "ref rc=$?"
B="python bench.py --steps 4 --warmup 2 --no-e2e --no-ess --no-cpu-baseline --no-pixels"
for v in dense pixels; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file $O/launches_r01f_$v.csv $B --data $v > $O/ncu_l_$v.log 2>&1; echo "launch list $v rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_fwd2|k_tc_gemm" -s 2 -c 2 -f -o $O/ncu_r01f_pixels_gemm python tools/profile_grad.py --evals 3 --data pixels > $O/ncu_f1.log 2>&1; echo "full gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_reduce_stream" -s 40 -c 1 -f -o $O/ncu_r01f_reduce_stream $B > $O/ncu_f2.log 2>&1; echo "full reduce rc=$?"
ls -la $O/*.ncu-rep
