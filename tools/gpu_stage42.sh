#!/bin/bash
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_bgemm|k_mlp_gemm|k_mlp_colsum|k_split_operand" -s 60 -c 24 -f -o $O/ncu_mlp16 python tools/bench_extra.py mlp --chains 16 --steps 4 > $O/ncu_mlp16f.log 2>&1; echo rc=$?
ls -la $O/ncu_mlp16.ncu-rep
