#!/bin/bash
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tc_fwd2|k_tc_gemm" -s 2 -c 2 -f -o $O/ncu_c8_gemm python tools/profile_grad.py --chains 8 --evals 3 > $O/ncu_c8.log 2>&1; echo rc=$?
