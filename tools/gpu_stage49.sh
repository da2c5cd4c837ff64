#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > $O/bench_r01g.json 2> $O/bench_r01g.err; echo "bench rc=$?"
OMP_NUM_THREADS=1 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r01g_ref.json 2> $O/bench_r01g_ref.err; echo "ref rc=$?"; cut -c1-200 $O/bench_r01g_ref.json
python tools/bench_extra.py mlp --chains 16 --steps 20 2>>$O/extra49.err | tee $O/extra49_mlp16.json
python tools/bench_extra.py mlp --chains 64 --steps 10 2>>$O/extra49.err | tee $O/extra49_mlp64.json
python tools/bench_extra.py sgld 2>>$O/extra49.err | tee $O/extra49_sgld.json
python tools/bench_extra.py sghmc 2>>$O/extra49.err | tee $O/extra49_sghmc.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tc_bgemm" -s 12 -c 3 -f -o $O/ncu_mlp16_bgemm2 python tools/bench_extra.py mlp --chains 16 --steps 3 > $O/ncu_mlp16_2.log 2>&1; echo "ncu rc=$?"
