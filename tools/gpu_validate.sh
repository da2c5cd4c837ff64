#!/bin/bash
# Round-end validation on one B200: full GPU test suite, smoke(), both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 | tee gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']
print('value=%.0f ms/step=%.1f e2e=%.0f frac=%.3f traffic=%s chains/launch=%.1f'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['frac'],r['traffic'],r['chains_per_launch_mean']), r['group_ms'], d['clocks'], 'cpu', d['cpu_baseline']['value'], 'ess', d.get('ess',{}).get('ess_min_per_s'), 'launches', d['gpu_launches'])
print(json.load(open('gpurun_out/bench_ref.json'))['value'])
PY
B="python bench.py --steps 4 --warmup 2 --no-e2e --no-ess --no-cpu-baseline"
$B > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/launches_r01_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
python tools/profile_grad.py --evals 4 > gpurun_out/profile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_fwd2|k_tc_gemm|k_tc_reduce|k_tc_prep" -s 4 -c 4 -o gpurun_out/prof_tc_r01f python tools/profile_grad.py --evals 4 > gpurun_out/ncu_full.log 2>&1
BZ="python bench.py --steps 1 --warmup 1 --no-e2e --no-ess --no-cpu-baseline --path-mode shared"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_softmax_from_z|k_tc_fwd2" -s 6 -c 2 -o gpurun_out/prof_fromz_r01 $BZ > gpurun_out/ncu_fromz.log 2>&1
B2="python tools/bench_extra.py stream --steps 3"
timeout 300 $B2 > gpurun_out/extra_stream.json 2>gpurun_out/extra_stream.err &&
timeout 900 ncu --set full --clock-control none -k regex:"k_stream_update|k_accept|k_hmc_begin|k_stream_kinetic" -s 40 -c 12 -o gpurun_out/prof_stream_r01 $B2 > gpurun_out/ncu_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
