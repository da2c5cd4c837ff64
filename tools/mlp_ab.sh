# A/B of the MLP's batched GEMM variants: per-GEMM durations from ncu launch lists (tools/mlp_gemm_times.py).
# usage: bash tools/mlp_ab.sh "<tag>:<ENV=V ENV=V>" ...      (CH = chains, default 16)
set -u
O=gpurun_out; mkdir -p $O
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/l_$tag.csv python tools/bench_extra.py mlp --chains ${CH:-16} --steps 3 > $O/l_$tag.log 2>&1
  echo "== $tag ($envs) rc=$?"; python tools/mlp_gemm_times.py $O/l_$tag.csv
done
