set -u
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/l_$tag.csv python tools/bench_extra.py mlp --chains ${CH:-16} --steps 3 > $O/l_$tag.log 2>&1
  echo "== $tag rc=$?"; python tools/mlp_gemm_times.py $O/l_$tag.csv; }
run base A=1
run base_noepi BHMC_BG_DEBUG_EPI=1
run bg2 BHMC_BG2=1
run bg2_noepi BHMC_BG2=1 BHMC_BG_DEBUG_EPI=1
run bg2_256 BHMC_BG2=256
run bg2_256_noepi BHMC_BG2=256 BHMC_BG_DEBUG_EPI=1
