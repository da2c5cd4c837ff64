#!/bin/bash
# One parametrised script for the GPU-box visits (replaces the 55 one-off tools/gpu_stage*.sh of round 1).
# Everything is logged under gpurun_out/ (merged back by gpurun); nothing here is a bench value unless it says so.
#   tools/gpu_session.sh tests [pytest -k expr]      GPU test suite (what the driver runs), per-test durations
#   tools/gpu_session.sh smoke                        __graft_entry__.smoke()
#   tools/gpu_session.sh bench [bench.py args]        one bench line -> gpurun_out/bench_<tag>.json (+ summary)
#   tools/gpu_session.sh ref                          reference arm
#   tools/gpu_session.sh launches <tag> -- <cmd>      ncu launch list (gpu__time_duration.sum) of a command
#   tools/gpu_session.sh ncu <tag> <kernel regex> <skip> <count> -- <cmd>   ncu --set full capture of matching kernels
#   tools/gpu_session.sh ncuhw <tag> <kernel regex> <skip> <count> -- <cmd>  same with the hardware-counter sections only
#                                                     (no SASS patching: what a cooperative persistent kernel survives)
#   tools/gpu_session.sh ab <VAR> <v1,v2,..> -- <cmd> run a command once per value of an environment switch
#   tools/gpu_session.sh sanitize <tool> -- <cmd>     compute-sanitizer (memcheck | racecheck | synccheck | initcheck)
set -u
O=gpurun_out
mkdir -p $O
TAG=${BHMC_TAG:-r02}
sub=$1; shift
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("value %.0f  e2e %.0f  ms/step %.2f  roofline %s frac %.3f avg %.3f ms chains/launch %.1f  clocks %s" % (
    d["value"], d.get("e2e", {}).get("value", 0), d["ms_per_step"], r.get("kernel_group"), r.get("frac", 0), r.get("avg_launch_ms", 0),
    r.get("chains_per_launch_mean", 0), d.get("clocks")))
for k in ("pixel_data", "ess", "cfg3", "cfg4", "cfg5_row_sharded"):
    b = d.get(k)
    if not b: continue
    if "error" in b: print(k, "ERROR", b["error"]); continue
    if k == "ess": print("ess min/s %.0f median/s %.0f accept %.2f | small_n: %s" % (b["ess_min_per_s"], b["ess_median_per_s"], b["mean_accept_prob"], {x: b["small_n"][x] for x in ("ess_min_per_s", "ess_median_per_s", "mean_accept_prob", "seconds")} if "small_n" in b else None)); continue
    if k == "cfg3":
        for w in ("sgld", "sghmc"): print("cfg3", w, "%.3g grad-evals/s  %.1f us/step  frac %.3f  groups %s" % (b[w]["value"], 1e3 * b[w]["ms_per_step"], b[w]["roofline"]["frac"], {g: (round(v["avg_ms"] * 1e3, 1) if v["avg_ms"] else None) for g, v in b[w]["roofline"]["groups"].items()}))
        continue
    rr = b.get("roofline", {})
    print(k, "%.4g grad-evals/s  %.3f ms/step  frac %.3f (%s)  %s" % (b["value"], b["ms_per_step"], rr.get("frac", 0), rr.get("kernel_group"), {x: b[x] for x in ("ms_per_grad_eval_all_chains", "check_grad_vs_1rank", "check_replicas_identical", "block_wall_s") if x in b}))
if "cpu_baseline" in d: print("cpu", {k: (v if k != "ess" else {x: v[x] for x in ("ess_min_per_s", "steps", "seconds")}) for k, v in d["cpu_baseline"].items() if k != "sample"})
PY
}
case $sub in
  tests)
    timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=12 ${1:+-k "$1"} > $O/tests_$TAG.log 2>&1
    echo "pytest rc=$?"; tail -25 $O/tests_$TAG.log ;;
  smoke)
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$TAG.log ;;
  bench)
    name=${BHMC_BENCH_NAME:-bench_$TAG}
    timeout 900 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "bench rc=$?"; tail -3 $O/$name.err; summ $O/$name.json ;;
  ref)
    timeout 600 python bench.py --impl reference "$@" > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref_$TAG.json ;;
  launches)
    tag=$1; shift; [ "$1" = "--" ] && shift
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_$tag.csv "$@" > $O/launches_$tag.log 2>&1
    echo "ncu launches rc=$?"; python tools/summarise_launches.py $O/launches_$tag.csv 2>&1 | tail -25 ;;
  ncu)
    tag=$1; regex=$2; skip=$3; count=$4; shift 4; [ "$1" = "--" ] && shift
    timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $count -f -o $O/ncu_$tag "$@" > $O/ncu_$tag.log 2>&1
    echo "ncu full rc=$?"; tail -3 $O/ncu_$tag.log ;;
  ncuhw)
    tag=$1; regex=$2; skip=$3; count=$4; shift 4; [ "$1" = "--" ] && shift
    timeout 600 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis \
      --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
      --clock-control none -k "regex:$regex" -s $skip -c $count -f -o $O/ncu_$tag "$@" > $O/ncu_$tag.log 2>&1
    echo "ncu hw rc=$?"; tail -3 $O/ncu_$tag.log ;;
  ab)
    var=$1; vals=$2; shift 2; [ "$1" = "--" ] && shift
    for v in ${vals//,/ }; do echo "== $var=$v"; env $var=$v "$@" 2>&1 | tail -${BHMC_AB_TAIL:-2}; done ;;
  sanitize)
    tool=$1; shift; [ "$1" = "--" ] && shift
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 "$@" > $O/sanitize_${tool}_$TAG.log 2>&1
    echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|=========     [A-Z]" $O/sanitize_${tool}_$TAG.log | tail -8 ;;
  *) echo "unknown subcommand $sub"; exit 2 ;;
esac
