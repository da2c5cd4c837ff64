#!/bin/bash
# N-GPU visit (gpurun --gpus N): multi-GPU parity of both sharding modes, then the bench line the driver's scaling
# run produces (chain-sharded headline + cfg3 / cfg4 blocks + the row-sharded cfg5 block with its in-run checks).
N=${1:-2}; shift
O=gpurun_out; mkdir -p $O
P=$((29500 + RANDOM % 500))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/multi_gpu_check.py > $O/multi_gpu_check_n$N.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "row-sharded|chain-sharded|MULTI_GPU_CHECK|Error|error" $O/multi_gpu_check_n$N.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 1)) bench.py --gpus $N "$@" > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench rc=$?"; tail -3 $O/bench_n$N.err
python - $O/bench_n$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d value %.0f e2e %.0f" % (d["n_gpus"], d["value"], d.get("e2e", {}).get("value", 0)))
for k in ("cfg3", "cfg4", "cfg5_row_sharded"):
    b = d.get(k, {})
    print(k, {x: b.get(x) for x in ("value", "ms_per_step", "ms_per_grad_eval_all_chains", "check_grad_vs_1rank", "check_replicas_identical", "check_moved", "error") if x in b})
PY
