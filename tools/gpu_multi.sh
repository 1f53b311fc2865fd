#!/bin/bash
# Multi-GPU check: bash tools/gpu_multi.sh TAG N   (default bench line incl. the workloads block, under torchrun)
TAG=${1:-x}; N=${2:-2}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "peer or two_gpu or multi" > $O/test_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $O/test_${TAG}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err; echo "bench N=$N rc=$?"; tail -5 $O/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("render", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["host_link_ceiling"]["value"],1))
    for k,w in d.get("workloads",{}).items():
        print(" ", k, round(w["value"],1), w["unit"], "ms", round(w["ms_per_step"],3), "exchange", w.get("exchange",{}))
except Exception as e: print("bench ERR", e)
PY
