#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (render + train), ncu launch list and one --set full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/test_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/test_$TAG.log
python bench.py --steps 50 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --workload train --steps 30 --warmup 5 > $O/train_$TAG.json 2>> $O/bench_$TAG.err; echo "train rc=$?"
python bench.py --workload train_full --steps 20 --warmup 5 > $O/train_full_$TAG.json 2>> $O/bench_$TAG.err; echo "train_full rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -f -o $O/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("bench_$TAG","train_$TAG","train_full_$TAG"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), d.get("roofline",{}).get("stage_ms"))
    except Exception as e: print(f, "ERR", e)
PY
