#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line (render + workloads block), measured parity margins,
# ncu launch list and (optionally) one --set full capture of the compositor kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG [full]'
TAG=${1:-x}
FULL=${2:-}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/test_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/test_$TAG.log
python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
python tools/diag_tolerances.py > $O/diag_$TAG.json 2> $O/diag_$TAG.err; echo "diag rc=$?"; tail -3 $O/diag_$TAG.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > $O/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -f -o $O/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
    print("render", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "serial", round(d["e2e"]["serial"]["value"],1))
    print(" stages", d["roofline"]["stage_ms"])
    for k,w in d.get("workloads",{}).items():
        print(" ", k, round(w["value"],1), w["unit"], "ms", round(w["ms_per_step"],3), "exchange", w.get("exchange",{}).get("ms"))
    print(" cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("bench ERR", e)
PY
cat $O/diag_$TAG.json
