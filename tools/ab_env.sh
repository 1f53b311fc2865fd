#!/bin/bash
# A/B of an environment toggle on one box: bash tools/ab_env.sh TAG VAR [A B]  (the GPU tests, then VAR=A,B,A,B;
# A B default to 1 0).  With A given the GPU tests run under VAR=A.
TAG=${1:-x}; VAR=${2:-FRB_OVERLAP}; A=${3:-1}; B=${4:-0}
export $VAR=$A
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/test_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/test_$TAG.log
for v in $A $B $A $B; do
  env $VAR=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads > $O/ab_${TAG}_$v.json 2>$O/ab_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/ab_${TAG}_$v.json").read().strip().splitlines()[-1])
print("$VAR=$v", round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "serial", round(d["e2e"]["serial"]["value"],1), "; stages", {k:v for k,v in d["roofline"]["stage_ms"].items()})
PY
done
tail -3 $O/ab_${TAG}.err
