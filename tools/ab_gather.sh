#!/bin/bash
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
FRB_GATHER=1 timeout 300 python -m pytest tests -m gpu -x -q -k "tile_renderer or config2 or batched"  > $O/test_$TAG.log 2>&1; echo "pytest(gather) rc=$?"; tail -4 $O/test_$TAG.log
for v in 1 0 1 0; do
  FRB_GATHER=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads > $O/ab_${TAG}_$v.json 2>$O/ab_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/ab_${TAG}_$v.json").read().strip().splitlines()[-1])
print("FRB_GATHER=$v", round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "; stages", {k:v for k,v in d["roofline"]["stage_ms"].items()})
PY
done
tail -3 $O/ab_${TAG}.err
