#!/bin/bash
# A/B of the compositor backward variants on one box + ncu --set full of the hot kernels.
#   gpurun --timeout 1200 -- 'bash tools/ab_bwd.sh TAG'
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
for v in 0 1 0 1; do
  FRB_BWD_V1=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads > $O/ab_${TAG}_v$v.json 2>$O/ab_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/ab_${TAG}_v$v.json").read().strip().splitlines()[-1])
print("V1=$v", round(d["value"],1), "fps; stages", {k:v for k,v in d["roofline"]["stage_ms"].items()})
PY
done
ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel|composite_fwd_kernel|tile_rank_gather|tile_emit|tile_count|radix_onesweep" -s 40 -c 14 -f -o $O/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null; echo "raw rc=$?"
