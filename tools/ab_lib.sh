#!/bin/bash
# A/B of two builds of the library on one box: bash tools/ab_lib.sh TAG ALT.so [WORKLOAD]  (current, ALT, current, ALT;
# with a third argument also bench.py --workload WORKLOAD: phase, multiview, train)
# A fourth argument skips the render line.
# ALT.so: a build of another revision copied aside before the call (e.g. tools/probes/_lib_prev.so; *.so is
# git-ignored but travels with the snapshot).
TAG=${1:-x}; ALT=$2
O=gpurun_out; mkdir -p $O
LIB=fresnel_b200/csrc/libfresnel_b200.so
cp $LIB /tmp/_cur.so
for v in cur alt cur alt; do
  if [ $v = alt ]; then cp $ALT $LIB; else cp /tmp/_cur.so $LIB; fi
  if [ -z "$4" ]; then
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads > $O/ab_${TAG}_$v.json 2>$O/ab_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/ab_${TAG}_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "serial", round(d["e2e"]["serial"]["value"],1), "; stages", {k:v for k,v in d["roofline"]["stage_ms"].items()})
PY
  fi
  if [ -n "$3" ]; then
    python bench.py --workload $3 --steps 20 --warmup 5 > $O/ab_${TAG}_$3_$v.json 2>>$O/ab_${TAG}.err
    python -c "
import json
d=json.loads(open('$O/ab_${TAG}_$3_$v.json').read().strip().splitlines()[-1]); print('$v $3', round(d['value'],2), round(d['ms_per_step'],4), d.get('stage_ms', d.get('roofline',{}).get('stage_ms')))"
  fi
done
cp /tmp/_cur.so $LIB
tail -3 $O/ab_${TAG}.err
