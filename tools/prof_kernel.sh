#!/bin/bash
# ncu --set full of selected kernels of one bench invocation, after the same command ran clean without ncu.
#   bash tools/prof_kernel.sh TAG "<bench.py args>" "<kernel regex>" [skip] [count]
TAG=${1:-x}; ARGS=${2:---no-cpu-baseline --no-workloads --steps 3 --warmup 3}; RE=${3:-composite}; SKIP=${4:-0}; CNT=${5:-6}
O=gpurun_out; mkdir -p $O
python bench.py $ARGS > $O/prof_${TAG}_plain.json 2> $O/prof_${TAG}_plain.err; echo "plain rc=$?"; tail -c 600 $O/prof_${TAG}_plain.json
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o $O/prof_$TAG \
    python bench.py $ARGS > $O/prof_${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --csv > $O/prof_${TAG}_source.csv 2>/dev/null
python tools/summarise_raw.py $O/prof_${TAG}_raw.csv
