#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full
# capture of the dominant kernel for the same command (B200_PROFILING.md recipe).
set -u
TAG=${1:-r1}
DTYPE=${2:-f64}
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --dtype $DTYPE"
$CMD > $OUT/${TAG}_${DTYPE}_plain.json 2> $OUT/${TAG}_${DTYPE}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_${DTYPE}_plain.err; exit 1; }
cat $OUT/${TAG}_${DTYPE}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_${DTYPE}_launches.csv $CMD > $OUT/${TAG}_${DTYPE}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_persistent -s 3 -c 1 -o $OUT/${TAG}_${DTYPE}_full -f $CMD > $OUT/${TAG}_${DTYPE}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_${DTYPE}_ncu_full.log
