#!/bin/bash
# C1 (N = 1000, T = 500: the cluster kernel): plain bench first, then the ncu launch list and one full capture
set -u
OUT=gpurun_out
CMD="python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu"
$CMD > $OUT/r2_c1_plain.json 2> $OUT/r2_c1_plain.err || { echo "plain run failed"; tail -5 $OUT/r2_c1_plain.err; exit 1; }
cat $OUT/r2_c1_plain.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_c1_launches.csv $CMD > $OUT/r2_c1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_cluster -s 3 -c 1 -o $OUT/r2_c1_full -f $CMD > $OUT/r2_c1_ncu_full.log 2>&1
tail -2 $OUT/r2_c1_ncu_full.log
