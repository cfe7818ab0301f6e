#!/bin/bash
# ncu capture of the batched workload (C3): plain run first, then launch list + one full capture
OUT=gpurun_out
CMD="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu"
$CMD > $OUT/r1_c3_plain.json 2> $OUT/r1_c3_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r1_c3_launches.csv $CMD > $OUT/r1_c3_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:pf_persistent -s 3 -c 1 -o $OUT/r1_c3_full $CMD > $OUT/r1_c3_ncu_full.log 2>&1
ncu -i $OUT/r1_c3_full.ncu-rep --page raw --csv > $OUT/r1_c3_raw.csv 2>/dev/null
rm -f $OUT/r1_c3_full.ncu-rep
tail -1 $OUT/r1_c3_ncu_full.log
