#!/bin/bash
# one full ncu capture of the persistent kernel at the latency-bound configurations (C1, C4)
OUT=gpurun_out
timeout 200 python tools/run_once.py --T 300 --N 1000 | tail -1
timeout 200 python tools/run_once.py --T 300 --N 65536 --model sv | tail -1
ncu --set full --clock-control none --import-source on -k regex:pf_persistent -s 1 -c 1 -o $OUT/r2_c1_full -f python tools/run_once.py --T 300 --N 1000 > $OUT/r2_c1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_persistent -s 1 -c 1 -o $OUT/r2_c4_full -f python tools/run_once.py --T 300 --N 65536 --model sv > $OUT/r2_c4_ncu.log 2>&1
tail -2 $OUT/r2_c4_ncu.log
