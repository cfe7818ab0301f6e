#!/bin/bash
OUT=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pf_cluster -s 1 -c 1 -o $OUT/r2_cl_n1000 -f python tools/run_once.py --T 300 --N 1000 --opt cluster=8 > $OUT/r2_cl_ncu.log 2>&1
tail -2 $OUT/r2_cl_ncu.log
