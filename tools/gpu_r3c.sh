#!/bin/bash
# staged phase A: timings with / without staging, then the GPU test tier
mkdir -p gpurun_out
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -1; }
echo "== base (see r3_stage1.log)"
unset PFB200_LIB
echo "== new, stage=1"
run --T 300; run --T 300 --dtype float32; run --T 100 --N 4096 --B 1024; run --T 300 --N 65536 --model sv --opt cluster=-1
echo "== new, stage=0"
run --T 300 --opt stage=0; run --T 300 --dtype float32 --opt stage=0; run --T 100 --N 4096 --B 1024 --opt stage=0
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
