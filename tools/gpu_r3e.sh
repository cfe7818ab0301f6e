#!/bin/bash
bash tools/gpu_r3.sh "$@"
unset PFB200_LIB
echo "== C4 shape / small"
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -1; }
run --T 300 --N 65536 --model sv --opt cluster=-1
run --T 300 --N 4096 --opt cluster=-1
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
