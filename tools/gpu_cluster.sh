#!/bin/bash
# cluster kernel: quick timings first (hang-safe), then the parity tier
for N in 1000 2000 4096 16384; do
  for C in -1 1 2 4 8 16; do
    timeout 120 python tools/run_once.py --T 300 --N $N --opt cluster=$C 2>&1 | tail -1
  done
done
timeout 120 python tools/run_once.py --T 300 --N 1000 --model sv 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 300 --N 1000 --dtype float32 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
