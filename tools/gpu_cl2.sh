#!/bin/bash
for args in "--N 1000 --opt cluster=8" "--N 1000 --opt cluster=16" "--N 1000 --opt cluster=8 --no-smoother" "--N 2000 --opt cluster=16" "--N 4096 --opt cluster=16" "--N 1000 --opt cluster=8 --model sv" "--N 1000 --opt cluster=8 --dtype float32"; do
  timeout 120 python tools/run_once.py --T 300 $args 2>&1 | tail -1
done
export PFB200_LIB=$PWD/build_variants/prof/libpfb200.so
for args in "--N 1000 --opt cluster=8" "--N 1000 --opt cluster=16 --no-smoother"; do
  echo "== $args"; timeout 120 python tools/run_once.py --T 300 $args 2>&1 | tail -19 | grep -v "  -  "
done
