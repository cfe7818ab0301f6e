#!/bin/bash
export PFB200_LIB=$PWD/build_variants/prof/libpfb200.so
for args in "--N 1000 --opt cluster=8" "--N 1000 --opt cluster=8 --no-smoother" "--N 1000 --opt cluster=1 --no-smoother" "--N 4096 --opt cluster=16"; do
  echo "== $args"; timeout 120 python tools/run_once.py --T 300 $args 2>&1 | tail -19
done
