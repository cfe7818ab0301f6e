#!/bin/bash
# A/B of library variants (build_variants/NAME): C2 fp64 / fp32, C3 slice, C4 shape, N = 4096 on 8 CTAs
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -1; }
for v in "$@"; do
  echo "== variant $v"; export PFB200_LIB=build_variants/$v/libpfb200.so
  run --T 300; run --T 300 --dtype float32; run --T 100 --N 4096 --B 1024
  run --T 300 --N 65536 --model sv --opt cluster=-1; run --T 300 --N 4096 --opt cluster=-1
done
