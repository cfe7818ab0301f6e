#!/bin/bash
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -${TAILN:-1}; }
for v in v5 v6; do
  echo "== variant $v"; export PFB200_LIB=build_variants/$v/libpfb200.so
  run --T 300; run --T 300 --dtype float32; run --T 100 --N 4096 --B 1024
  run --T 300 --N 65536 --model sv --opt cluster=-1; run --T 300 --N 4096 --opt cluster=-1
done
echo "== profile C4"; export PFB200_LIB=build_variants/v5_prof/libpfb200.so
TAILN=19 run --T 300 --N 65536 --model sv --opt cluster=-1
