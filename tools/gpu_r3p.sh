#!/bin/bash
for v in "$@"; do
  echo "== variant $v"
  export PFB200_LIB=build_variants/$v/libpfb200.so
  timeout 200 python tools/run_once.py --T 300 2>&1 | tail -20
  timeout 200 python tools/run_once.py --T 300 --dtype float32 2>&1 | tail -20
done
