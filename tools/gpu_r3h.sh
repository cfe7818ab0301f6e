#!/bin/bash
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -${TAILN:-1}; }
for v in v5 i15 i11; do
  echo "== variant $v"; export PFB200_LIB=build_variants/$v/libpfb200.so
  run --T 300; run --T 300 --dtype float32
done
