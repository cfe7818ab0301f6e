#!/bin/bash
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -1; }
export PFB200_LIB=build_variants/base/libpfb200.so
echo "== base: smoother share"
for dt in float64 float32; do
  run --T 300 --dtype $dt --no-smoother
  run --T 300 --dtype $dt --L 1
  run --T 300 --dtype $dt --L 5
  run --T 300 --dtype $dt --L 10
  run --T 300 --dtype $dt --L 20
done
