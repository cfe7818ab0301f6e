#!/bin/bash
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -${TAILN:-1}; }
for o in 1 0; do
  echo "== stage=$o"
  run --T 300 --opt stage=$o; run --T 300 --dtype float32 --opt stage=$o; run --T 100 --N 4096 --B 1024 --opt stage=$o
  run --T 300 --N 65536 --model sv --opt cluster=-1 --opt stage=$o; run --T 300 --N 4096 --opt cluster=-1 --opt stage=$o
done
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
