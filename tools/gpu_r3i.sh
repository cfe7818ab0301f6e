#!/bin/bash
run() { timeout 200 python tools/run_once.py "$@" 2>&1 | tail -1; }
for G in 148 100 64; do run --N 4194304 --T 50 --G $G --reps 1; done
for G in 148 64; do run --N 16777216 --T 50 --G $G --reps 1; done
