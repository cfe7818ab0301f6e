#!/bin/bash
OUT=gpurun_out/r2_qmh_chains_sv.txt
: > $OUT
run() { echo "== python examples/run_qmh.py $*" | tee -a $OUT; timeout 900 python examples/run_qmh.py "$@" 2>&1 | grep -v "Numerical problems\|true_divide" | tail -8 | tee -a $OUT; }
run sv --iters 3000 --burn-in 500 --particles 65536 --obs 1000
run sv_leverage --iters 2000 --burn-in 400 --particles 65536 --obs 1000 --step-size 0.05
