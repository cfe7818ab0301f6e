#!/bin/bash
# north-star evidence: complete particle-qMH chains on the B200 estimator (C1 / C4 shapes)
OUT=gpurun_out/r2_qmh_chains.txt
: > $OUT
run() { echo "== python examples/run_qmh.py $*" | tee -a $OUT; timeout 900 python examples/run_qmh.py "$@" 2>&1 | grep -v "Numerical problems" | tail -8 | tee -a $OUT; }
run lgss --iters 10000 --burn-in 1000 --particles 2000 --obs 500
run lgss --iters 10000 --burn-in 1000 --particles 1000 --obs 250
run lgss --iters 3000 --burn-in 500 --particles 2000 --obs 500 --alg mh1
run sv --iters 2000 --burn-in 500 --particles 65536 --obs 1000
run sv_leverage --iters 2000 --burn-in 500 --particles 65536 --obs 1000 --step-size 0.1
