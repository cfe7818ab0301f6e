python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python examples/run_qmh.py lgss --iters 1000 --burn-in 200 2>&1 | tail -3
python examples/run_qmh.py lgss --iters 1000 --burn-in 200 --particles 1000 --alg mh1 2>&1 | tail -2
python examples/run_qmh.py sv --iters 300 --burn-in 50 2>&1 | tail -2
python tools/run_once.py --T 200 2>&1 | tail -1
