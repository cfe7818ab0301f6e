timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 120 python tools/run_once.py --T 200 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --dtype f32 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --N 2000 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --N 65536 --model sv 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 50 --N 16777216 2>&1 | tail -1
