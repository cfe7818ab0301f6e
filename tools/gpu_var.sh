timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 120 python tools/run_once.py --T 200 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 1024 --dtype f32 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 50 --N 16777216 2>&1 | tail -1
