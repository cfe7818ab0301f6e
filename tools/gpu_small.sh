timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
for N in 600 1000 1024 1025 2000; do
timeout 120 python tools/run_once.py --T 200 --N $N 2>&1 | tail -1
done
python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/r1_c1_f64_bench.json; python -c "import json; d=json.load(open('gpurun_out/r1_c1_f64_bench.json')); print('c1', d['value']/1e9, d['ms_per_step'])"
