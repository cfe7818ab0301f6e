bash tools/profile_bench.sh r1e f32 | tail -1
for W in c1 c3 c4; do for D in f64 f32; do
  python bench.py --workload $W --dtype $D --steps 5 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/r1e_${W}_${D}_bench.json
  python -c "import json; d=json.load(open('gpurun_out/r1e_${W}_${D}_bench.json')); print('$W $D', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"
done; done
python bench.py > gpurun_out/r1e_default_bench.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r1e_default_bench.json')); print('default', d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac'], d['cpu_baseline']['value']/1e6)"
python __graft_entry__.py smoke 2>&1 | tail -1
