
bash tools/profile_bench.sh r1d f32 | tail -3
for W in c1 c3 c4; do for D in f64 f32; do
  python bench.py --workload $W --dtype $D --steps 5 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/r1d_${W}_${D}_bench.json
  python -c "import json; d=json.load(open('gpurun_out/r1d_${W}_${D}_bench.json')); print('$W $D', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"
done; done
python bench.py > gpurun_out/r1d_default_bench.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r1d_default_bench.json')); print('default', d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac'], d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_reference_arm.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r1d_reference_arm.json')); print('reference arm', d['value']/1e6, d['cpu_baseline']['cores'])"
python __graft_entry__.py smoke 2>&1 | tail -2
