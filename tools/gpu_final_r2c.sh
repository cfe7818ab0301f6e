#!/bin/bash
# last refresh of round 2 (no ncu): GPU test tier and the bench lines of every workload with the final library
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/r2_gputests_final.log; cat $OUT/r2_gputests_final.log
timeout 900 python bench.py > $OUT/r2_default_bench.json 2> $OUT/r2_default_bench.err; tail -c 300 $OUT/r2_default_bench.json
for w in c1 c3 c4; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 > $OUT/r2_${w}_f64_bench.json 2>> $OUT/r2_wl.err
done
timeout 300 python bench.py --workload c1 --dtype f32 --no-cpu --steps 5 > $OUT/r2_c1_f32_bench.json 2>> $OUT/r2_wl.err
timeout 300 python bench.py --workload c3 --dtype f32 --no-cpu --steps 5 > $OUT/r2_c3_f32_bench.json 2>> $OUT/r2_wl.err
timeout 300 python bench.py --dtype f32 --no-cpu --steps 5 > $OUT/r2_f32_bench.json 2>> $OUT/r2_wl.err
for f in default c1_f64 c1_f32 c3_f64 c3_f32 c4_f64 f32; do python - $OUT/r2_${f}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:44], d["dtype"], "value %.3f G"%(d["value"]/1e9), "e2e %.3f G"%(d["e2e"]["value"]/1e9), "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"], d["roofline"]["kernel"])
PY
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
