#!/bin/bash
# full GPU tier + the workload bench lines (no CPU leg)
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
for w in c2 c1 c3 c4; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 > $OUT/cur_${w}.json 2>> $OUT/cur.err
  python - $OUT/cur_${w}.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:40], "value %.3f G"%(d["value"]/1e9), "e2e %.3f G"%(d["e2e"]["value"]/1e9), "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"], d["geometry"])
PY
done
timeout 300 python bench.py --dtype f32 --no-cpu --steps 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32 value %.3f G frac %.4f'%(d['value']/1e9, d['roofline']['frac']))"
