#!/bin/bash
# multi-GPU checks (run with gpurun --gpus 2)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_distributed.py -q 2>&1 | tail -2
timeout 600 $TR --master-port 29531 tools/run_dist.py --log2 22 24 --T 50 2>&1 | grep -E "C5|rror" | tail -6
$TR --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default x2:', d['value'], d['scaling'], d['e2e']['value'], d['n_gpus'])"
$TR --master-port 29523 bench.py --gpus 2 --workload c5 --steps 3 --warmup 3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 x2:', d['value'], d['scaling'], d['ms_per_step'])"
